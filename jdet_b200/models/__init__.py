"""Thin mirrors of the reference modules that CALL the hot-path ops (jdet.models.*)."""
