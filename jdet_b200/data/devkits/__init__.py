from . import dota_utils, result_merge  # noqa: F401
