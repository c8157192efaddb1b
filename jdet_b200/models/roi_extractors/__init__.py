from .oriented_single_level import OrientedSingleRoIExtractor  # noqa: F401
from .rbox_single_level import RboxSingleRoIExtractor  # noqa: F401
