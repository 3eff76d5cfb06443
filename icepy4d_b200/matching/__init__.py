from .enums import GeometricVerification, Quality, TileSelection  # noqa: F401
from .geometric_verification import geometric_verification  # noqa: F401
from .matchers import FeaturesBase, ImageMatcherBase, LightGlueMatcher, SuperGlueMatcher  # noqa: F401
from .tiling import Tiler  # noqa: F401
