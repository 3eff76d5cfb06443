from .features import Feature, Features  # noqa: F401
from .points import Point, Points  # noqa: F401
