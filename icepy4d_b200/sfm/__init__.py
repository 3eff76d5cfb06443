from .absolute_orientation import Absolute_orientation  # noqa: F401
from .geometry import estimate_pose, undistort_points  # noqa: F401
from .triangulation import Triangulate  # noqa: F401
from .two_view_geometry import RelativeOrientation  # noqa: F401
