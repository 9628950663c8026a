from .CameraCalibration import CameraCalibration  # noqa: F401
from .LensDistortion import LensDistortion  # noqa: F401
from .PerspectiveCorrection import PerspectiveCorrection  # noqa: F401
