"""Error taxonomy of the reference (src/camera/mod.rs:79-113, src/util/mod.rs:39-63)."""
from . import _native as N


class AcmError(RuntimeError):
    """CUDA / NCCL / argument failure inside libacm (no reference counterpart)."""


class CameraModelError(Exception):
    pass


class ProjectionOutSideImage(CameraModelError):
    def __init__(self):
        super().__init__("Projection is outside the image")


class PointIsOutSideImage(CameraModelError):
    def __init__(self):
        super().__init__("Input point is outside the image")


class PointAtCameraCenter(CameraModelError):
    def __init__(self):
        super().__init__("z is close to zero, point is at camera center")


class FocalLengthMustBePositive(CameraModelError):
    def __init__(self):
        super().__init__("Focal length must be positive")


class PrincipalPointMustBeFinite(CameraModelError):
    def __init__(self):
        super().__init__("Principal point must be finite")


class InvalidParams(CameraModelError):
    def __init__(self, msg):
        super().__init__(f"Invalid camera parameters: {msg}")
        self.detail = msg


class YamlError(CameraModelError):
    def __init__(self, msg):
        super().__init__(f"Failed to load YAML: {msg}")


class IOError_(CameraModelError):
    def __init__(self, msg):
        super().__init__(f"IO Error: {msg}")


class NumericalError(CameraModelError):
    def __init__(self, msg):
        super().__init__(f"NumericalError: {msg}")
        self.detail = msg


class UtilError(Exception):
    pass


class ZeroProjectionPoints(UtilError):
    def __init__(self):
        super().__init__("Zero projection points")


# per-point status byte -> the `Err(..)` the reference's scalar call returns
_NUMERICAL_MESSAGES = {1: "Jacobian is singular", 2: "Unprojection failed to converge"}


def raise_point_status(status: int, model_id: int = -1):
    if status == 0:
        return
    if status == 1:
        raise PointIsOutSideImage()
    if status == 2:
        raise PointAtCameraCenter()
    if status == 3:
        raise ProjectionOutSideImage()
    raise NumericalError(_NUMERICAL_MESSAGES.get(model_id, "numerical failure"))


def raise_call_status(rc: int, message: str):
    """Map a negative libacm return code to the reference's error variant."""
    if rc == N.OK:
        return
    if rc == N.ERR_INVALID_PARAMS:
        raise InvalidParams(message)
    if rc == N.ERR_NUMERICAL:
        raise NumericalError(message)
    if rc == N.ERR_FOCAL_LENGTH:
        raise FocalLengthMustBePositive()
    if rc == N.ERR_PRINCIPAL_POINT:
        raise PrincipalPointMustBeFinite()
    if rc == N.ERR_ZERO_PROJECTION_POINTS:
        raise ZeroProjectionPoints()
    raise AcmError(f"libacm error {rc}: {message}")
