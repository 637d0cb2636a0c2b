"""util hot loops of the reference: undistort_image (src/util/undistort.rs:14-105), sample_points
(src/util/point_sampling.rs:46-120), compute_reprojection_error (src/util/error_metrics.rs:62-121)."""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .camera import CameraModel, Intrinsics
from .errors import UtilError
from .runtime import Points

_lib = N.lib


class InterpolationMethod(enum.IntEnum):  # undistort.rs:8-12
    Nearest = 0
    Bilinear = 1


@dataclass
class ProjectionError:  # error_metrics.rs:17-31
    rmse: float
    min: float
    max: float
    mean: float
    stddev: float
    median: float
    count: int = 0


def _target_array(target: Intrinsics | None):
    if target is None:
        return None
    return (C.c_double * 4)(target.fx, target.fy, target.cx, target.cy)


def undistort_images(frames: np.ndarray, camera_model: CameraModel, target_intrinsics: Intrinsics | None = None,
                     interpolation: InterpolationMethod = InterpolationMethod.Bilinear) -> np.ndarray:
    """Batch form: frames (F, H, W, 3) uint8 -> same shape."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    if frames.ndim != 4 or frames.shape[3] != 3:
        raise UtilError("Invalid parameters: expected (F, H, W, 3) uint8 frames")
    res = camera_model.get_resolution()
    F, H, W, _ = frames.shape
    if W != res.width or H != res.height:  # undistort.rs:23-28
        raise UtilError(f"Invalid parameters: Image {W}x{H} doesn't match model {res.width}x{res.height}")
    ctx = camera_model.ctx
    cam = camera_model.camera_block()
    out = np.empty_like(frames)
    ctx.check(_lib.acm_undistort_rgb8_host(ctx.handle, C.byref(cam), _target_array(target_intrinsics), frames.ctypes.data_as(C.c_void_p),
                                           out.ctypes.data_as(C.c_void_p), F, int(interpolation)))
    return out


def undistort_image(input_image: np.ndarray, camera_model: CameraModel, target_intrinsics: Intrinsics | None = None,
                    interpolation: InterpolationMethod = InterpolationMethod.Bilinear) -> np.ndarray:
    """`undistort_image(&RgbImage, &dyn CameraModel, Option<Intrinsics>, InterpolationMethod)`;
    image is (H, W, 3) uint8 (image::RgbImage memory order)."""
    img = np.ascontiguousarray(input_image, dtype=np.uint8)
    return undistort_images(img[None], camera_model, target_intrinsics, interpolation)[0]


def undistort_map(camera_model: CameraModel, target_intrinsics: Intrinsics | None = None) -> np.ndarray:
    """(H, W, 2) source coordinates of every output pixel (NaN where the projection fails)."""
    ctx = camera_model.ctx
    cam = camera_model.camera_block()
    res = camera_model.get_resolution()
    n = res.width * res.height
    d = ctx.device_alloc(max(n, 1) * 16)
    ctx.check(_lib.acm_undistort_map(ctx.handle, C.byref(cam), _target_array(target_intrinsics), C.c_void_p(d)))
    out = np.empty((res.height, res.width, 2))
    ctx.d2h(out, d)
    ctx.sync()
    ctx.device_free(d)
    return out


def sample_points(camera_model: CameraModel, n: int, device: bool = False, shard: tuple[int, int] | None = None):
    """`sample_points(Some(&model), n) -> (Matrix2xX, Matrix3xX)`: (points_2d, points_3d).

    `shard=(rank, world)` returns that rank's contiguous slice of the grid (one slice per GPU; the
    slices concatenated in rank order are the unsharded result)."""
    if camera_model is None:
        raise UtilError("Camera model does not exist")  # the reference panics on None (point_sampling.rs:53)
    ctx = camera_model.ctx
    cam = camera_model.camera_block()
    uv_h, xyz_h, kept = C.c_void_p(), C.c_void_p(), C.c_size_t()
    rank, world = shard if shard is not None else (0, 1)
    ctx.check(_lib.acm_sample_points_shard(ctx.handle, C.byref(cam), n, rank, world, C.byref(uv_h), C.byref(xyz_h), C.byref(kept)))
    uv = Points(ctx, 2, kept.value, _handle=uv_h)
    xyz = Points(ctx, 3, kept.value, _handle=xyz_h)
    if device:
        return uv, xyz
    a, b = uv.numpy(), xyz.numpy()
    uv.free(); xyz.free()
    return a, b


def compute_reprojection_error(camera_model: CameraModel, points3d, points2d) -> ProjectionError:
    if camera_model is None:
        raise UtilError("Camera model does not exist")
    ctx = camera_model.ctx
    own = []
    def dev(a, dim):
        if isinstance(a, Points):
            return a
        p = Points.from_numpy(ctx, np.ascontiguousarray(a, dtype=np.float64).reshape(-1, dim))
        own.append(p)
        return p
    X, UV = dev(points3d, 3), dev(points2d, 2)
    cam = camera_model.camera_block()
    out = N.ProjectionError()
    rc = _lib.acm_reprojection_error(ctx.handle, C.byref(cam), X.handle, UV.handle, C.byref(out))
    for p in own:
        p.free()
    ctx.check(rc)
    return ProjectionError(out.rmse, out.min, out.max, out.mean, out.stddev, out.median, int(out.count))
