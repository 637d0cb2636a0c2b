"""`camera_converter` on the GPU path (reference bin/camera_converter.rs): load one model from YAML,
sample correspondences from it, fit every other model (linear estimate -> Levenberg-Marquardt) and
report the reprojection statistics and the five-region validation.

    python -m apex_camera_models_b200.camera_converter -i kb -p samples/kannala_brandt.yaml -n 500
    torchrun --nproc-per-node 8 -m apex_camera_models_b200.camera_converter -i kb -p kb.yaml -n 10000000

Flags are the reference's (`-i/--input-model`, `-p/--input-path`, `-n/--num-points`, `-m/--image-path`;
camera_converter.rs:66-83).  Under torchrun every rank samples its slice of the grid and the fits run on
the sharded correspondences (normal equations and statistics are combined inside libacm).  The image
quality diagnostics of the reference (PSNR / SSIM of the rendered projection images)
run on the GPU as well (image_quality.py); `--image-path` is the optional reference image the projections
are drawn on, and the display images go to `<output-dir>/<model>_projection.png` like the reference's.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time
from dataclasses import dataclass, field

import numpy as np

from . import distributed as D
from .camera import (CameraModel, DoubleSphereModel, EucmModel, FovModel, Intrinsics, KannalaBrandtModel, PinholeModel, RadTanModel,
                     UcmModel)
from .errors import AcmError
from .image_quality import ImageQualityMetrics, compute_image_quality_metrics, model_projection_visualization
from .reporting import export_conversion_results, export_point_correspondences
from .optimization import OptimizationCost
from .runtime import Context, default_context
from .util import ProjectionError, compute_reprojection_error, sample_points

# camera_converter.rs:86-124
INPUT_ALIASES = {
    "kb": KannalaBrandtModel, "kannala_brandt": KannalaBrandtModel, "ds": DoubleSphereModel, "double_sphere": DoubleSphereModel,
    "radtan": RadTanModel, "rad_tan": RadTanModel, "ucm": UcmModel, "unified": UcmModel, "eucm": EucmModel,
    "extended_unified": EucmModel, "fov": FovModel, "field_of_view": FovModel, "pinhole": PinholeModel,
}
# target order and initial distortion of bin/camera_converter.rs (:364-370 DS, :507-511 KB, :644-650 RadTan, :787-792 UCM,
# :918-924 EUCM, :1051-1056 FOV)
TARGETS = [
    ("Double Sphere", DoubleSphereModel, [0.5, 0.1]),
    ("Kannala-Brandt", KannalaBrandtModel, [0.0, 0.0, 0.0, 0.0]),
    ("Radial-Tangential", RadTanModel, [0.0, 0.0, 0.0, 0.0, 0.0]),
    ("Unified Camera Model", UcmModel, [0.5]),
    ("Extended Unified Camera Model", EucmModel, [0.5, 1.0]),
    ("Field-of-View", FovModel, [1.0]),
]
# `output_model_name` handed to compute_image_quality_metrics -> output/<name>_projection.png (camera_converter.rs:469, :608, :750, :880, :1014, :1144)
IMAGE_LABELS = {DoubleSphereModel: "double_sphere_apex", KannalaBrandtModel: "kannala_brandt_apex", RadTanModel: "radial_tangential_apex",
                UcmModel: "unified_camera_model_apex", EucmModel: "extended_unified_camera_model_apex", FovModel: "fov_apex"}
REGIONS = [("Center", 0.5), ("Near Center", 0.55), ("Mid Region", 0.65), ("Edge Region", 0.8), ("Far Edge", 0.95)]  # validation.rs:106-112


@dataclass
class ValidationResults:  # validation.rs:21-42
    region_errors: list
    average_error: float
    max_error: float
    status: str
    region_data: list = field(default_factory=list)


@dataclass
class ConversionMetrics:  # reporting.rs ConversionMetrics
    model: CameraModel
    model_name: str
    final_reprojection_error: ProjectionError
    initial_reprojection_error: ProjectionError
    optimization_time_ms: float
    convergence_status: str
    validation_results: ValidationResults
    iterations: int = 0
    image_quality: "ImageQualityMetrics | None" = None   # reporting.rs:36-37


def load_input_model(model_type: str, path: str, ctx: Context | None = None) -> CameraModel:
    cls = INPUT_ALIASES.get(model_type.lower())
    if cls is None:
        raise ValueError(f"Unsupported input model type: {model_type}. Supported types: kb, ds, radtan, ucm, eucm, pinhole, fov")
    return cls.load_from_yaml(path, ctx=ctx)


def validate_conversion_accuracy(output_model: CameraModel, input_model: CameraModel) -> ValidationResults:
    """validation.rs:93-213: five test pixels on the diagonal -> unproject with the input model ->
    project with both -> distance; EXCELLENT < 0.001 px, GOOD < 0.1 px."""
    res = input_model.get_resolution()
    w, h = float(res.width), float(res.height)
    px = np.array([[w * f, h * f] for _, f in REGIONS])
    rays, st_u = input_model.unproject_batch(px)
    uv_in, st_a = input_model.project_batch(rays)
    uv_out, st_b = output_model.project_batch(rays)
    errors, data, total, worst, valid = [], [], 0.0, 0.0, 0
    for i, (name, _) in enumerate(REGIONS):
        if st_u[i] == 0 and st_a[i] == 0 and st_b[i] == 0:
            e = float(np.hypot(*(uv_in[i] - uv_out[i])))
            total += e; worst = max(worst, e); valid += 1
            errors.append(e)
            data.append((name, tuple(uv_in[i]), tuple(uv_out[i]), e))
        else:
            errors.append(math.nan)
            data.append((name, None, None, math.nan))
    avg = total / valid if valid else math.nan
    status = "NEEDS IMPROVEMENT" if math.isnan(avg) else "EXCELLENT" if avg < 0.001 else "GOOD" if avg < 0.1 else "NEEDS IMPROVEMENT"
    return ValidationResults(errors, avg, worst, status, data)


def _load_rgb(path: str) -> np.ndarray:
    """Decode the optional reference image (`--image-path`; load_image, image_quality.rs:225-230).  File decoding is
    host work outside the GPU path: PIL when the environment has it."""
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)


def _save_png(path: str, img: np.ndarray) -> bool:
    try:
        from PIL import Image
    except ImportError:
        return False
    Image.fromarray(img).save(path)
    return True


def convert(input_model: CameraModel, name: str, cls, init, points_3d, points_2d, reference_image=None, output_dir=None) -> ConversionMetrics:
    """One `convert_to_*` of the reference (camera_converter.rs:355-488 and clones): target initialised with
    the input intrinsics / resolution, initial error, linear estimate, LM with the converter's bounds and
    tolerances ("Linear Only" when the solver reports an error), final error, validation."""
    t0 = time.perf_counter()
    model = cls(input_model.get_intrinsics(), input_model.get_resolution(), init, ctx=input_model.ctx)
    initial = compute_reprojection_error(model, points_3d, points_2d)
    model.linear_estimation(points_3d, points_2d)
    status, iterations = "Converged", 0
    try:
        # the reference reports "Converged" for every Ok(..) of the solver, whatever its stop reason (:417-446)
        iterations = OptimizationCost(model, points_3d, points_2d).optimize().iterations
    except AcmError:
        status = "Linear Only"
    elapsed = (time.perf_counter() - t0) * 1e3
    final = compute_reprojection_error(model, points_3d, points_2d)
    try:
        val = validate_conversion_accuracy(model, input_model)
    except AcmError:
        val = ValidationResults([math.nan] * 5, math.nan, math.nan, "NEEDS IMPROVEMENT")
    # compute_image_quality_metrics (camera_converter.rs:465-473; an Err is logged and becomes None)
    quality = None
    try:
        want = output_dir is not None
        res = compute_image_quality_metrics(input_model, model, points_3d, reference_image, return_image=want)
        quality, img = res if want else (res, None)
        if img is not None:  # save_model_projection_image (image_quality.rs:521-536): output/<model>_projection.png
            _save_png(os.path.join(output_dir, IMAGE_LABELS[cls].lower().replace(" ", "_") + "_projection.png"), img)
    except AcmError:
        pass
    return ConversionMetrics(model, name, final, initial, elapsed, status, val, iterations, quality)


def convert_all(input_model: CameraModel, num_points: int, shard=None, log=print, reference_image=None, output_dir=None, input_label=None):
    uv, xyz = sample_points(input_model, num_points, device=True, shard=shard)
    kept = len(uv)
    # camera_converter.rs:203, :213-218: the correspondences as CSV / Rust literals and the input-model projection image
    # (single-GPU runs of converter size only: these are host files of every point)
    if output_dir is not None and shard is None and kept <= 200_000:
        p2, p3 = uv.numpy(), xyz.numpy()
        export_point_correspondences(p3, p2, "point_correspondences_apex", output_dir)
        res = input_model.get_resolution()
        img = model_projection_visualization(p2, reference_image, (res.width, res.height), input_model.ctx)
        _save_png(os.path.join(output_dir, f"{(input_label or input_model.get_model_name()).lower()}_projection.png"), img)   # image_quality.rs:607
    metrics = []
    for name, cls, init in TARGETS:
        if cls is type(input_model):
            continue
        try:
            metrics.append(convert(input_model, name, cls, init, xyz, uv, reference_image, output_dir))
        except AcmError as e:  # the reference skips a target whose conversion returns Err (camera_converter.rs:232-241)
            log(f"  {name}: skipped ({e})")
    return kept, metrics, (uv, xyz)


def _fmt_params(m: CameraModel) -> str:
    i = m.get_intrinsics()
    dist = ", ".join(f"{n}={v:.6g}" for n, v in zip(m.DISTORTION_NAMES, m.get_distortion()))
    return f"fx={i.fx:.4f} fy={i.fy:.4f} cx={i.cx:.4f} cy={i.cy:.4f}" + (f" | {dist}" if dist else "")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="camera_converter", description="Camera model conversion tool (GPU path of apex-camera-models)")
    ap.add_argument("-i", "--input-model", required=True, help="kb, ds, radtan, ucm, eucm, fov, pinhole")
    ap.add_argument("-p", "--input-path", required=True, help="input model YAML")
    ap.add_argument("-n", "--num-points", type=int, default=500)
    ap.add_argument("-m", "--image-path", default=None)
    ap.add_argument("-o", "--output-dir", default="output", help="converted models are written here as <model>.yaml")
    args = ap.parse_args(argv)

    rank, local, world = D.env_rank_world()
    ctx = default_context() if world == 1 else Context(local)
    shard = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        D.attach_communicator(ctx)
        D.attach_peers(ctx)
        shard = (rank, world)
    say = print if rank == 0 else (lambda *a, **k: None)

    say("CAMERA MODEL CONVERTER - B200 path of apex-camera-models")
    say("=========================================================")
    say(f"Input model: {args.input_model.lower()} -> converting to all supported target models")
    say(f"Input file: {args.input_path}")
    say(f"Sample points: {args.num_points}" + (f" (sharded over {world} GPUs)" if world > 1 else ""))
    reference_image = None
    if args.image_path:
        reference_image = _load_rgb(args.image_path)
        say(f"Input image: {args.image_path} ({reference_image.shape[1]}x{reference_image.shape[0]})")
    input_model = load_input_model(args.input_model, args.input_path, ctx)
    say(f"\nInput {input_model.get_model_name()} parameters: {_fmt_params(input_model)}")
    res = input_model.get_resolution()
    say(f"Resolution: {res.width}x{res.height}")

    t0 = time.perf_counter()
    img_dir = args.output_dir if (rank == 0 and args.output_dir) else None
    if img_dir:
        os.makedirs(img_dir, exist_ok=True)
    # every rank takes part in the collectives of the diagnostics; only rank 0 asks for (and saves) the display images
    kept, metrics, pts = convert_all(input_model, args.num_points, shard, say, reference_image, img_dir, args.input_model)
    total_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([kept], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        kept = int(t.item())
    say(f"\nValid 3D-2D correspondences: {kept} / {args.num_points}")
    say("\nCONVERSION RESULTS")
    say(f"{'Target model':32s} {'initial px':>12s} {'final px':>12s} {'rmse':>10s} {'max':>10s} {'median':>10s} {'iters':>6s} {'ms':>9s}  status / validation")
    for m in metrics:
        f, i0, v = m.final_reprojection_error, m.initial_reprojection_error, m.validation_results
        say(f"{m.model_name:32s} {i0.mean:12.6f} {f.mean:12.6f} {f.rmse:10.6f} {f.max:10.6f} {f.median:10.6f} {m.iterations:6d} {m.optimization_time_ms:9.3f}  "
            f"{m.convergence_status} / {v.status} (avg {v.average_error:.6f} px)")
    say("")
    for m in metrics:   # reporting.rs:198-201
        if m.image_quality is not None:
            say(f"{m.model_name}: PSNR {m.image_quality.psnr:.2f} dB, SSIM {m.image_quality.ssim:.4f}")
    say("")
    for m in metrics:
        say(f"{m.model_name}: {_fmt_params(m.model)}")
    say(f"\nTotal (sampling + {len(metrics)} conversions): {total_ms:.2f} ms")
    if rank == 0 and args.output_dir:
        os.makedirs(args.output_dir, exist_ok=True)
        for m in metrics:
            m.model.save_to_yaml(os.path.join(args.output_dir, m.model.get_model_name() + ".yaml"))
        say(f"Converted models written to {args.output_dir}/")
        say(f"Report: {export_conversion_results(metrics, args.input_model, args.output_dir)}")   # reporting.rs:225-413
    for p in pts:
        p.free()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
