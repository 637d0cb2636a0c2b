"""Text report of a conversion run in the reference's format (reference src/util/reporting.rs:225-413,
`export_conversion_results`): `<output_dir>/camera_conversion_results_<input>.txt`.  Host-side formatting
only; the numbers come from the GPU path.  Rust's float formatting (`{}`, `{:?}`, `{:.N}`) is reproduced so
that the file can be diffed against the reference's."""
from __future__ import annotations

import math
import os


def _digits_exp(v: float):
    """Shortest round-trip decimal digits and decimal exponent: v = 0.d1d2... x 10^exp."""
    r = repr(abs(v))
    mant, _, e = r.partition("e")
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    exp = len(ip.lstrip("0")) if ip.strip("0") else -(len(fp) - len(fp.lstrip("0")))
    if not digits:
        return "0", 1
    exp += int(e) if e else 0
    return digits.rstrip("0") or "0", exp


def rust_display_f64(v: float) -> str:
    """`format!("{}", v)`: shortest round-trip digits, never an exponent, no trailing `.0`."""
    if v != v:
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    sign = "-" if math.copysign(1.0, v) < 0 else ""
    if v == 0:
        return sign + "0"
    d, e = _digits_exp(v)
    if e <= 0:
        s = "0." + "0" * (-e) + d
    elif e >= len(d):
        s = d + "0" * (e - len(d))
    else:
        s = d[:e] + "." + d[e:]
    return sign + s


def rust_debug_f64(v: float) -> str:
    """`format!("{:?}", v)`: like Display with a forced `.0`; scientific below 1e-4 and from 1e16 (core::fmt::float,
    `already_rounded_value_should_use_exponential`)."""
    if v != v or math.isinf(v):
        return rust_display_f64(v)
    sign = "-" if math.copysign(1.0, v) < 0 else ""
    if v == 0:
        return sign + "0.0"
    a = abs(v)
    if 1e-4 <= a < 1e16:
        s = rust_display_f64(a)
        return sign + (s if "." in s else s + ".0")
    d, e = _digits_exp(a)
    mant = d[0] + ("." + d[1:] if len(d) > 1 else "")
    return f"{sign}{mant}e{e - 1}"


_DEBUG_NAMES = {"double_sphere": ("DoubleSphere", "DoubleSphere"), "eucm": ("Eucm", "EUCM"), "fov": ("Fov", "FOV"),
                "kannala_brandt": ("KannalaBrandt", "KannalaBrandt"), "pinhole": ("Pinhole", None), "rad_tan": ("RadTan", "RadTan"),
                "ucm": ("Ucm", "UCM")}
_SCALAR_NAMES = {"double_sphere": ("alpha", "xi"), "eucm": ("alpha", "beta"), "ucm": ("alpha",), "fov": ("w",)}


def model_debug(model) -> str:
    """`{:?}` of `CameraModelEnum` (mod.rs:37-46) with the models' hand-written Debug impls
    (double_sphere.rs:294-307, rad_tan.rs:238-250, kannala_brandt.rs:276-288, ucm.rs:262-273, eucm.rs:292-304,
    fov.rs:255-262; Pinhole derives Debug)."""
    name = model.get_model_name()
    variant, label = _DEBUG_NAMES[name]
    i = model.get_intrinsics()
    dist = [float(v) for v in model.get_distortion()]
    if name == "pinhole":
        r = model.get_resolution()
        inner = (f"PinholeModel {{ intrinsics: Intrinsics {{ fx: {rust_debug_f64(i.fx)}, fy: {rust_debug_f64(i.fy)}, cx: {rust_debug_f64(i.cx)}, "
                 f"cy: {rust_debug_f64(i.cy)} }}, resolution: Resolution {{ width: {r.width}, height: {r.height} }} }}")
        return f"{variant}({inner})"
    head = f"{label} [fx: {rust_display_f64(i.fx)} fy: {rust_display_f64(i.fy)} cx: {rust_display_f64(i.cx)} cy: {rust_display_f64(i.cy)}"
    if name in _SCALAR_NAMES:
        tail = "".join(f" {n}: {rust_display_f64(v)}" for n, v in zip(_SCALAR_NAMES[name], dist))
    else:
        tail = " distortions: [" + ", ".join(rust_debug_f64(v) for v in dist) + "]"
    return f"{variant}({head}{tail}])"


def format_conversion_report(metrics, input_model_type: str) -> str:
    w = []
    w.append("FISHEYE CAMERA MODEL CONVERSION ANALYSIS REPORT - RUST IMPLEMENTATION")
    w.append("=====================================================================")
    w.append("")
    w.append(f"INPUT MODEL TYPE: {input_model_type.upper()}")
    w.append("OPTIMIZATION FRAMEWORK: tiny-solver")   # the reference's own (stale) banner, reporting.rs:251
    w.append("ALGORITHM: Levenberg-Marquardt")
    w.append("")
    if not metrics:
        w.append("❌ No conversions performed (input model type not supported for conversion or no target models available)")
        return "\n".join(w) + "\n"
    w.append("CONVERSION RESULTS TABLE")
    w.append("========================")
    w.append(f"{'Target Model':<32} | {'Final Error':>15} | {'Improvement':>15} | {'Time (ms)':>13} | {'Convergence':>15}")
    w.append(f"{'':<32} | {'(pixels)':>15} | {'(pixels)':>15} | {'':>13} | {'Status':>15}")
    w.append(f"{'':-<32}-+-{'':-<15}-+-{'':-<15}-+-{'':-<13}-+-{'':-<15}")
    for m in metrics:
        imp = m.initial_reprojection_error.mean - m.final_reprojection_error.mean
        w.append(f"{m.model_name:<32} | {m.final_reprojection_error.mean:>13.6f}   | {imp:>13.6f}   | {m.optimization_time_ms:>11.2f}   | {m.convergence_status:<15}")
    w.append("")
    w.append("PERFORMANCE ANALYSIS")
    w.append("====================")
    best = min(metrics, key=lambda m: m.final_reprojection_error.mean)
    fastest = min(metrics, key=lambda m: m.optimization_time_ms)
    w.append(f"🏆 Best Accuracy: {best.model_name} ({best.final_reprojection_error.mean:.6f} pixels)")
    w.append(f"⚡ Fastest Conversion: {fastest.model_name} ({fastest.optimization_time_ms:.2f} ms)")
    avg_e = sum(m.final_reprojection_error.mean for m in metrics) / len(metrics)
    avg_t = sum(m.optimization_time_ms for m in metrics) / len(metrics)
    w.append(f"📊 Average Reprojection Error: {avg_e:.6f} pixels")
    w.append(f"📊 Average Optimization Time: {avg_t:.2f} ms")
    w.append("")
    w.append("DETAILED MODEL RESULTS")
    w.append("======================")
    for m in metrics:
        e, v = m.final_reprojection_error, m.validation_results
        w.append(f"\n{m.model_name.upper()} MODEL:")
        w.append("-" * (len(m.model_name) + 7))
        w.append(f"Final Parameters: {model_debug(m.model)}")
        w.append(f"Optimization Time: {m.optimization_time_ms:.2f} ms")
        w.append(f"Convergence Status: {m.convergence_status}")
        w.append("\nReprojection Error Statistics:")
        for label, val in (("Mean", e.mean), ("RMSE", e.rmse), ("Min", e.min), ("Max", e.max), ("Std Dev", e.stddev), ("Median", e.median)):
            w.append(f"  {label}: {val:.8f} px")
        w.append("\nConversion Accuracy:")
        w.append(f"  Average Error: {_fixed(v.average_error, 4)} px")
        w.append(f"  Max Error: {_fixed(v.max_error, 4)} px")
        w.append(f"  Status: {v.status}")
        q = getattr(m, "image_quality", None)
        if q is not None:
            w.append("\nImage Quality Assessment:")
            w.append(f"  PSNR: {_fixed(q.psnr, 2)} dB")
            w.append(f"  SSIM: {_fixed(q.ssim, 4)}")
    return "\n".join(w) + "\n"


def _fixed(v: float, prec: int) -> str:
    """`{:.N}` of Rust: NaN / inf spelled `NaN` / `inf`."""
    if v != v:
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    return f"{v:.{prec}f}"


def export_conversion_results(metrics, input_model_type: str, output_dir: str = "output") -> str:
    """reporting.rs:225-413; returns the path written."""
    os.makedirs(output_dir, exist_ok=True)   # ensure_output_dir, util/mod.rs:29-37
    path = os.path.join(output_dir, f"camera_conversion_results_{input_model_type.lower()}.txt")
    with open(path, "w", encoding="utf-8") as f:
        f.write(format_conversion_report(metrics, input_model_type))
    return path


def export_point_correspondences(points_3d, points_2d, filename_prefix: str, output_dir: str = "output"):
    """`util::export_point_correspondences` (reference src/util/point_sampling.rs:153-237): `<prefix>.csv` and
    `<prefix>_rust.txt` with `{:.15}` coordinates.  Host-side file output of the converter (camera_converter.rs:203).
    Returns the two paths."""
    import numpy as np
    from .errors import UtilError
    p3 = np.asarray(points_3d, dtype=np.float64).reshape(-1, 3)
    p2 = np.asarray(points_2d, dtype=np.float64).reshape(-1, 2)
    if len(p3) != len(p2):
        raise UtilError("Invalid parameters: 3D and 2D point counts must match")
    os.makedirs(output_dir, exist_ok=True)
    n = len(p3)
    csv_path = os.path.join(output_dir, f"{filename_prefix}.csv")
    with open(csv_path, "w", encoding="utf-8") as f:
        f.write("# 3D-2D Point Correspondences from Rust Implementation\n# Format: x3d,y3d,z3d,x2d,y2d\n")
        f.write(f"# Total points: {n}\n")
        for a, b in zip(p3, p2):
            f.write(f"{a[0]:.15f},{a[1]:.15f},{a[2]:.15f},{b[0]:.15f},{b[1]:.15f}\n")
    rust_path = os.path.join(output_dir, f"{filename_prefix}_rust.txt")
    with open(rust_path, "w", encoding="utf-8") as f:
        f.write("// 3D-2D Point Correspondences for Rust Import\n// Generated from Rust fisheye-tools\n")
        f.write("let points_3d = Matrix3xX::from_columns(&[\n")
        f.write(",\n".join(f"    Vector3::new({a[0]:.15f}, {a[1]:.15f}, {a[2]:.15f})" for a in p3) + ("\n" if n else ""))
        f.write("]);\n\nlet points_2d = Matrix2xX::from_columns(&[\n")
        f.write(",\n".join(f"    Vector2::new({b[0]:.15f}, {b[1]:.15f})" for b in p2) + ("\n" if n else ""))
        f.write("]);\n")
    return csv_path, rust_path
