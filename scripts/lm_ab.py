"""Device time of the KB -> Double Sphere solve at a few sizes (one line; used for same-box A/B of builds via ACM_LIB_PATH)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from apex_camera_models_b200 import _native as N
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
lib = N.lib; ctx = acm.Context(0)
kb = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), KB[4:], ctx=ctx)
out = []
for n in [int(v) for v in os.environ.get("SIZES", "450,1250000,10000000").split(",")]:
    X = acm.Points(ctx, 3, n)
    ctx.check(lib.acm_synth_points3(ctx.handle, 0xACE50004, 0, float(np.cos(np.deg2rad(85.0))), 0, X.handle))
    U, st = kb.project_batch(X); ctx.device_free(st)
    target = os.environ.get("TARGET", "ds")
    if target == "ds":
        ds = acm.DoubleSphereModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.6467, 0.0], ctx=ctx)
        cost = acm.DoubleSphereOptimizationCost(ds, X, U)
    elif target == "eucm":
        ds = acm.EucmModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.6, 1.0], ctx=ctx)
        cost = acm.EucmOptimizationCost(ds, X, U)
    elif target == "kb":
        ds = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.0] * 4, ctx=ctx)
        cost = acm.KannalaBrandtOptimizationCost(ds, X, U)
    elif target == "radtan":
        ds = acm.RadTanModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.0] * 5, ctx=ctx)
        cost = acm.RadTanOptimizationCost(ds, X, U)
    elif target == "fov":
        ds = acm.FovModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.9], ctx=ctx)
        cost = acm.FovOptimizationCost(ds, X, U)
    else:
        ds = acm.UcmModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.6], ctx=ctx)
        cost = acm.UcmOptimizationCost(ds, X, U)
    start = ds.params().copy()
    best = 1e9
    for _ in range(8):
        ds.set_params(start)
        r = cost.optimize()
        best = min(best, r.device_ms)
    out.append(f"{target} n={n}: {r.passes} passes {best:.4f} ms ({1e3 * best / r.passes:.2f} us/pass)")
    X.free(); U.free()
print(" | ".join(out))
