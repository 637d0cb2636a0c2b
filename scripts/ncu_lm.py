"""Two KB -> Double Sphere solves of N correspondences (default 1.25 M) on one GPU, for an `ncu --set full -k regex:lin_kernel` capture
of the persistent solve kernel (the first solve is the warm-up)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from apex_camera_models_b200 import _native as N
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
lib = N.lib; ctx = acm.Context(0)
kb = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), KB[4:], ctx=ctx)
n = int(os.environ.get("N", "1250000"))
X = acm.Points(ctx, 3, n)
ctx.check(lib.acm_synth_points3(ctx.handle, 0xACE50004, 0, float(np.cos(np.deg2rad(85.0))), 0, X.handle))
U, st = kb.project_batch(X); ctx.device_free(st)
ds = acm.DoubleSphereModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.6467, 0.0], ctx=ctx)
cost = acm.DoubleSphereOptimizationCost(ds, X, U)
start = ds.params().copy()
for _ in range(2):
    ds.set_params(start)
    r = cost.optimize()
print(f"n={n}: {r.passes} passes, device {r.device_ms:.3f} ms, status {r.status}")
