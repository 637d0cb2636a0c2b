"""Per-iteration cost of the device-resident LM loop at several problem sizes / poll intervals."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from apex_camera_models_b200 import _native as N
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
lib = N.lib; ctx = acm.Context(0)
kb = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), KB[4:], ctx=ctx)
for n in (450, 100_000, 1_250_000, 10_000_000):
    X = acm.Points(ctx, 3, n); 
    ctx.check(lib.acm_synth_points3(ctx.handle, 0xACE50004, 0, float(np.cos(np.deg2rad(85.0))), 0, X.handle))
    U, st = kb.project_batch(X); ctx.device_free(st)
    for ce in (1, 4, 16):
        ds = acm.DoubleSphereModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.6467, 0.0], ctx=ctx)
        cost = acm.DoubleSphereOptimizationCost(ds, X, U)
        cfg = acm.LevenbergMarquardtConfig(max_iterations=64, cost_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0, check_every=ce)
        start = ds.params().copy()
        cost.optimize(config=cfg); ds.set_params(start)
        r = cost.optimize(config=cfg)
        ds.set_params(start)
        r2 = cost.optimize(config=acm.LevenbergMarquardtConfig(check_every=ce))
        print(f"n={n:9d} check_every={ce:2d}: fixed {r.iterations} it / {r.passes} passes {r.elapsed_ms:.3f} ms = {1e3*r.elapsed_ms/r.passes:.1f} us/pass | converter cfg: {r2.iterations} it {r2.elapsed_ms:.3f} ms status {r2.status}")
    X.free(); U.free()
