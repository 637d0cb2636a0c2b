"""Launches each kernel of the path once on 100 M points (for an `ncu --set full` capture)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from apex_camera_models_b200 import _native as N
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
lib = N.lib; ctx = acm.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
X = acm.Points(ctx, 3, n)
ctx.check(lib.acm_synth_points3(ctx.handle, 0xACE50003, 0, float(np.cos(np.deg2rad(85.0))), 0, X.handle))
kb = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), KB[4:], ctx=ctx)
UV, st = kb.project_batch(X)
intr = KB[:4]
dist = {0: [], 1: [0.01, 0.001, 0.0, 0.0, 0.0], 2: KB[4:], 3: [0.6], 4: [0.6, 1.0], 5: [0.6, 0.1], 6: [0.9]}
UV2 = acm.Points(ctx, 2, n); X2 = acm.Points(ctx, 3, n)
ctx.check(lib.acm_synth_pixels(ctx.handle, 7, 0, 512.0, 512.0, UV2.handle))
for mid in (5, 2, 1, 6, 4):
    m = acm.MODEL_CLASSES[mid](acm.Intrinsics(*intr), acm.Resolution(512, 512), dist[mid], ctx=ctx)
    cam = m.camera_block()
    ctx.check(lib.acm_linearize_async(ctx.handle, C.byref(cam), 0, X.handle, UV.handle))
    ctx.check(lib.acm_unproject(ctx.handle, C.byref(cam), UV2.handle, X2.handle, C.c_void_p(st)))
    if mid in (5, 2):
        UV3 = acm.Points(ctx, 2, n)
        ctx.check(lib.acm_project(ctx.handle, C.byref(cam), X.handle, UV3.handle, C.c_void_p(st)))
        ctx.sync(); UV3.free()
ctx.sync()
print("done", ctx.kernel_launches())
