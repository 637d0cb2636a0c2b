"""One launch of the fused pass per model (after two warm-up launches) on 100 M points, for an `ncu --set full` capture:
    MODELS=2,6,1,5 python scripts/ncu_lin.py
plus (UNPROJECT=1) one RadTan / KB unproject and (UNDISTORT=1) one bilinear undistort of 8 frames at 4096^2."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from apex_camera_models_b200 import _native as N
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
lib = N.lib; ctx = acm.Context(0)
n = int(os.environ.get("N", "100000000"))
X = acm.Points(ctx, 3, n)
ctx.check(lib.acm_synth_points3(ctx.handle, 0xACE50003, 0, float(np.cos(np.deg2rad(85.0))), 0, X.handle))
kb = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), KB[4:], ctx=ctx)
UV, st = kb.project_batch(X)
dist = {0: [], 1: [0.01, 0.001, 0.0, 0.0, 0.0], 2: KB[4:], 3: [0.6], 4: [0.6, 1.0], 5: [0.6, 0.1], 6: [0.9]}
for mid in [int(v) for v in os.environ.get("MODELS", "2,6,1,5").split(",")]:
    m = acm.MODEL_CLASSES[mid](acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), dist[mid], ctx=ctx)
    cam = m.camera_block()
    for _ in range(3):
        ctx.check(lib.acm_linearize_async(ctx.handle, C.byref(cam), 0, X.handle, UV.handle))
    ctx.sync()
if os.environ.get("UNPROJECT"):
    UV2 = acm.Points(ctx, 2, n); X2 = acm.Points(ctx, 3, n)
    ctx.check(lib.acm_synth_pixels(ctx.handle, 7, 0, 752.0, 480.0, UV2.handle))
    rt = acm.RadTanModel(acm.Intrinsics(461.629, 460.152, 362.68, 246.049), acm.Resolution(752, 480), [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0], ctx=ctx)
    cam = rt.camera_block()
    for _ in range(3):
        ctx.check(lib.acm_unproject(ctx.handle, C.byref(cam), UV2.handle, X2.handle, C.c_void_p(st)))
    ctx.sync()
if os.environ.get("UNDISTORT"):
    W = H = 4096; F = 8
    kb8 = acm.KannalaBrandtModel(acm.Intrinsics(*(v * 8 for v in KB[:4])), acm.Resolution(W, H), KB[4:], ctx=ctx)
    cam8 = kb8.camera_block()
    fb = W * H * 3
    d_in = ctx.device_alloc(fb * F); d_out = ctx.device_alloc(fb * F)
    ctx.check(lib.acm_synth_bytes(ctx.handle, 0xACE50005, 0, C.c_void_p(d_in), fb * F))
    for _ in range(2):
        ctx.check(lib.acm_undistort_rgb8(ctx.handle, C.byref(cam8), None, C.c_void_p(d_in), C.c_void_p(d_out), F, 1))
    ctx.sync()
print("done", ctx.kernel_launches())
