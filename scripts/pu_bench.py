"""Times project / unproject / fused round trip for the models in MODELS (default all) on 100 M f64 points."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from apex_camera_models_b200 import _native as N
lib = N.lib
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
n = int(os.environ.get("N", "100000000")); reps = int(os.environ.get("REPS", "10"))
ctx = acm.Context(0)
X = acm.Points(ctx, 3, n); UV = acm.Points(ctx, 2, n); X2 = acm.Points(ctx, 3, n)
ctx.check(lib.acm_synth_points3(ctx.handle, 0xACE50003, 0, float(np.cos(np.deg2rad(85.0))), 0, X.handle))
st = ctx.device_alloc(n); st2 = ctx.device_alloc(n)
dist_init = {0: [], 1: [0.01, 0.001, 0.0, 0.0, 0.0], 2: KB[4:], 3: [0.6], 4: [0.6, 1.0], 5: [0.6, 0.1], 6: [0.9]}
names = {0: "pinhole", 1: "rad_tan", 2: "kannala_brandt", 3: "ucm", 4: "eucm", 5: "double_sphere", 6: "fov"}
def timeit(f):
    for _ in range(3): f()
    ctx.sync(); ctx.timer_start()
    for _ in range(reps): f()
    return ctx.timer_stop() / reps
for mid in [int(v) for v in os.environ.get("MODELS", "0,1,2,3,4,5,6").split(",")]:
    m = acm.MODEL_CLASSES[mid](acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), dist_init[mid], ctx=ctx)
    cam = m.camera_block()
    a = timeit(lambda: ctx.check(lib.acm_project(ctx.handle, C.byref(cam), X.handle, UV.handle, C.c_void_p(st))))
    ctx.check(lib.acm_synth_pixels(ctx.handle, 7, 0, 512.0, 512.0, UV.handle))
    b = timeit(lambda: ctx.check(lib.acm_unproject(ctx.handle, C.byref(cam), UV.handle, X2.handle, C.c_void_p(st))))
    c = timeit(lambda: ctx.check(lib.acm_project_unproject(ctx.handle, C.byref(cam), X.handle, UV.handle, X2.handle, C.c_void_p(st), C.c_void_p(st2))))
    print(f"{names[mid]:15s} project {a:.3f} ms {n*41/a/1e6:5.0f} GB/s | unproject {b:.3f} ms {n*41/b/1e6:5.0f} GB/s | round trip {c:.3f} ms {n*66/c/1e6:5.0f} GB/s", flush=True)
