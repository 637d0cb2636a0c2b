"""Times acm_undistort_rgb8 on 4096x4096 KB fisheye frames resident in HBM (BASELINE config 5 per-GPU share)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from apex_camera_models_b200 import _native as N
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
lib = N.lib; ctx = acm.Context(0)
W = H = 4096
kb8 = acm.KannalaBrandtModel(acm.Intrinsics(*(v * 8 for v in KB[:4])), acm.Resolution(W, H), KB[4:], ctx=ctx)
cam = kb8.camera_block(); fb = W * H * 3
frames = [int(a) for a in sys.argv[1:]] or [8, 32]
for F in frames:
    d_in = ctx.device_alloc(fb * F); d_out = ctx.device_alloc(fb * F)
    ctx.check(lib.acm_synth_bytes(ctx.handle, 0xACE50005, 0, C.c_void_p(d_in), fb * F))
    for interp in (1, 0):
        f = lambda: ctx.check(lib.acm_undistort_rgb8(ctx.handle, C.byref(cam), None, C.c_void_p(d_in), C.c_void_p(d_out), F, interp))
        for _ in range(3): f()
        ctx.sync(); ctx.timer_start()
        for _ in range(5): f()
        ms = ctx.timer_stop() / 5
        print(f"undistort 4096^2 KB interp={interp} frames={F}: {ms:.3f} ms  {F/ms*1e3:.0f} frames/s  {2*fb*F/ms/1e6:.0f} GB/s  ({ms/F*1e3:.1f} us/frame)")
    ctx.device_free(d_in); ctx.device_free(d_out)
