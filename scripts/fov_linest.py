"""FOV grid-search linear_estimation at the converter's 10 M scale (timing / ncu launch list)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
ctx = acm.Context(0)
kb = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), KB[4:], ctx=ctx)
uv, xyz = acm.sample_points(kb, int(os.environ.get("N", "10000000")), device=True)
m = acm.FovModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [1.0], ctx=ctx)
for rep in range(3):
    t0 = time.perf_counter(); m.linear_estimation(xyz, uv); ctx.sync(); t1 = time.perf_counter()
    print(f"fov linear_estimation n={len(uv)}: {(t1 - t0) * 1e3:.2f} ms -> w={m.params()[4]}")
