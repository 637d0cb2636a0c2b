import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
ctx = acm.Context(0)
kb = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), KB[4:], ctx=ctx)
uv, xyz = acm.sample_points(kb, 10_000_000, device=True)
ds = acm.DoubleSphereModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [0.5, 0.1], ctx=ctx)
print(acm.compute_reprojection_error(ds, xyz, uv))
fov = acm.FovModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), [1.0], ctx=ctx)
fov.linear_estimation(xyz, uv); print(fov.params())
