import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from oracle import oracle as O
cams = json.load(open("tests/golden/cameras.json")); c = cams["rad_tan"]
ctx = acm.Context(0)
src = acm.RadTanModel(acm.Intrinsics(*c["params"][:4]), acm.Resolution(752, 480), c["params"][4:], ctx=ctx)
uv, xyz = acm.sample_points(src, 50)
print("n", len(uv), xyz[:2], uv[:2])
est = acm.RadTanModel.new(c["params"][:4] + [0.0] * 5, ctx=ctx); est.resolution = src.get_resolution()
est.linear_estimation(xyz, uv); print("gpu", est.params())
om = O.make_model(1, c["params"][:4] + [0.0] * 5, 752, 480); print(O.linear_estimation(om, xyz, uv), om.params())
