"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle."""
import json
import os

import numpy as np


def run():
    import apex_camera_models_b200 as acm
    from oracle import oracle as O  # checker only
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cams = json.load(open(os.path.join(root, "tests", "golden", "cameras.json")))
    ctx = acm.Context(0)
    c = cams["double_sphere"]
    ds = acm.DoubleSphereModel(acm.Intrinsics(*c["params"][:4]), acm.Resolution(c["width"], c["height"]), c["params"][4:], ctx=ctx)
    om = O.make_model(O.DS, c["params"], c["width"], c["height"])
    n = 65537
    xyz = O.synth_points3(0xACE50001, 0, n, np.cos(np.deg2rad(100.0)), True)
    uv, st = ds.project_batch(xyz)
    uvo, sto = O.project(om, xyz)
    assert np.array_equal(st, sto), "status masks differ"
    ok = sto == 0
    assert np.allclose(uv[ok], uvo[ok], rtol=1e-9, atol=0), "projections differ"
    ray, st2 = ds.unproject_batch(uv[ok])
    rayo, st2o = O.unproject(om, uvo[ok])
    assert np.array_equal(st2, st2o) and np.allclose(ray[st2 == 0], rayo[st2o == 0], rtol=1e-9, atol=1e-15)
    # fused project + Jacobian + J^T J / J^T r and a short LM
    obs = np.where(np.isnan(uvo), 0.0, uvo) + 0.125
    cost = acm.DoubleSphereOptimizationCost(ds, xyz, obs, residual_kind=0)
    H, g, cst, nv = cost.linearize()
    Ho, go, co, nvo = O.linearize(om, 0, xyz, obs)
    assert nv == nvo and np.allclose(H, Ho, rtol=1e-9) and np.allclose(g, go, rtol=1e-8, atol=1e-8 * np.abs(go).max()) and np.isclose(cst, co, rtol=1e-9)
    res = cost.optimize(bounds=None)
    assert res.converged and res.final_cost < res.initial_cost
    print(f"smoke ok: n={n} valid={int(ok.sum())} lm_iters={res.iterations} cost {res.initial_cost:.3e}->{res.final_cost:.3e} launches={ctx.kernel_launches()}")
