"""Round-2 smoke of the persistent LM kernel: small and large solves against the oracle, linearize vs oracle,
scalar host path latency.  Prints one line per check; exits non-zero on the first mismatch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apex_camera_models_b200 as acm
from apex_camera_models_b200 import _native as N
from oracle import oracle as O
O.build()
KB = [190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182]
lib = N.lib; ctx = acm.Context(0)
kb = acm.KannalaBrandtModel(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), KB[4:], ctx=ctx)
okb = O.make_model(O.KB, KB, 512, 512)
ok = True
for n in (1, 2, 450, 4097, 100_001, 2_000_000):
    xyz = O.synth_points3(0xACE50004, 0, n, float(np.cos(np.deg2rad(85.0))), False)
    uv, _ = O.project(okb, xyz)
    for name, cls, mid, kind, init in (("ds", acm.DoubleSphereModel, O.DS, 1, [0.6, 0.05]), ("ds", acm.DoubleSphereModel, O.DS, 0, [0.6, 0.05]),
                                       ("kb", acm.KannalaBrandtModel, O.KB, 0, [0.0] * 4), ("radtan", acm.RadTanModel, O.RADTAN, 0, [0.0] * 5),
                                       ("fov", acm.FovModel, O.FOV, 0, [0.9]), ("pinhole", acm.PinholeModel, O.PINHOLE, 0, [])):
        m = cls(acm.Intrinsics(*KB[:4]), acm.Resolution(512, 512), init, ctx=ctx)
        cost = acm.OptimizationCost(m, xyz, uv, residual_kind=kind)
        H, g, c, nv = cost.linearize()
        Ho, go, co, no = O.linearize(O.make_model(mid, KB[:4] + init, 512, 512), kind, xyz, uv)
        scale = np.sqrt(np.outer(np.diag(Ho), np.diag(Ho))) + 1e-300
        e = max(np.max(np.abs(H - Ho) / scale), abs(c - co) / max(co, 1e-300))
        good = e < 1e-9 and nv == no
        t0 = time.perf_counter()
        r = cost.optimize(bounds=None) if n >= 450 and name in ("ds", "kb", "fov") else None
        line = f"n={n:8d} {name:8s} kind={kind} lin_err={e:.2e} nv={nv}/{no}"
        if r is not None:
            oo, ores = O.lm_solve(O.make_model(mid, KB[:4] + init, 512, 512), kind, xyz, uv, None, None, nthreads=8)
            pe = float(np.max(np.abs(r.parameters - oo) / np.maximum(np.abs(oo), 1e-12)))
            same = (r.status, r.iterations, r.passes) == (ores.status, ores.iterations, ores.passes)
            good = good and same and pe < 1e-8
            line += f" | LM st={r.status} it={r.iterations} passes={r.passes} same={same} perr={pe:.1e} wall={r.elapsed_ms:.3f}ms dev={r.device_ms:.3f}ms ({1e3*r.device_ms/max(r.passes,1):.1f} us/pass)"
        print(("OK  " if good else "BAD ") + line, flush=True)
        ok = ok and good
        cost.free()
p = np.array([[0.1, 0.2, 1.0]])
for _ in range(100):
    kb.project(p[0])
t0 = time.perf_counter()
for _ in range(2000):
    kb.project(p[0])
print(f"scalar project through the Python trait mirror: {(time.perf_counter()-t0)/2000*1e6:.1f} us/call")
sys.exit(0 if ok else 1)
