"""Early GPU sanity run (development aid): parity of project/unproject/linearize/LM against the
oracle on small sizes + a first timing of the 100M-point kernels.  Not part of the test suite."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
L = C.CDLL(os.path.join(ROOT, "apex_camera_models_b200/lib/libacm.so"))
vp = C.c_void_p
class Camera(C.Structure):
    _fields_ = [("model", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32), ("n_params", C.c_int32), ("params", C.c_double * 9)]
class NE(C.Structure):
    _fields_ = [("n_params", C.c_int32), ("H", C.c_double * 81), ("g", C.c_double * 9), ("cost", C.c_double), ("n_valid", C.c_uint64)]
class LMC(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("cost_tolerance", C.c_double), ("parameter_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("lambda0", C.c_double), ("invalid_penalty", C.c_double), ("check_every", C.c_int32)]
class LMR(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32), ("passes", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double), ("n_valid", C.c_uint64), ("elapsed_ms", C.c_double)]
L.acm_last_error.restype = C.c_char_p; L.acm_last_error.argtypes = [vp]
L.acm_points_component.restype = vp
ctx = vp()
rc = L.acm_ctx_create(0, None, C.byref(ctx)); assert rc == 0, L.acm_last_error(None)
def ck(rc):
    assert rc == 0, (rc, L.acm_last_error(ctx))
def cam(c):
    k = Camera(); k.model = c["model_id"]; k.width = c["width"]; k.height = c["height"]; k.n_params = len(c["params"])
    for i, v in enumerate(c["params"]): k.params[i] = v
    return k
def pts(dim, n, dtype=0):
    p = vp(); ck(L.acm_points_create(ctx, dim, C.c_size_t(n), dtype, C.byref(p))); return p
def up(p, a):
    a = np.ascontiguousarray(a, dtype=np.float64); ck(L.acm_points_upload_aos_f64(ctx, p, a.ctypes.data_as(vp), C.c_size_t(a.shape[0]))); ck(L.acm_ctx_sync(ctx))
def down(p, n, dim):
    a = np.empty((n, dim)); ck(L.acm_points_download_aos_f64(ctx, p, a.ctypes.data_as(vp), C.c_size_t(n))); return a
def dalloc(nbytes):
    p = vp(); ck(L.acm_device_alloc(ctx, C.c_size_t(nbytes), C.byref(p))); return p
cams = json.load(open(os.path.join(ROOT, "tests/golden/cameras.json")))
n = 100003
for name in ["pinhole", "rad_tan", "kannala_brandt", "ucm", "eucm", "double_sphere", "fov"]:
    c = cams[name]; om = O.make_model(c["model_id"], c["params"], c["width"], c["height"]); k = cam(c)
    cosmax = np.cos(np.deg2rad(40.0 if name in ("pinhole", "rad_tan") else 100.0))
    xyz = O.synth_points3(0xACE50002, 0, n, cosmax, True)
    X = pts(3, n); UV = pts(2, n); S = dalloc(n)
    # device generator must equal the oracle's bit for bit
    ck(L.acm_synth_points3(ctx, C.c_uint64(0xACE50002), C.c_size_t(0), C.c_double(cosmax), 1, X))
    xyz_dev = down(X, n, 3)
    synth_equal = np.array_equal(xyz_dev, xyz)
    up(X, xyz)
    ck(L.acm_project(ctx, C.byref(k), X, UV, S))
    uv = down(UV, n, 2); st = np.empty(n, np.uint8); ck(L.acm_memcpy_d2h(ctx, st.ctypes.data_as(vp), S, C.c_size_t(n))); ck(L.acm_ctx_sync(ctx))
    uvo, sto = O.project(om, xyz)
    ok = sto == 0
    rel = np.abs(uv[ok] - uvo[ok]).max() if ok.any() else 0
    print(f"{name:16s} project: synth_equal={synth_equal} status_equal={np.array_equal(st, sto)} valid={ok.mean():.3f} max_abs_px={rel:.3e} nan_match={np.array_equal(np.isnan(uv), np.isnan(uvo))}")
    px = O.synth_pixels(0xACE50002, 0, n, c["width"], c["height"])
    up(UV, px); ck(L.acm_unproject(ctx, C.byref(k), UV, X, S))
    ray = down(X, n, 3); ck(L.acm_memcpy_d2h(ctx, st.ctypes.data_as(vp), S, C.c_size_t(n))); ck(L.acm_ctx_sync(ctx))
    rayo, sto = O.unproject(om, px); ok = sto == 0
    print(f"{'':16s} unproject: status_equal={np.array_equal(st, sto)} valid={ok.mean():.3f} max_abs={np.abs(ray[ok]-rayo[ok]).max() if ok.any() else 0:.3e} codes={np.unique(sto).tolist()}")
    # linearize
    uvobs = np.where(np.isnan(uvo), 0.0, uvo) + 0.3
    uvobs, _ = O.project(om, xyz); uvobs = np.where(np.isnan(uvobs), 1.0, uvobs) + 0.25
    up(X, xyz); up(UV, uvobs)
    kinds = [0, 1] if name in ("ucm", "eucm", "double_sphere") else [0]
    for kind in kinds:
        ne = NE(); ck(L.acm_linearize(ctx, C.byref(k), kind, X, UV, C.byref(ne)))
        P = len(c["params"]); H = np.array(ne.H[:P*P]).reshape(P, P); g = np.array(ne.g[:P])
        Ho, go, co, nvo = O.linearize(om, kind, xyz, uvobs, nthreads=8)
        print(f"{'':16s} linearize kind={kind}: nvalid {ne.n_valid}=={nvo} H_rel={np.abs(H-Ho).max()/np.abs(Ho).max():.2e} Hij_rel={(np.abs(H-Ho)/(np.abs(Ho)+1e-300)).max():.2e} g_rel={np.abs(g-go).max()/np.abs(go).max():.2e} cost_rel={abs(ne.cost-co)/co:.2e}")
    L.acm_points_destroy(ctx, X); L.acm_points_destroy(ctx, UV); L.acm_device_free(ctx, S)

# LM: the converter's config-1 problem
c = cams["kannala_brandt"]; kb = O.make_model(2, c["params"], 512, 512)
uv, xyz = O.sample_points(kb, 500)
X = pts(3, len(uv)); UV = pts(2, len(uv)); up(X, xyz); up(UV, uv)
intr = c["params"][:4]
for kind in (1, 0):
    ds = O.make_model(O.DS, intr + [0.5, 0.1], 512, 512); O.linear_estimation(ds, xyz, uv)
    k = Camera(); k.model = 5; k.width = 512; k.height = 512; k.n_params = 6
    for i, v in enumerate(ds.params()): k.params[i] = v
    lo = (C.c_double * 6)(1, 1, 0, 0, 1e-6, -5); hi = (C.c_double * 6)(2000, 2000, 2000, 2000, 1, 5)
    for tight in (False, True):
        cfg = LMC(); L.acm_lm_default_config(C.byref(cfg)); ocfg = O.lm_default_config()
        if tight:
            cfg.max_iterations = 500; cfg.cost_tolerance = 0; cfg.parameter_tolerance = 1e-15; cfg.gradient_tolerance = 0
            ocfg.max_iterations = 500; ocfg.cost_tolerance = 0; ocfg.parameter_tolerance = 1e-15; ocfg.gradient_tolerance = 0
        out = (C.c_double * 9)(); res = LMR()
        ck(L.acm_lm_solve(ctx, C.byref(k), kind, X, UV, lo, hi, C.byref(cfg), out, C.byref(res)))
        oo, ores = O.lm_solve(ds, kind, xyz, uv, list(lo), list(hi), ocfg)
        o = np.array(out[:6])
        print(f"LM DS kind={kind} tight={tight}: gpu status={res.status} it={res.iterations} passes={res.passes} cost={res.final_cost:.12e} ms={res.elapsed_ms:.2f} | oracle status={ores.status} it={ores.iterations} cost={ores.final_cost:.12e} | rel diff {np.abs(o-oo)/np.abs(oo)}")

# timing at 100M points
N = 100_000_000
X = pts(3, N); UV = pts(2, N); S = dalloc(N)
ck(L.acm_synth_points3(ctx, C.c_uint64(0xACE50003), C.c_size_t(0), C.c_double(np.cos(np.deg2rad(85.0))), 0, X))
kbk = cam(cams["kannala_brandt"])
ck(L.acm_project(ctx, C.byref(kbk), X, UV, S)); ck(L.acm_ctx_sync(ctx))
ms = C.c_float()
def timeit(fn, reps=10):
    for _ in range(3): fn()
    ck(L.acm_ctx_sync(ctx)); ck(L.acm_timer_start(ctx))
    for _ in range(reps): fn()
    ck(L.acm_timer_stop(ctx, C.byref(ms))); return ms.value / reps
for name in ["double_sphere", "eucm", "ucm", "kannala_brandt", "fov", "rad_tan", "pinhole"]:
    c = dict(cams[name]); c["width"] = 512; c["height"] = 512
    if name not in ("kannala_brandt",): c["params"] = cams["kannala_brandt"]["params"][:4] + {"double_sphere": [0.6, 0.1], "eucm": [0.6, 1.0], "ucm": [0.6], "fov": [0.9], "rad_tan": [0.0]*5, "pinhole": []}[name]
    k = cam(c)
    for kind in ([0, 1] if name in ("ucm", "eucm", "double_sphere") else [0]):
        t = timeit(lambda: ck(L.acm_linearize_async(ctx, C.byref(k), kind, X, UV)))
        print(f"linearize {name:16s} kind={kind}: {t:.3f} ms  {N/t/1e6:.1f} Gpts/s  {N*40/t/1e6:.0f} GB/s")
    XO = None
for name in ["double_sphere", "kannala_brandt", "pinhole", "rad_tan", "ucm", "eucm", "fov"]:
    c = dict(cams[name]); k = cam(c)
    UV2 = pts(2, N)
    t = timeit(lambda: ck(L.acm_project(ctx, C.byref(k), X, UV2, S)))
    print(f"project   {name:16s}: {t:.3f} ms  {N/t/1e6:.1f} Gpts/s  {N*41/t/1e6:.0f} GB/s")
    X2 = pts(3, N)
    ck(L.acm_synth_pixels(ctx, C.c_uint64(7), C.c_size_t(0), C.c_double(c["width"]), C.c_double(c["height"]), UV2))
    t = timeit(lambda: ck(L.acm_unproject(ctx, C.byref(k), UV2, X2, S)))
    print(f"unproject {name:16s}: {t:.3f} ms  {N/t/1e6:.1f} Gpts/s  {N*41/t/1e6:.0f} GB/s")
    L.acm_points_destroy(ctx, UV2); L.acm_points_destroy(ctx, X2)
print("launches", L.acm_ctx_kernel_launches(ctx))
