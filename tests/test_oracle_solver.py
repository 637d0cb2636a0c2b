"""Pin the oracle's solver side (Jacobians, normal equations, LM, linear estimation, util) against
independent implementations.  The reference keeps this code in the un-vendored apex-solver crate
(PARITY UNPINNED, see oracle/acm_oracle.h), so the anchors are: OpenCV, mpmath, scipy and the
survey's restated README figures."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, oracle_model

TIGHT = dict(max_iterations=500, cost_tolerance=0.0, parameter_tolerance=1e-15, gradient_tolerance=0.0)


def _cfg(O, **kw):
    c = O.lm_default_config()
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@pytest.mark.parametrize("cam", ["kannala_brandt", "kannala_brandt_inline", "rad_tan", "pinhole"])
def test_projection_matches_opencv(O, cameras, cam):
    g = load_golden("opencv_cross.json")[cam]
    c = cameras[cam]
    m = O.make_model(c["model_id"], c["params"], 0, 0)
    for X, ref in zip(g["points"], g["uv"]):
        uv = np.empty(2)
        st = O.lib().orc_project_nobounds(C.byref(m), O._dp(np.ascontiguousarray(X, dtype=np.float64)), O._dp(uv))
        assert st == 0
        assert np.max(np.abs(uv - np.array(ref))) < 1e-10  # px


@pytest.mark.parametrize("cam", ["pinhole", "rad_tan", "kannala_brandt", "ucm", "eucm", "double_sphere", "fov"])
def test_param_jacobian_matches_mpmath(O, cameras, cam):
    rows = load_golden("mpmath_jacobians.json")[cam]
    c = cameras[cam]
    m = O.make_model(c["model_id"], c["params"], 0, 0)
    for row in rows:
        st, uv, J = O.project_jacobian1(m, row["point"])
        assert st == 0
        Jr = np.array(row["J"])
        assert np.allclose(uv, row["uv"], rtol=1e-13, atol=0)
        assert np.allclose(J, Jr, rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("cam", ["ucm", "eucm", "double_sphere"])
def test_algebraic_residual_jacobian_fd(O, cameras, cam):
    c = cameras[cam]
    p0 = np.array(c["params"])
    rng = np.random.default_rng(7)
    for _ in range(8):
        X = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(0.8, 3.0)])
        uvo = np.array([rng.uniform(0, 752), rng.uniform(0, 480)])
        m = O.make_model(c["model_id"], p0, 0, 0)
        st, r, J = O.residual_jacobian1(m, O.RES_ALGEBRAIC, X, uvo)
        assert st == 0
        for k in range(len(p0)):
            h = 1e-6 * max(1.0, abs(p0[k]))
            pp, pm = p0.copy(), p0.copy()
            pp[k] += h; pm[k] -= h
            _, rp, _ = O.residual_jacobian1(O.make_model(c["model_id"], pp, 0, 0), O.RES_ALGEBRAIC, X, uvo)
            _, rm, _ = O.residual_jacobian1(O.make_model(c["model_id"], pm, 0, 0), O.RES_ALGEBRAIC, X, uvo)
            fd = (rp - rm) / (2 * h)
            assert np.allclose(J[:, k], fd, rtol=2e-6, atol=1e-6 * (1 + np.abs(r).max()))


def _kb_correspondences(O, cameras, n=500):
    c = cameras["kannala_brandt"]
    kb = oracle_model(O, c)
    uv, xyz = O.sample_points(kb, n)
    return c["params"][:4], uv, xyz


def test_linearize_is_jtj(O, cameras):
    intr, uv, xyz = _kb_correspondences(O, cameras)
    for model_id, dist, kinds in [(O.DS, [0.6, 0.1], (0, 1)), (O.EUCM, [0.6, 1.1], (0, 1)), (O.UCM, [0.6], (0, 1)),
                                  (O.KB, [0.01, 0.0, 0.0, 0.0], (0,)), (O.FOV, [0.9], (0,)),
                                  (O.RADTAN, [0.01, 0.0, 0.001, 0.0, 0.0], (0,)), (O.PINHOLE, [], (0,))]:
        m = O.make_model(model_id, intr + dist, 512, 512)
        for kind in kinds:
            H, g, cost, nv = O.linearize(m, kind, xyz, uv)
            Js, rs = [], []
            for X, p in zip(xyz, uv):
                st, r, J = O.residual_jacobian1(m, kind, X, p)
                if st == 0:
                    Js.append(J); rs.append(r)
            J = np.concatenate(Js); r = np.concatenate(rs)
            assert nv == len(Js)
            assert np.allclose(H, J.T @ J, rtol=1e-12)
            assert np.allclose(g, J.T @ r, rtol=1e-10, atol=1e-9)
            assert np.isclose(cost, 0.5 * r @ r, rtol=1e-12)
            H4, g4, c4, n4 = O.linearize(m, kind, xyz, uv, nthreads=4)
            assert np.allclose(H4, H, rtol=1e-12) and np.allclose(g4, g, rtol=1e-9, atol=1e-9) and n4 == nv
    with pytest.raises(ValueError):
        O.linearize(O.make_model(O.KB, intr + [0, 0, 0, 0], 512, 512), O.RES_ALGEBRAIC, xyz, uv)


def test_linear_estimation_anchors(O, cameras):
    """alpha = 0.6467229596331426 for DS/UCM/EUCM (identical 2Nx1 system), FOV grid
    search -> w = 1.03; initial DS error 10.0321 px, linear-only error 0.31413 px (== README's
    EUCM "0.314 px", reference README.md:165)."""
    A = load_golden("restated_kats.json")["lm_anchors_450"]
    intr, uv, xyz = _kb_correspondences(O, cameras)
    assert len(uv) == 450
    ds = O.make_model(O.DS, intr + [0.5, 0.1], 512, 512)
    assert abs(O.reprojection_error(ds, xyz, uv).mean - A["ds_initial_mean_px"]) < 1e-4
    for mid, dist in [(O.DS, [0.5, 0.1]), (O.UCM, [0.5]), (O.EUCM, [0.5, 1.0])]:
        m = O.make_model(mid, intr + dist, 512, 512)
        assert O.linear_estimation(m, xyz, uv) == 0
        assert abs(m.params()[4] - A["linear_alpha"]) < 1e-13
        assert abs(O.reprojection_error(m, xyz, uv).mean - A["linear_only_mean_px"]) < 1e-5
    assert ds.params()[5] == 0.1  # untouched copy
    fov = O.make_model(O.FOV, intr + [1.0], 512, 512)
    assert O.linear_estimation(fov, xyz, uv) == 0 and fov.params()[4] == A["fov_linear_w"]
    kbm = O.make_model(O.KB, intr + [0.0] * 4, 512, 512)
    assert O.linear_estimation(kbm, xyz, uv) == 0
    assert np.allclose(kbm.params()[4:], cameras["kannala_brandt"]["params"][4:], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name,mid,dist,kind,lo,hi,rtol", [
    ("ds_algebraic", 5, [0.5, 0.1], 1, [1, 1, 0, 0, 1e-6, -5], [2000, 2000, 2000, 2000, 1, 5], 5e-9),
    ("ds_pixel", 5, [0.5, 0.1], 0, [1, 1, 0, 0, 1e-6, -5], [2000, 2000, 2000, 2000, 1, 5], 2e-8),
    ("ucm_algebraic", 3, [0.5], 1, [1, 1, 0, 0, 1e-6], [2000, 2000, 2000, 2000, 10], 1e-10),
    ("ucm_pixel", 3, [0.5], 0, [1, 1, 0, 0, 1e-6], [2000, 2000, 2000, 2000, 10], 1e-9),
    ("fov_pixel", 6, [1.0], 0, [1, 1, 0, 0, 1e-6], [2000, 2000, 2000, 2000, 3], 1e-9),
])
def test_lm_matches_scipy_anchors(O, cameras, name, mid, dist, kind, lo, hi, rtol):
    """Converged parameters of the converter's config-1 problem (450 correspondences, inits and
    bounds of bin/camera_converter.rs:364-400 and clones) vs scipy.least_squares run at 1e-15
    tolerances during the survey.  rtol reflects cond(J^T J) ~ 1e9 for Double Sphere."""
    A = load_golden("restated_kats.json")["lm_anchors_450"][name]
    intr, uv, xyz = _kb_correspondences(O, cameras)
    m = O.make_model(mid, intr + dist, 512, 512)
    assert O.linear_estimation(m, xyz, uv) == 0
    out, res = O.lm_solve(m, kind, xyz, uv, lo, hi, _cfg(O, **TIGHT))
    assert res.status in (0, 1, 2)
    assert np.allclose(out, A["params"], rtol=rtol, atol=0)
    fit = O.make_model(mid, out, 512, 512)
    assert abs(O.reprojection_error(fit, xyz, uv).mean - A["mean_px"]) < 5e-6  # anchors carry 5-7 digits


def test_lm_eucm_anchors(O, cameras):
    A = load_golden("restated_kats.json")["lm_anchors_450"]
    intr, uv, xyz = _kb_correspondences(O, cameras)
    for kind, key in [(1, "eucm_algebraic"), (0, "eucm_pixel")]:
        m = O.make_model(O.EUCM, intr + [0.5, 1.0], 512, 512)
        O.linear_estimation(m, xyz, uv)
        out, res = O.lm_solve(m, kind, xyz, uv, [1, 1, 0, 0, 1e-6, 1e-6], [2000, 2000, 2000, 2000, 1, 5], _cfg(O, **TIGHT))
        assert abs(out[4] - A[key]["alpha"]) < 1e-8 and abs(out[5] - A[key]["beta"]) < 1e-8


def test_lm_reference_config_reaches_readme_figure(O, cameras):
    """With the converter's own tolerances (camera_converter.rs:410-415) KB->DS ends at the
    README's 0.008 px (README.md:163) and KB->UCM at 0.145 px (README.md:164)."""
    intr, uv, xyz = _kb_correspondences(O, cameras)
    ds = O.make_model(O.DS, intr + [0.5, 0.1], 512, 512)
    O.linear_estimation(ds, xyz, uv)
    out, res = O.lm_solve(ds, O.RES_ALGEBRAIC, xyz, uv, [1, 1, 0, 0, 1e-6, -5], [2000, 2000, 2000, 2000, 1, 5])
    assert res.status == 0 and res.iterations < 100
    assert round(O.reprojection_error(O.make_model(O.DS, out, 512, 512), xyz, uv).mean, 3) == 0.008
    ucm = O.make_model(O.UCM, intr + [0.5], 512, 512)
    O.linear_estimation(ucm, xyz, uv)
    out, res = O.lm_solve(ucm, O.RES_ALGEBRAIC, xyz, uv, [1, 1, 0, 0, 1e-6], [2000, 2000, 2000, 2000, 10])
    assert round(O.reprojection_error(O.make_model(O.UCM, out, 512, 512), xyz, uv).mean, 3) == 0.145


def test_lm_recovers_same_model(O, cameras):
    """Self-consistency: data generated by a DS camera is fitted back to its own parameters."""
    c = cameras["double_sphere"]
    truth = oracle_model(O, c)
    xyz = O.synth_points3(0xACE50003, 0, 4000, np.cos(np.deg2rad(70.0)), False)
    uv, st = O.project(truth, xyz)
    assert np.all(st == 0)
    start = np.array(c["params"]) * np.array([1.02, 0.98, 1.01, 0.99, 0.9, 0.8])
    out, res = O.lm_solve(O.make_model(O.DS, start, 752, 480), O.RES_PIXEL, xyz, uv, None, None, _cfg(O, **TIGHT))
    assert np.allclose(out, c["params"], rtol=1e-9)


def test_reprojection_error_stats(O, cameras):
    """error_metrics.rs:62-121 vs numpy."""
    intr, uv, xyz = _kb_correspondences(O, cameras)
    m = O.make_model(O.DS, intr + [0.6, 0.05], 512, 512)
    e = O.reprojection_error(m, xyz, uv)
    p, st = O.project(m, xyz)
    err = np.linalg.norm(p[st == 0] - uv[st == 0], axis=1)
    assert e.count == len(err)
    assert np.isclose(e.mean, err.mean(), rtol=1e-13) and np.isclose(e.stddev, err.std(), rtol=1e-11)
    assert np.isclose(e.rmse, np.sqrt(np.mean(err ** 2)), rtol=1e-13)
    assert e.min == err.min() and e.max == err.max() and np.isclose(e.median, np.median(err), rtol=1e-15)
    behind = np.tile([0.0, 0.0, -1.0], (4, 1))
    with pytest.raises(ValueError):
        O.reprojection_error(m, behind, uv[:4])


def test_undistort_semantics(O):
    """undistort.rs:14-105: output pixel <- bilinear sample at project(pinhole ray); the last
    row/column can never be sampled (x1 >= W rejects); failures stay black."""
    W, H = 16, 12
    m = O.make_model(O.PINHOLE, [8.0, 8.0, 4.0, 2.0], W, H)  # powers of two: the map is exact
    img = O.synth_bytes(1, 0, W * H * 3).reshape(H, W, 3)
    mp = O.undistort_map(m, [8.0, 8.0, 4.0, 2.0])
    uu, vv = np.meshgrid(np.arange(W), np.arange(H))
    assert np.array_equal(mp[..., 0], uu) and np.array_equal(mp[..., 1], vv)
    out = O.undistort_rgb8(m, [8.0, 8.0, 4.0, 2.0], img, 1)
    assert np.array_equal(out[:-1, :-1], img[:-1, :-1])
    assert not out[-1].any() and not out[:, -1].any()
    out_n = O.undistort_rgb8(m, [8.0, 8.0, 4.0, 2.0], img, 0)
    assert np.array_equal(out_n, img)
    # half-pixel shift: bilinear average of horizontal neighbours, round half away from zero
    out_s = O.undistort_rgb8(m, [8.0, 8.0, 3.5, 2.0], img, 1)
    a = img[:-1, :-1].astype(np.float64); b = img[:-1, 1:].astype(np.float64)
    exp = np.floor((a * 0.5 * 1.0 + b * 0.5 * 1.0) + 0.5)
    assert np.array_equal(out_s[:-1, :-1], exp.astype(np.uint8))
    # threaded == scalar
    assert np.array_equal(O.undistort_rgb8(m, [8.0, 8.0, 3.5, 2.0], img, 1, nthreads=4), out_s)


def test_synth_generators(O):
    a = O.synth_points3(5, 0, 1000, np.cos(np.deg2rad(100.0)), True)
    b = O.synth_points3(5, 500, 500, np.cos(np.deg2rad(100.0)), True)
    assert np.array_equal(a[500:], b)  # counter-based: any sub-range reproduces
    rho = np.linalg.norm(a, axis=1)
    plain = np.ones(1000, bool); plain[63::64] = False
    assert np.all((rho[plain] >= 0.5) & (rho[plain] < 10.0))
    assert np.all(a[plain, 2] / rho[plain] >= np.cos(np.deg2rad(100.0)) - 1e-12)
    assert a[63].tolist() == [0.0, 0.0, 0.0] and a[127].tolist() == [0.1, 0.2, -1.0]
    px = O.synth_pixels(9, 0, 100, 752.0, 480.0)
    assert np.all((px[:, 0] >= 0) & (px[:, 0] < 752) & (px[:, 1] >= 0) & (px[:, 1] < 480))
    assert O.lib().orc_splitmix64(0) == 0xE220A8397B1DCDAF  # published splitmix64 first output


def test_point_jacobians_match_mpmath_and_differences(O, cameras):
    """2x3 Jacobian w.r.t. the 3-D point (trait doc mod.rs:246-252; no reference code): oracle vs mpmath
    50-digit differences of the model definitions, and vs central differences of orc_project."""
    from conftest import load_golden, oracle_model
    g = load_golden("mpmath_point_jacobians.json")
    for name in ("pinhole", "rad_tan", "kannala_brandt", "ucm", "eucm", "double_sphere", "fov"):
        m = oracle_model(O, cameras[name])
        for r in g[name]:
            st, uv, J = O.project_point_jacobian1(m, r["point"])
            assert st == 0
            Jm = np.array(r["J"])
            assert np.max(np.abs(J - Jm)) <= 1e-13 * np.abs(Jm).max(), (name, r["point"])
            X = np.array(r["point"])
            for k in range(3):
                d = np.zeros(3); d[k] = 1e-6
                _, up, _ = O.project_jacobian1(m, X + d); _, um, _ = O.project_jacobian1(m, X - d)
                assert np.allclose(J[:, k], (up - um) / 2e-6, rtol=1e-5, atol=1e-6 * np.abs(Jm).max())
    # invalid projection: status, zero Jacobian
    kb = oracle_model(O, cameras["kannala_brandt"])
    st, uv, J = O.project_point_jacobian1(kb, [0.1, 0.2, -1.0])
    assert st != 0 and np.all(J == 0.0)
    # on the KB axis: the analytic limit fx / z
    st, uv, J = O.project_point_jacobian1(kb, [0.0, 0.0, 2.0])
    assert st == 0 and np.allclose(J, [[kb.p[0] / 2.0, 0, 0], [0, kb.p[1] / 2.0, 0]])
