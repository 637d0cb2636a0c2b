"""GPU parity of the image-quality diagnostics against the oracle.
PSNR is integer arithmetic up to the last division (bit-exact); the SSIM window term is evaluated
with the reference's operations in the reference's order, only the order of the sum over windows
differs (1e-12); drawings and the combined display image are byte-exact."""
import math

import numpy as np
import pytest

from conftest import load_golden, oracle_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acm():
    import apex_camera_models_b200 as m
    return m


@pytest.fixture(scope="module")
def ctx(acm):
    c = acm.Context(0)
    yield c
    c.close()


def gpu_model(acm, ctx, cam, width=None, height=None):
    cls = acm.MODEL_CLASSES[cam["model_id"]]
    return cls(acm.Intrinsics(*cam["params"][:4]), acm.Resolution(width or cam["width"], height or cam["height"]), cam["params"][4:], ctx=ctx)


def same(a, b, rtol=0.0):
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= rtol * abs(b)


@pytest.mark.parametrize("W,H", [(16, 12), (3, 3), (2, 9), (1, 1), (33, 9), (97, 41), (640, 480), (1023, 517)])
def test_psnr_ssim_match_oracle(acm, ctx, O, W, H):
    rng = np.random.default_rng(W * 1000 + H)
    a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-30, 31, a.shape), 0, 255).astype(np.uint8)
    hole = rng.random((H, W)) < 0.2
    a[hole] = 0; b[hole & (rng.random((H, W)) < 0.7)] = 0     # pixels black in both images are skipped by PSNR
    assert acm.calculate_psnr(a, b, ctx) == O.image_psnr(a, b)
    assert same(acm.calculate_ssim(a, b, ctx), O.image_ssim(a, b), 1e-12)
    assert acm.calculate_psnr(a, a, ctx) == math.inf and same(acm.calculate_ssim(a, a, ctx), 1.0, 1e-12)
    z = np.zeros_like(a)
    assert acm.calculate_psnr(z, z, ctx) == math.inf
    with pytest.raises(acm.UtilError):
        acm.calculate_psnr(a, np.zeros((H + 1, W, 3), np.uint8), ctx)


def test_golden_vectors_on_gpu(acm, ctx):
    g = load_golden("image_quality.json")
    for c in g["cases"]:
        a = np.array(c["a"], np.uint8).reshape(c["H"], c["W"], 3); b = np.array(c["b"], np.uint8).reshape(c["H"], c["W"], 3)
        want = math.inf if c["psnr"] == "inf" else c["psnr"]
        assert acm.calculate_psnr(a, b, ctx) == want
        assert same(acm.calculate_ssim(a, b, ctx), c["ssim"], 1e-12)
    d = g["draw"]
    pts = np.array([[float(v) for v in p] for p in d["points"]])
    img = acm.create_projection_image(pts, (255, 255, 255), d["W"], d["H"], ctx)
    assert np.array_equal(img.ravel(), d["image"])


def test_drawings_are_byte_exact(acm, ctx, O):
    rng = np.random.default_rng(77)
    W, H = 211, 157
    pin = rng.uniform(-5, [W + 5, H + 5], (4000, 2)); pout = pin + rng.normal(0, 1.5, pin.shape)
    pin[::97] = np.floor(pin[::97]) + 0.5                      # rounding ties
    ref = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    want = O.draw_points(O.draw_points(ref.copy(), pin, (0, 255, 0)), pout, (255, 0, 255))
    assert np.array_equal(acm.create_combined_projection_image_on_reference(pin, pout, ref, ctx), want)
    want = O.draw_points(O.draw_points(np.zeros_like(ref), pin, (0, 255, 0)), pout, (255, 0, 255))
    assert np.array_equal(acm.create_combined_projection_image(pin, pout, W, H, ctx), want)
    assert np.array_equal(acm.model_projection_visualization(pin, ref, (W, H), ctx), O.draw_points(ref.copy(), pin, (0, 255, 0)))
    assert np.array_equal(acm.model_projection_visualization(pin, None, (W, H), ctx), O.draw_points(np.zeros_like(ref), pin, (0, 255, 0)))


@pytest.mark.parametrize("target", ["double_sphere", "ucm", "eucm", "fov", "rad_tan", "kannala_brandt"])
def test_image_quality_metrics_match_oracle(acm, ctx, O, cameras, target):
    """compute_image_quality_metrics of the converter's KB sample against every target family
    (the targets carry the sample parameters of their family on the KB resolution)."""
    kbc = cameras["kannala_brandt"]
    kb, okb = gpu_model(acm, ctx, kbc), oracle_model(O, kbc)
    _, xyz = O.sample_points(okb, 10_000)
    tc = cameras[target]
    tm = gpu_model(acm, ctx, tc, kbc["width"], kbc["height"])
    otm = O.make_model(tc["model_id"], tc["params"], kbc["width"], kbc["height"])
    rng = np.random.default_rng(3)
    ref = rng.integers(0, 256, (kbc["height"], kbc["width"], 3), dtype=np.uint8)
    for reference in (None, ref):
        kept, psnr, ssim, comb = O.image_quality_metrics(okb, otm, xyz, kbc["width"], kbc["height"], reference, want_image=True)
        if kept == 0:
            with pytest.raises(acm.ZeroProjectionPoints):
                acm.compute_image_quality_metrics(kb, tm, xyz, reference)
            continue
        m, img = acm.compute_image_quality_metrics(kb, tm, xyz, reference, return_image=True)
        assert m.psnr == psnr and same(m.ssim, ssim, 1e-12), (m, psnr, ssim)
        assert np.array_equal(img, comb)
    # device-resident points give the same answer
    X = acm.Points.from_numpy(ctx, xyz)
    if kept:
        assert acm.compute_image_quality_metrics(kb, tm, X).psnr == psnr
    X.free()


def test_zero_projection_points(acm, ctx, cameras):
    kb = gpu_model(acm, ctx, cameras["kannala_brandt"])
    behind = np.array([[0.1, 0.2, -1.0], [0.0, 0.0, -2.0]])
    with pytest.raises(acm.ZeroProjectionPoints):
        acm.compute_image_quality_metrics(kb, kb, behind)


def test_large_image_many_points(acm, ctx, O, cameras):
    """4096^2 image, 1 M points: size-independent properties (a model against itself is a perfect
    match; PSNR from the pixel counts of the two drawings) plus the oracle on the same input."""
    kbc = dict(cameras["kannala_brandt"])
    s = 8.0
    kbc["params"] = [p * s for p in kbc["params"][:4]] + kbc["params"][4:]
    kbc["width"], kbc["height"] = 4096, 4096
    kb, okb = gpu_model(acm, ctx, kbc), oracle_model(O, kbc)
    uv, xyz = acm.sample_points(kb, 1_000_000, device=True)
    m = acm.compute_image_quality_metrics(kb, kb, xyz)
    assert m.psnr == math.inf and same(m.ssim, 1.0, 1e-12)
    ds = dict(cameras["double_sphere"]); ds["params"] = kbc["params"][:4] + ds["params"][4:]   # KB's (scaled) intrinsics, DS distortion
    dsm = gpu_model(acm, ctx, ds, 4096, 4096)
    ods = O.make_model(ds["model_id"], ds["params"], 4096, 4096)
    m2 = acm.compute_image_quality_metrics(kb, dsm, xyz)
    kept, psnr, ssim, _ = O.image_quality_metrics(okb, ods, xyz.numpy(), 4096, 4096)
    # 16.7 M window terms of mixed sign: the reference's (and the oracle's) sequential f64 sum carries ~1e-11 of
    # rounding error itself, the device's tree sum less; the contract's bar (1e-9 relative) applies
    assert kept > 0 and m2.psnr == psnr and same(m2.ssim, ssim, 1e-9)
    uv.free(); xyz.free()


@pytest.mark.parametrize("target", ["double_sphere", "ucm", "eucm", "fov", "rad_tan", "pinhole"])
def test_validate_conversion_accuracy_matches_oracle(acm, ctx, O, cameras, target):
    """util::validate_conversion_accuracy (validation.rs:93-213): five probe pixels -> unproject(input) ->
    project(both) -> distance; the host function over the GPU batch entry points vs the oracle's scalar loop."""
    from apex_camera_models_b200.camera_converter import validate_conversion_accuracy
    for src in ("kannala_brandt", "double_sphere", "rad_tan"):
        sc, tc = cameras[src], cameras[target]
        im, om_in = gpu_model(acm, ctx, sc), oracle_model(O, sc)
        tm = gpu_model(acm, ctx, tc, sc["width"], sc["height"])
        om_out = O.make_model(tc["model_id"], tc["params"], sc["width"], sc["height"])
        got = validate_conversion_accuracy(tm, im)
        valid, err, avg, mx = O.validate_conversion(om_out, om_in)
        e = np.array(got.region_errors, dtype=np.float64)
        assert np.array_equal(np.isnan(e), np.isnan(err)), (src, target, e, err)
        fin = ~np.isnan(err)
        assert np.allclose(e[fin], err[fin], rtol=1e-9, atol=1e-12)
        if valid:
            assert np.isclose(got.average_error, avg, rtol=1e-9, atol=1e-12) and np.isclose(got.max_error, mx, rtol=1e-9, atol=1e-12)
            want = "EXCELLENT" if avg < 0.001 else "GOOD" if avg < 0.1 else "NEEDS IMPROVEMENT"
            assert got.status == want
        else:
            assert math.isnan(got.average_error) and got.status == "NEEDS IMPROVEMENT"
