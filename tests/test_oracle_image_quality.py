"""Pin oracle/acm_oracle_image.c (reference src/util/image_quality.rs, src/util/validation.rs).

The reference holds no test and no golden number for these functions, so the pins are (i) the
independent float64 re-evaluation in plain Python loops (tests/golden/image_quality.json, made by
tests/golden/make_image_quality_golden.py) and (ii) closed-form cases."""
import math

import numpy as np

from conftest import load_golden, oracle_model


def _img(case, key):
    return np.array(case[key], dtype=np.uint8).reshape(case["H"], case["W"], 3)


def test_psnr_ssim_gray_match_the_restated_vectors(O):
    g = load_golden("image_quality.json")
    for c in g["cases"]:
        a, b = _img(c, "a"), _img(c, "b")
        want = math.inf if c["psnr"] == "inf" else c["psnr"]
        assert O.image_psnr(a, b) == want, c["mode"]
        assert O.image_ssim(a, b) == c["ssim"], c["mode"]          # same operations in the same order: same bits
        assert np.array_equal(O.rgb_to_grayscale(a).ravel(), c["gray_a"]), c["mode"]


def test_drawing_matches_the_restated_vectors(O):
    d = load_golden("image_quality.json")["draw"]
    pts = np.array([[float(v) for v in p] for p in d["points"]])
    img = O.draw_points(np.zeros((d["H"], d["W"], 3), np.uint8), pts, (255, 255, 255))
    assert np.array_equal(img.ravel(), d["image"])
    img2 = O.draw_points(np.zeros((d["H"], d["W"], 3), np.uint8), pts + [0.8, -0.6], (255, 255, 255))
    assert O.image_psnr(img, img2) == d["psnr_vs_shifted"] and O.image_ssim(img, img2) == d["ssim_vs_shifted"]
    assert int(O.rgb_to_grayscale(np.full((1, 1, 3), 255, np.uint8))[0, 0]) == d["luma_white"] == 255


def test_closed_form_cases(O):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (11, 13, 3), dtype=np.uint8)
    assert O.image_psnr(a, a) == math.inf and O.image_ssim(a, a) == 1.0   # image_quality.rs:83-84; ssim of x with x is 1
    z = np.zeros_like(a)
    assert O.image_psnr(z, z) == math.inf                                   # :77-79 no valid pixel
    # one white pixel against black: the only non-black pixel has 3 channels of difference 255 => mse = 255^2 => 0 dB
    one = z.copy(); one[4, 5] = 255
    assert O.image_psnr(one, z) == 0.0
    # a single interior disc: 13 pixels (dx^2 + dy^2 <= 4), clipped at the border to 6 (corner: dx, dy >= 0)
    assert int((O.draw_points(np.zeros((9, 9, 3), np.uint8), [[4.0, 4.0]], (1, 2, 3)).sum(axis=2) > 0).sum()) == 13
    assert int((O.draw_points(np.zeros((9, 9, 3), np.uint8), [[0.0, 0.0]], (1, 2, 3)).sum(axis=2) > 0).sum()) == 6
    # images narrower than 3 pixels have no window: 1.0 (:184-188)
    assert O.image_ssim(a[:2], a[:2] // 2) == 1.0


def test_image_quality_metrics_pipeline(O, cameras):
    """compute_image_quality_metrics (image_quality.rs:254-324) on the converter's own data: the KB sample
    against itself is a perfect match; against a perturbed model the images differ."""
    kb = oracle_model(O, cameras["kannala_brandt"])
    _, xyz = O.sample_points(kb, 500)
    kept, psnr, ssim, comb = O.image_quality_metrics(kb, kb, xyz, kb.width, kb.height, want_image=True)
    assert kept == len(xyz) == 450 and psnr == math.inf and ssim == 1.0
    assert set(map(tuple, comb.reshape(-1, 3))) == {(0, 0, 0), (255, 0, 255)}          # magenta drawn over green everywhere
    ds = oracle_model(O, cameras["double_sphere"])
    ds.width, ds.height = kb.width, kb.height
    kept2, psnr2, ssim2, comb2 = O.image_quality_metrics(kb, ds, xyz, kb.width, kb.height, want_image=True)
    assert 0 < kept2 <= 450 and math.isfinite(psnr2) and 0.0 < ssim2 < 1.0
    assert {(0, 255, 0), (255, 0, 255)} <= set(map(tuple, comb2.reshape(-1, 3)))
    # nothing projects: zero kept points (the reference returns ZeroProjectionPoints, :306-308)
    kept3, p3, s3, _ = O.image_quality_metrics(kb, kb, -np.abs(xyz), kb.width, kb.height)
    assert kept3 == 0 and math.isnan(p3) and math.isnan(s3)


def test_validate_conversion_accuracy(O, cameras):
    """validation.rs:93-213: a model against itself has zero error in all five regions."""
    kb = oracle_model(O, cameras["kannala_brandt"])
    valid, err, avg, mx = O.validate_conversion(kb, kb)
    fin = ~np.isnan(err)   # the far-edge probe (0.95 W, 0.95 H) of the 512^2 KB sample lies outside its field of view
    assert valid == fin.sum() >= 4 and np.all(err[fin] == 0.0) and avg == 0.0 and mx == 0.0
    ds = oracle_model(O, cameras["double_sphere"])
    valid, err, avg, mx = O.validate_conversion(ds, kb)
    fin = err[~np.isnan(err)]
    assert valid == len(fin) and (valid == 0 or (avg == fin.sum() / valid and mx == fin.max()))
