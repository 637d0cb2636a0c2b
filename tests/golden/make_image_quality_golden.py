"""Independent float64 re-evaluation of the reference's image-quality functions
(/root/reference/src/util/image_quality.rs: calculate_psnr :45-89, calculate_ssim :108-189,
rgb_to_grayscale :194-210, the radius-2 disc of create_projection_image :338-373) in plain Python
loops, in the reference's operation order.  Writes tests/golden/image_quality.json, which pins
oracle/acm_oracle_image.c (the reference holds no test or golden number for these functions and
cannot be built here).  Deterministic: inputs come from numpy's PCG64 with fixed seeds and are
stored in the fixture."""
import json, math, os
import numpy as np


def psnr(a, b):
    H, W, _ = a.shape
    mse, valid = 0.0, 0
    for y in range(H):
        for x in range(W):
            p, q = a[y, x], b[y, x]
            if any(int(v) != 0 for v in p) or any(int(v) != 0 for v in q):
                for c in range(3):
                    d = float(p[c]) - float(q[c])
                    mse += d * d
                valid += 3
    if valid == 0:
        return math.inf
    mse /= float(valid)
    if mse <= 1e-10:
        return math.inf
    return 10.0 * math.log10(255.0 * 255.0 / mse)


def gray(img):
    H, W, _ = img.shape
    g = np.zeros((H, W), dtype=np.uint8)
    for y in range(H):
        for x in range(W):
            v = 0.299 * float(img[y, x, 0]) + 0.587 * float(img[y, x, 1]) + 0.114 * float(img[y, x, 2])
            g[y, x] = min(255, max(0, int(v)))  # `as u8`: truncate toward zero, saturate
    return g


def ssim(a, b):
    g1, g2 = gray(a), gray(b)
    c1 = (0.01 * 255.0) * (0.01 * 255.0)
    c2 = (0.03 * 255.0) * (0.03 * 255.0)
    H, W = g1.shape
    s, count = 0.0, 0
    for y in range(1, H - 1):
        for x in range(1, W - 1):
            l1 = l2 = 0.0
            n = 0
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    l1 += float(g1[y + dy, x + dx]); l2 += float(g2[y + dy, x + dx]); n += 1
            mu1, mu2 = l1 / float(n), l2 / float(n)
            s1 = s2 = s12 = 0.0
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    v1, v2 = float(g1[y + dy, x + dx]), float(g2[y + dy, x + dx])
                    s1 += (v1 - mu1) * (v1 - mu1); s2 += (v2 - mu2) * (v2 - mu2); s12 += (v1 - mu1) * (v2 - mu2)
            s1 /= float(n - 1); s2 /= float(n - 1); s12 /= float(n - 1)
            num = (2.0 * mu1 * mu2 + c1) * (2.0 * s12 + c2)
            den = (mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2)
            if den > 0.0:
                s += num / den; count += 1
    return s / float(count) if count > 0 else 1.0


def rust_round(v):  # f64::round: half away from zero; `as i32` saturates, NaN -> 0
    if v != v:
        return 0
    r = math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)
    return int(max(-2**31, min(2**31 - 1, r)))


def draw(img, pts, color):
    H, W, _ = img.shape
    for (u, v) in pts:
        cx, cy = rust_round(u), rust_round(v)
        for dy in range(-2, 3):
            for dx in range(-2, 3):
                if dx * dx + dy * dy <= 4:
                    x, y = cx + dx, cy + dy
                    if 0 <= x < W and 0 <= y < H:
                        img[y, x] = color
    return img


def main():
    rng = np.random.default_rng(0xACE5F4)
    cases = []
    for (W, H, mode) in [(16, 12, "random"), (9, 7, "sparse"), (24, 10, "smooth"), (5, 5, "identical"), (3, 3, "tiny"), (8, 6, "black")]:
        if mode == "random":
            a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8); b = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        elif mode == "sparse":
            a = np.where(rng.random((H, W, 1)) < 0.3, rng.integers(0, 256, (H, W, 3)), 0).astype(np.uint8)
            b = np.where(rng.random((H, W, 1)) < 0.3, rng.integers(0, 256, (H, W, 3)), 0).astype(np.uint8)
        elif mode == "smooth":
            yy, xx = np.mgrid[0:H, 0:W]
            a = np.stack([(xx * 10) % 256, (yy * 20) % 256, ((xx + yy) * 7) % 256], axis=2).astype(np.uint8)
            b = np.clip(a.astype(int) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
        elif mode == "identical":
            a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8); b = a.copy()
        elif mode == "tiny":
            a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8); b = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        else:
            a = np.zeros((H, W, 3), np.uint8); b = np.zeros((H, W, 3), np.uint8)
        p = psnr(a, b)
        cases.append({"mode": mode, "W": W, "H": H, "a": a.ravel().tolist(), "b": b.ravel().tolist(),
                      "psnr": ("inf" if math.isinf(p) else p), "ssim": ssim(a, b), "gray_a": gray(a).ravel().tolist()})
    # drawing: discs incl. clipped ones, ties at .5, negative and far-outside centres
    W, H = 20, 14
    pts = [[3.4, 4.6], [0.5, 0.5], [-0.5, 2.5], [19.49, 13.5], [10.5, -1.5], [25.0, 7.0], [7.0, 7.0], [8.2, 7.9], [float("nan"), 5.0], [1e12, 3.0], [-1e12, 3.0]]
    img = draw(np.zeros((H, W, 3), np.uint8), pts, (255, 255, 255))
    img2 = draw(np.zeros((H, W, 3), np.uint8), [[p[0] + 0.8, p[1] - 0.6] for p in pts], (255, 255, 255))
    white = {"W": W, "H": H, "points": [[("nan" if v != v else v) for v in p] for p in pts], "image": img.ravel().tolist(),
             "psnr_vs_shifted": psnr(img, img2), "ssim_vs_shifted": ssim(img, img2), "luma_white": int(gray(np.full((1, 1, 3), 255, np.uint8))[0, 0])}
    out = {"_comment": "made by tests/golden/make_image_quality_golden.py (pure-Python restatement of reference src/util/image_quality.rs)",
           "cases": cases, "draw": white}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "image_quality.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
