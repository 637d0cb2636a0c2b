"""Generate tests/golden/opencv_cross.json and tests/golden/mpmath_jacobians.json.

Independent implementations used to pin the oracle where the reference ships no numbers:
  * OpenCV (cv2.fisheye.projectPoints == Kannala-Brandt, cv2.projectPoints == RadTan): same
    camera models written by other people.
  * mpmath (50 digits): the seven projection functions written directly from the model
    definitions, differentiated numerically in high precision -> parameter Jacobians.

Run here (no GPU needed):  python tests/golden/make_golden.py
"""
import json
import os

import cv2
import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
cams = json.load(open(os.path.join(HERE, "cameras.json")))


def opencv_cross():
    rng = np.random.default_rng(0xACE5)
    out = {}
    pts = np.stack([rng.uniform(-0.6, 0.6, 64), rng.uniform(-0.4, 0.4, 64), rng.uniform(1.0, 4.0, 64)], axis=1)
    z = np.zeros(3)
    for name in ("kannala_brandt", "kannala_brandt_inline"):
        p = cams[name]["params"]
        K = np.array([[p[0], 0, p[2]], [0, p[1], p[3]], [0, 0, 1.0]])
        uv, _ = cv2.fisheye.projectPoints(pts.reshape(1, -1, 3), z, z, K, np.array(p[4:8]))
        out[name] = {"points": pts.tolist(), "uv": uv.reshape(-1, 2).tolist()}
    p = cams["rad_tan"]["params"]
    K = np.array([[p[0], 0, p[2]], [0, p[1], p[3]], [0, 0, 1.0]])
    uv, _ = cv2.projectPoints(pts, z, z, K, np.array(p[4:9]))  # (k1,k2,p1,p2,k3)
    out["rad_tan"] = {"points": pts.tolist(), "uv": uv.reshape(-1, 2).tolist()}
    p = cams["pinhole"]["params"]
    K = np.array([[p[0], 0, p[2]], [0, p[1], p[3]], [0, 0, 1.0]])
    uv, _ = cv2.projectPoints(pts, z, z, K, None)
    out["pinhole"] = {"points": pts.tolist(), "uv": uv.reshape(-1, 2).tolist()}
    out["_comment"] = "cv2 %s; see make_golden.py" % cv2.__version__
    return out


def project_mp(model_id, p, X):
    """Model definitions (Usenko et al. 2018 'The Double Sphere Camera Model' section 2-4 for
    UCM/EUCM/DS/KB/FOV; Brown-Conrady for RadTan) in arbitrary precision."""
    fx, fy, cx, cy = p[:4]
    x, y, z = X
    if model_id == 0:
        mx, my = x / z, y / z
    elif model_id == 1:
        k1, k2, p1, p2, k3 = p[4:9]
        a, b = x / z, y / z
        r2 = a * a + b * b
        rad = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
        mx = a * rad + 2 * p1 * a * b + p2 * (r2 + 2 * a * a)
        my = b * rad + p1 * (r2 + 2 * b * b) + 2 * p2 * a * b
    elif model_id == 2:
        k1, k2, k3, k4 = p[4:8]
        r = mp.sqrt(x * x + y * y)
        th = mp.atan2(r, z)
        thd = th + k1 * th ** 3 + k2 * th ** 5 + k3 * th ** 7 + k4 * th ** 9
        mx, my = thd * x / r, thd * y / r
    elif model_id == 3:
        al = p[4]
        d = mp.sqrt(x * x + y * y + z * z)
        den = al * d + (1 - al) * z
        mx, my = x / den, y / den
    elif model_id == 4:
        al, be = p[4:6]
        d = mp.sqrt(be * (x * x + y * y) + z * z)
        den = al * d + (1 - al) * z
        mx, my = x / den, y / den
    elif model_id == 5:
        al, xi = p[4:6]
        d1 = mp.sqrt(x * x + y * y + z * z)
        g = xi * d1 + z
        d2 = mp.sqrt(x * x + y * y + g * g)
        den = al * d2 + (1 - al) * g
        mx, my = x / den, y / den
    elif model_id == 6:
        w = p[4]
        r = mp.sqrt(x * x + y * y)
        rd = mp.atan2(2 * mp.tan(w / 2) * r, z) / (r * w)
        mx, my = x * rd, y * rd
    return fx * mx + cx, fy * my + cy


def mp_jacobians():
    mp.mp.dps = 50
    out = {"_comment": "mpmath 50-digit central differences (h=1e-20) of the model definitions"}
    pts = [(0.3, -0.2, 1.5), (-0.7, 0.4, 1.0), (0.05, 0.02, 2.0), (1.2, 0.9, 0.8)]
    for name in ("pinhole", "rad_tan", "kannala_brandt", "ucm", "eucm", "double_sphere", "fov"):
        c = cams[name]
        p0 = [mp.mpf(repr(v)) if False else mp.mpf(v) for v in c["params"]]
        rows = []
        for X in pts:
            Xm = [mp.mpf(v) for v in X]
            u0, v0 = project_mp(c["model_id"], p0, Xm)
            J = [[], []]
            h = mp.mpf(10) ** -20
            for k in range(len(p0)):
                pp = list(p0); pm = list(p0)
                pp[k] += h; pm[k] -= h
                up, vp = project_mp(c["model_id"], pp, Xm)
                um, vm = project_mp(c["model_id"], pm, Xm)
                J[0].append(float((up - um) / (2 * h)))
                J[1].append(float((vp - vm) / (2 * h)))
            rows.append({"point": list(X), "uv": [float(u0), float(v0)], "J": J})
        out[name] = rows
    return out


def mp_point_jacobians():
    """2x3 Jacobians of (u, v) w.r.t. the 3-D point: 50-digit central differences of the same model definitions."""
    mp.mp.dps = 50
    out = {"_comment": "mpmath 50-digit central differences (h=1e-20) of the model definitions w.r.t. the 3-D point"}
    pts = [(0.3, -0.2, 1.5), (-0.7, 0.4, 1.0), (0.05, 0.02, 2.0), (1.2, 0.9, 0.8), (0.0, 0.3, 1.0)]
    for name in ("pinhole", "rad_tan", "kannala_brandt", "ucm", "eucm", "double_sphere", "fov"):
        c = cams[name]
        p0 = [mp.mpf(v) for v in c["params"]]
        rows = []
        for X in pts:
            Xm = [mp.mpf(v) for v in X]
            J = [[], []]
            h = mp.mpf(10) ** -20
            for k in range(3):
                Xp = list(Xm); Xn = list(Xm)
                Xp[k] += h; Xn[k] -= h
                up, vp = project_mp(c["model_id"], p0, Xp)
                um, vm = project_mp(c["model_id"], p0, Xn)
                J[0].append(float((up - um) / (2 * h)))
                J[1].append(float((vp - vm) / (2 * h)))
            rows.append({"point": list(X), "J": J})
        out[name] = rows
    return out


if __name__ == "__main__":
    json.dump(opencv_cross(), open(os.path.join(HERE, "opencv_cross.json"), "w"), indent=1)
    json.dump(mp_jacobians(), open(os.path.join(HERE, "mpmath_jacobians.json"), "w"), indent=1)
    json.dump(mp_point_jacobians(), open(os.path.join(HERE, "mpmath_point_jacobians.json"), "w"), indent=1)
    print("wrote opencv_cross.json, mpmath_jacobians.json, mpmath_point_jacobians.json")
