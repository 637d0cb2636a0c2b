"""Pin the CPU oracle against every known-answer test the reference holds for the hot path.

Each test restates one reference `#[test]` (file:line in the docstring) against the oracle.  The
reference has no golden *numbers* for this path, so these are its
behavioural KATs: round-trip tolerances, error classification, validator thresholds,
sample_points / linear_estimation behaviour.  `test_restated_vectors` additionally compares with
the float64 re-evaluation made independently during the survey.
"""
import math

import numpy as np
import pytest

from conftest import load_golden, oracle_model

SQRT_EPS = 2.0 ** -26


def _norm(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


def _roundtrip(O, m, X):
    st, uv = O.project1(m, X)
    assert st == O.OK
    st2, ray = O.unproject1(m, uv)
    assert st2 == O.OK
    return uv, ray


@pytest.mark.parametrize("cam,X,tol", [
    ("double_sphere", (0.5, -0.3, 2.0), 1e-6),        # double_sphere.rs:735-758
    ("rad_tan", (0.5, -0.3, 2.0), 1e-6),              # rad_tan.rs:866-890
    ("kannala_brandt_inline", (0.1, 0.2, 1.0), 1e-5), # kannala_brandt.rs:899-944
    ("ucm", (0.1, 0.1, 3.0), 1e-4),                   # ucm.rs:587-617
    ("eucm", (0.1, 0.1, 3.0), 1e-4),                  # eucm.rs:577-606
    ("fov", (0.1, 0.1, 3.0), 1e-4),                   # fov.rs:577-606
    ("ucm", (0.0, 0.0, 1.0), 1e-6),                   # ucm.rs:621-642
    ("eucm", (0.0, 0.0, 1.0), 1e-6),                  # eucm.rs:609-631
    ("fov", (0.0, 0.0, 1.0), 1e-6),                   # fov.rs:610-631
])
def test_project_unproject_round_trip(O, cameras, cam, X, tol):
    m = oracle_model(O, cameras[cam])
    uv, ray = _roundtrip(O, m, X)
    assert np.all(np.abs(ray - _norm(X)) < tol)
    if cam == "kannala_brandt_inline":
        assert 0 <= uv[0] < 752 and 0 <= uv[1] < 480


def test_pinhole_round_trip(O):
    """pinhole.rs:412-430 on samples/pinhole.yaml: (1,1,5), 1e-6."""
    cams = load_golden("cameras.json")
    m = oracle_model(O, cams["pinhole"])
    _, ray = _roundtrip(O, m, (1.0, 1.0, 5.0))
    assert np.all(np.abs(ray - _norm((1.0, 1.0, 5.0))) < 1e-6)


def test_pinhole_doc_test_numbers(O):
    """pinhole.rs:145-164, the one numeric known answer the reference itself states: fx = fy = 500, c = (320, 240),
    640 x 480, X = (0.1, 0.2, 1.0) -> (370, 340) within 1e-6."""
    m = O.make_model(O.PINHOLE, [500.0, 500.0, 320.0, 240.0], 640, 480)
    st, uv = O.project1(m, (0.1, 0.2, 1.0))
    assert st == O.OK
    assert abs(uv[0] - 370.0) < 1e-6 and abs(uv[1] - 340.0) < 1e-6
    assert uv[0] == 500.0 * 0.1 / 1.0 + 320.0 and uv[1] == 500.0 * 0.2 / 1.0 + 240.0  # the reference's operation order, exactly


def test_rad_tan_ten_points(O, cameras):
    """rad_tan.rs:894-943: directions preserved (dot > 0.99) over a spread of points."""
    m = oracle_model(O, cameras["rad_tan"])
    pts = [(0.1, 0.1, 1.0), (0.3, 0.0, 1.5), (-0.2, 0.3, 2.0), (-0.3, -0.2, 1.8), (0.15, -0.25, 2.5),
           (0.0, 0.0, 1.0), (0.2, 0.1, 1.5), (-0.1, -0.2, 2.0), (0.4, 0.2, 3.0), (-0.25, 0.15, 2.2)]
    ok = 0
    for X in pts:
        st, uv = O.project1(m, X)
        if st == O.OK:
            st2, ray = O.unproject1(m, uv)
            assert st2 == O.OK and float(np.dot(_norm(X), ray)) > 0.99
            ok += 1
    assert ok > 0


@pytest.mark.parametrize("cam,X,expected", [
    # behind camera (0.1,0.2,-1): double_sphere.rs:795-801, ucm.rs:678-684, eucm.rs:666-672,
    # kannala_brandt.rs:957-962 -> PointIsOutSideImage; fov.rs:660-666 -> PointAtCameraCenter
    ("double_sphere", (0.1, 0.2, -1.0), 1), ("ucm", (0.1, 0.2, -1.0), 1), ("eucm", (0.1, 0.2, -1.0), 1),
    ("kannala_brandt_inline", (0.1, 0.2, -1.0), 1), ("fov", (0.1, 0.2, -1.0), 2),
    # origin: double_sphere.rs:782-790 (+ucm/eucm) -> PointIsOutSideImage; kannala_brandt.rs:948-953,
    # fov.rs:648-655 -> PointAtCameraCenter
    ("double_sphere", (0.0, 0.0, 0.0), 1), ("ucm", (0.0, 0.0, 0.0), 1), ("eucm", (0.0, 0.0, 0.0), 1),
    ("kannala_brandt_inline", (0.0, 0.0, 0.0), 2), ("fov", (0.0, 0.0, 0.0), 2),
    ("pinhole", (0.0, 0.0, 0.0), 2), ("rad_tan", (0.0, 0.0, 0.0), 2),
    ("pinhole", (0.1, 0.2, -1.0), 2), ("rad_tan", (0.1, 0.2, -1.0), 2),
])
def test_error_classification(O, cameras, cam, X, expected):
    m = oracle_model(O, cameras[cam])
    st, uv = O.project1(m, X)
    assert st == expected
    assert np.all(np.isnan(uv))


def test_kb_unproject_out_of_bounds(O, cameras):
    """kannala_brandt.rs:966-974."""
    m = oracle_model(O, cameras["kannala_brandt_inline"])
    for uv in [(-1.0, 100.0), (752.0, 100.0), (100.0, -0.5), (100.0, 480.0)]:
        st, _ = O.unproject1(m, uv)
        assert st == O.POINT_OUTSIDE_IMAGE


def test_pinhole_unprojection_validates_bounds(O):
    """tests/projection_accuracy.rs:29-44 and consistency :47-70."""
    m = O.make_model(O.PINHOLE, [500.0, 500.0, 320.0, 240.0], 640, 480)
    assert O.unproject1(m, (-100.0, 100.0))[0] == O.POINT_OUTSIDE_IMAGE
    assert O.unproject1(m, (1000.0, 1000.0))[0] == O.POINT_OUTSIDE_IMAGE
    for X in [(0.0, 0.0, 1.0), (0.2, 0.1, 1.5), (-0.1, -0.2, 2.0)]:
        st, uv = O.project1(m, X)
        if st == O.OK:
            st2, ray = O.unproject1(m, uv)
            if st2 == O.OK:
                assert abs(float(np.dot(_norm(X), ray)) - 1.0) < 1e-6


def test_validator_thresholds(O):
    """mod.rs:629-679: bounds are half-open [0,W); z < sqrt(EPS) is the centre error (1e-10 fails,
    0.001 passes)."""
    m = O.make_model(O.PINHOLE, [500.0, 500.0, 320.0, 240.0], 640, 480)
    assert O.project1(m, (0.0, 0.0, 1e-10))[0] == O.POINT_AT_CENTER
    assert O.project1(m, (0.0, 0.0, 0.001))[0] == O.OK
    assert O.project1(m, (0.0, 0.0, SQRT_EPS))[0] == O.OK
    assert O.project1(m, (0.0, 0.0, np.nextafter(SQRT_EPS, 0.0)))[0] == O.POINT_AT_CENTER
    assert O.unproject1(m, (0.0, 0.0))[0] == O.OK
    assert O.unproject1(m, (640.0, 0.0))[0] == O.POINT_OUTSIDE_IMAGE
    assert O.unproject1(m, (np.nextafter(640.0, 0.0), np.nextafter(480.0, 0.0)))[0] == O.OK
    # projection that lands outside the image: pinhole/rad_tan test bounds (pinhole.rs:173-179)
    assert O.project1(m, (10.0, 0.0, 1.0))[0] == O.PROJECTION_OUTSIDE_IMAGE


FIVE = [(0.1, 0.1, 1.0), (0.3, 0.0, 1.5), (-0.2, 0.3, 2.0), (-0.3, -0.2, 1.8), (0.15, -0.25, 2.5)]


@pytest.mark.parametrize("cam", ["double_sphere", "kannala_brandt", "rad_tan", "ucm", "eucm"])
def test_model_conversions_basic_operations(O, cameras, cam):
    """tests/model_conversions.rs:18-131: five fixed points, directions preserved (dot > 0.99)."""
    c = cameras[cam]
    m = oracle_model(O, c)
    projected = 0
    for X in FIVE:
        st, uv = O.project1(m, X)
        if st != O.OK:
            continue
        projected += 1
        inside = 0 <= uv[0] < c["width"] and 0 <= uv[1] < c["height"]
        if cam in ("double_sphere", "kannala_brandt", "rad_tan"):
            assert inside
        if inside:
            st2, ray = O.unproject1(m, uv)
            if st2 == O.OK:
                assert float(np.dot(_norm(X), ray)) > 0.99
    assert projected > 0


def test_boundary_projections(O, cameras):
    """tests/projection_accuracy.rs:73-110."""
    m = oracle_model(O, cameras["double_sphere"])
    for X in [(0.5, 0.0, 2.0), (-0.5, 0.0, 2.0), (0.0, 0.5, 2.0), (0.0, -0.5, 2.0)]:
        st, uv = O.project1(m, X)
        assert st in (O.OK, O.PROJECTION_OUTSIDE_IMAGE)
        if st == O.OK:
            assert 0 <= uv[0] < 752 and 0 <= uv[1] < 480


def test_sample_points_behaviour(O, cameras):
    """src/util/mod.rs:70-95: sample_points(DS, 100) -> non-empty, equal counts, all z > 0."""
    m = oracle_model(O, cameras["double_sphere"])
    uv, xyz = O.sample_points(m, 100)
    assert len(uv) == len(xyz) > 0
    assert np.all(xyz[:, 2] > 0)
    assert np.allclose(np.linalg.norm(xyz, axis=1), 1.0, atol=1e-12)


def test_rad_tan_linear_estimation(O, cameras):
    """tests/parameter_estimation.rs:6-33, :36-59 (n=2 -> Err)."""
    c = cameras["rad_tan"]
    src = oracle_model(O, c)
    uv, xyz = O.sample_points(src, 50)
    est = O.make_model(O.RADTAN, c["params"][:4] + [0.0] * 5, c["width"], c["height"])
    assert O.linear_estimation(est, xyz, uv) == 0
    assert np.any(np.abs(est.params()[4:]) > 1e-10)
    uv2, xyz2 = O.sample_points(src, 2)
    est2 = O.make_model(O.RADTAN, c["params"][:4] + [0.0] * 5, c["width"], c["height"])
    assert len(uv2) < 3 and O.linear_estimation(est2, xyz2, uv2) != 0


def test_restated_vectors(O, cameras):
    """SURVEY.md section 8c known-answer table (independent float64 re-evaluation of the reference
    formulas made during the survey): must agree to the last bit printed."""
    kat = load_golden("restated_kats.json")
    for row in kat["project_unproject"]:
        m = oracle_model(O, cameras[row["camera"]])
        st, uv = O.project1(m, row["point"])
        assert st == O.OK
        assert uv.tolist() == row["uv"], row["camera"]
        if row.get("ray"):
            st2, ray = O.unproject1(m, uv)
            assert st2 == O.OK
            assert ray.tolist() == row["ray"], row["camera"]
    m = oracle_model(O, cameras["kannala_brandt"])
    for n_req, grid, kept in kat["sample_points_kb"]:
        total = O.lib().orc_sample_grid_size(__import__("ctypes").byref(m), n_req, None, None)
        uv, xyz = O.sample_points(m, n_req)
        assert total == grid and len(uv) == kept
    # KB project(unproject) closes to 6.1e-12 px on the 450 correspondences
    uv, xyz = O.sample_points(m, 500)
    back, st = O.project(m, xyz)
    assert np.all(st == 0)
    assert np.max(np.linalg.norm(back - uv, axis=1)) < 1e-11


def test_kb_unproject_hole_and_newton_quirks(O, cameras):
    """kannala_brandt.rs:474-534: 0 < ru <= 1e-6 fails, ru == 0 succeeds."""
    c = cameras["kannala_brandt"]
    m = oracle_model(O, c)
    fx, fy, cx, cy = c["params"][:4]
    st, ray = O.unproject1(m, (cx, cy))
    assert st == O.OK and ray.tolist() == [0.0, 0.0, 1.0]
    st, _ = O.unproject1(m, (cx + 0.5e-6 * fx, cy))
    assert st == O.NUMERICAL
    st, _ = O.unproject1(m, (cx + 2e-6 * fx, cy))
    assert st == O.OK


def test_eucm_precedence_quirk(O):
    """eucm.rs:196: bound is (1/beta)*(2a-1), not 1/(beta*(2a-1))."""
    m = O.make_model(O.EUCM, [100.0, 100.0, 0.0, 0.0, 0.8, 2.0], 0, 0)
    # (1/2)*(0.6) = 0.3 ; correct bound would be 1/(1.2) = 0.833.  det = 1 - 0.6*2*r2 >= 1e-3 => r2 <= 0.8325
    r2_ok, r2_bad = 0.29, 0.31
    assert O.unproject1(m, (100.0 * math.sqrt(r2_ok), 0.0))[0] == O.OK
    assert O.unproject1(m, (100.0 * math.sqrt(r2_bad), 0.0))[0] == O.POINT_OUTSIDE_IMAGE
