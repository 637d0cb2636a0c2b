// C++ host-layer test: the reference's own unit tests for the path, restated against include/acm.hpp
// (double_sphere.rs:735-801, kannala_brandt.rs:899-974, pinhole.rs:412-430, tests/model_conversions.rs,
// tests/projection_accuracy.rs, tests/parameter_estimation.rs).  Needs a GPU; prints "HOST_TEST_OK".
#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "acm.hpp"

using namespace acm;

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)

template <class F> static ErrorKind kind_of(F&& f) {
    try { f(); } catch (const CameraModelError& e) { return e.kind; }
    std::fprintf(stderr, "expected an error\n"); std::exit(1);
}

static Vector3 unit(const Vector3& p) { double n = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]); return {p[0] / n, p[1] / n, p[2] / n}; }

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : ".";
    Context ctx(0);

    // --- new(): length check for all, validation for Pinhole / RadTan only
    CHECK(kind_of([] { DoubleSphereModel::create({500.0, 500.0}); }) == ErrorKind::InvalidParams);
    CHECK(kind_of([] { PinholeModel::create({-500.0, 500.0, 320.0, 240.0}); }) == ErrorKind::FocalLengthMustBePositive);
    CHECK(kind_of([] { PinholeModel::create({500.0, 500.0, INFINITY, 240.0}); }) == ErrorKind::PrincipalPointMustBeFinite);
    DoubleSphereModel bad = DoubleSphereModel::create({500.0, 500.0, 320.0, 240.0, 1.5, 0.0});  // not validated in new()
    CHECK(kind_of([&] { bad.validate_params(); }) == ErrorKind::InvalidParams);

    // --- YAML round trip + the sample Double Sphere camera
    DoubleSphereModel ds = DoubleSphereModel::create({348.112754378549, 347.1109973814674, 365.8121721753254, 249.3555778487899, 0.5657413673629862, -0.24425190195168348});
    ds.resolution = {752, 480};
    ds.save_to_yaml(dir + "/ds.yaml");
    DoubleSphereModel ds2 = DoubleSphereModel::load_from_yaml(dir + "/ds.yaml");
    CHECK(ds2.intrinsics.fx == 348.112754378549 && ds2.distortions[1] == -0.24425190195168348 && ds2.resolution.width == 752);
    ds2.bind(ctx);

    // double_sphere.rs:735-758 round trip, :782-801 error classification
    Vector3 p{0.5, -0.3, 2.0};
    Vector2 uv = ds2.project(p);
    CHECK(uv[0] == 477.8635401766414 && uv[1] == 182.3182258143974);  // tests/golden/restated_kats.json
    Vector3 ray = ds2.unproject(uv), d = unit(p);
    for (int i = 0; i < 3; ++i) CHECK(std::fabs(ray[i] - d[i]) < 1e-6);
    CHECK(kind_of([&] { ds2.project({0.1, 0.2, -1.0}); }) == ErrorKind::PointIsOutSideImage);
    CHECK(kind_of([&] { ds2.project({0.0, 0.0, 0.0}); }) == ErrorKind::PointIsOutSideImage);

    // kannala_brandt.rs:948-974 + YAML asymmetry (saves distortion_coeffs, loads distortion)
    KannalaBrandtModel kb = KannalaBrandtModel::create({190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504,
                                                        0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182});
    kb.resolution = {512, 512};
    kb.bind(ctx);
    CHECK(kind_of([&] { kb.project({0.0, 0.0, 0.0}); }) == ErrorKind::PointAtCameraCenter);
    CHECK(kind_of([&] { kb.unproject({-1.0, 100.0}); }) == ErrorKind::PointIsOutSideImage);
    kb.save_to_yaml(dir + "/kb.yaml");
    CHECK(kind_of([&] { KannalaBrandtModel::load_from_yaml(dir + "/kb.yaml"); }) == ErrorKind::InvalidParams);

    // pinhole: tests/projection_accuracy.rs:29-70
    PinholeModel pin = PinholeModel::create({500.0, 500.0, 320.0, 240.0});
    pin.resolution = {640, 480};
    pin.bind(ctx);
    CHECK(kind_of([&] { pin.unproject({-100.0, 100.0}); }) == ErrorKind::PointIsOutSideImage);
    CHECK(kind_of([&] { pin.project({10.0, 0.0, 1.0}); }) == ErrorKind::ProjectionOutSideImage);

    // --- the converter's config 1: sample_points(KB, 500) -> 450 correspondences -> KB -> DS
    auto [pts2, pts3] = sample_points(kb, 500);
    const size_t n = pts2.size() / 2;
    CHECK(n == 450 && pts3.size() == 3 * n);
    DoubleSphereModel target = DoubleSphereModel::create({kb.intrinsics.fx, kb.intrinsics.fy, kb.intrinsics.cx, kb.intrinsics.cy, 0.5, 0.1});
    target.resolution = kb.resolution;
    target.bind(ctx);
    CHECK(std::fabs(compute_reprojection_error(target, pts3.data(), pts2.data(), n).mean - 10.0321) < 1e-3);
    OptimizationCost cost(target, pts3.data(), pts2.data(), n);
    cost.linear_estimation();
    CHECK(std::fabs(target.distortions[0] - 0.6467229596331426) < 1e-12 && target.distortions[1] == 0.0);
    const double lo[6] = {1, 1, 0, 0, 1e-6, -5}, hi[6] = {2000, 2000, 2000, 2000, 1, 5};  // camera_converter.rs:395-400
    acm_lm_result res = cost.optimize(lo, hi);
    CHECK(res.status == 0 && res.iterations < 100);
    double mean = compute_reprojection_error(target, pts3.data(), pts2.data(), n).mean;
    CHECK(std::fabs(mean - 0.0077324) < 1e-5);  // README.md:163 "0.008 px"

    // tests/parameter_estimation.rs: RadTan linear estimation on 50 sampled points
    RadTanModel rt = RadTanModel::create({461.629, 460.152, 362.680, 246.049, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0});
    rt.resolution = {752, 480};
    rt.bind(ctx);
    auto [r2, r3] = sample_points(rt, 50);
    RadTanModel est = RadTanModel::create({461.629, 460.152, 362.680, 246.049, 0, 0, 0, 0, 0});
    est.resolution = rt.resolution;
    est.bind(ctx);
    {
        Points X(ctx, 3, r3.data(), r3.size() / 3), U(ctx, 2, r2.data(), r2.size() / 2);
        est.linear_estimation(X, U);
    }
    CHECK(std::fabs(est.distortions[0]) > 1e-10 && est.distortions[2] == 0.0 && est.distortions[3] == 0.0);

    // undistort: identity-like pinhole map keeps interior pixels (undistort.rs:69-103)
    PinholeModel small = PinholeModel::create({8.0, 8.0, 4.0, 2.0});
    small.resolution = {16, 12};
    small.bind(ctx);
    std::vector<uint8_t> img(16 * 12 * 3);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + 3);
    auto out = undistort_image(img, 16, 12, small);
    for (int v = 0; v < 11; ++v) for (int u = 0; u < 15; ++u) for (int c = 0; c < 3; ++c) CHECK(out[3 * (v * 16 + u) + c] == img[3 * (v * 16 + u) + c]);
    for (int u = 0; u < 16; ++u) CHECK(out[3 * (11 * 16 + u)] == 0);
    CHECK(kind_of([&] { undistort_image(img, 15, 12, small); }) == ErrorKind::InvalidParams);

    // image-quality diagnostics (image_quality.rs): a model against itself is a perfect match; one white pixel
    // against black is 0 dB; the converted DS model is close to, but not identical with, the KB input
    {
        ImageQualityMetrics q = compute_image_quality_metrics(kb, kb, pts3.data(), n, 512, 512);
        CHECK(std::isinf(q.psnr) && q.psnr > 0 && std::fabs(q.ssim - 1.0) < 1e-12);
        std::vector<uint8_t> comb;
        ImageQualityMetrics q2 = compute_image_quality_metrics(kb, target, pts3.data(), n, 512, 512, nullptr, &comb);
        CHECK(q2.psnr > 0.0 && q2.ssim > 0.9 && q2.ssim <= 1.0 && comb.size() == 512 * 512 * 3);
        std::vector<uint8_t> z(9 * 7 * 3, 0), one(9 * 7 * 3, 0);
        one[3 * (3 * 9 + 4)] = one[3 * (3 * 9 + 4) + 1] = one[3 * (3 * 9 + 4) + 2] = 255;
        CHECK(calculate_psnr(ctx, one, z, 9, 7) == 0.0 && std::isinf(calculate_psnr(ctx, z, z, 9, 7)));
        CHECK(std::fabs(calculate_ssim(ctx, one, one, 9, 7) - 1.0) < 1e-12);
        const double c[2] = {4.0, 4.0};
        auto disc = create_projection_image(ctx, c, 1, 255, 255, 255, 9, 9);
        int lit = 0; for (size_t i = 0; i < disc.size(); i += 3) lit += disc[i] != 0;
        CHECK(lit == 13);
        CHECK(kind_of([&] { calculate_psnr(ctx, one, z, 9, 8); }) == ErrorKind::InvalidParams);
    }

    // 2x3 point Jacobian: pinhole closed form, fx * (1/z, 0, -x/z^2)
    {
        const double X[3] = {0.2, -0.1, 2.0};
        std::vector<double> uvj, jac; std::vector<uint8_t> stj;
        project_point_jacobian(pin, X, 1, uvj, jac, stj);
        CHECK(stj[0] == 0 && std::fabs(jac[0] - 250.0) < 1e-12 && jac[1] == 0.0 && std::fabs(jac[2] + 500.0 * 0.2 / 4.0) < 1e-12);
        CHECK(jac[3] == 0.0 && std::fabs(jac[4] - 250.0) < 1e-12 && std::fabs(jac[5] - 500.0 * 0.1 / 4.0) < 1e-12);
    }

    std::printf("HOST_TEST_OK mean_px=%.7f lm_iters=%d\n", mean, res.iterations);
    return 0;
}
