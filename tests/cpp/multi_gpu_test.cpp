// Plain C-ABI test of the single-process multi-GPU form (acm_comm_init_all + acm_*_multi) -- what a
// single-threaded, launcher-less host like the reference's converter `main`
// (bin/camera_converter.rs:127-343) binds -- and of the scalar host path (acm_project_host with n = 1,
// the call behind the trait's `project(&p)`).  Needs >= 1 GPU; the multi-GPU part needs >= 2 and prints
// MULTI_GPU_SKIPPED otherwise.  Prints "MULTI_TEST_OK".
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "acm.h"

#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)
#define OK(ctx, call) do { int32_t _rc = (call); if (_rc != ACM_OK) { std::fprintf(stderr, "%s -> %d: %s (%s:%d)\n", #call, (int)_rc, acm_last_error(ctx), __FILE__, __LINE__); std::exit(1); } } while (0)

static const double KB[8] = {190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504,
                             0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202, 0.00020293673591811182};

static acm_camera make(int model, const std::vector<double>& p, uint32_t w, uint32_t h) {
    acm_camera c;
    std::memset(&c, 0, sizeof(c));
    c.model = model; c.width = w; c.height = h; c.n_params = (int32_t)p.size();
    for (size_t i = 0; i < p.size(); ++i) c.params[i] = p[i];
    return c;
}
static bool close_rel(double a, double b, double rtol) { return std::fabs(a - b) <= rtol * std::fmax(std::fabs(a), std::fabs(b)) + 1e-300; }

int main(int argc, char** argv) {
    int ndev_want = argc > 1 ? std::atoi(argv[1]) : 2;
    acm_ctx* c0 = nullptr;
    if (acm_ctx_create(0, nullptr, &c0) != ACM_OK) { std::fprintf(stderr, "%s\n", acm_last_error(nullptr)); return 2; }
    int ndev = 1;   // count the devices through the ABI itself (no CUDA headers needed to build this test)
    for (;; ++ndev) {
        acm_ctx* probe = nullptr;
        if (acm_ctx_create(ndev, nullptr, &probe) != ACM_OK) break;
        acm_ctx_destroy(probe);
    }
    acm_camera kb = make(ACM_MODEL_KANNALA_BRANDT, std::vector<double>(KB, KB + 8), 512, 512);

    // ---- scalar host path: the trait's project(&p) / unproject(&uv) -------------------------------------------
    {
        const double p[3] = {0.1, 0.2, 1.0};
        double uv[2], ray[3];
        uint8_t st = 9;
        OK(c0, acm_project_host(c0, &kb, p, 1, uv, &st));
        CHECK(st == ACM_POINT_OK && std::isfinite(uv[0]));
        OK(c0, acm_unproject_host(c0, &kb, uv, 1, ray, &st));
        const double n = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        CHECK(st == ACM_POINT_OK && std::fabs(ray[0] - p[0] / n) < 1e-9 && std::fabs(ray[2] - p[2] / n) < 1e-9);
        // the batch kernels must give the same bits as the small-batch path
        const int m = 1000;
        std::vector<double> P(3 * m), U1(2 * m), U2(2 * 4096);
        std::vector<uint8_t> S1(m), S2(4096);
        for (int i = 0; i < m; ++i) { P[3 * i] = 0.001 * i - 0.4; P[3 * i + 1] = 0.3 - 0.0007 * i; P[3 * i + 2] = (i % 17 == 0) ? -1.0 : 1.0 + 0.01 * i; }
        OK(c0, acm_project_host(c0, &kb, P.data(), m, U1.data(), S1.data()));           // small-batch path
        std::vector<double> Pbig(3 * 4096, 1.0);
        std::memcpy(Pbig.data(), P.data(), sizeof(double) * 3 * m);
        OK(c0, acm_project_host(c0, &kb, Pbig.data(), 4096, U2.data(), S2.data()));     // staged batch path
        for (int i = 0; i < m; ++i) {
            CHECK(S1[i] == S2[i]);
            CHECK(std::memcmp(&U1[2 * i], &U2[2 * i], 16) == 0);
        }
        const int reps = 2000;
        for (int i = 0; i < 50; ++i) acm_project_host(c0, &kb, p, 1, uv, &st);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) acm_project_host(c0, &kb, p, 1, uv, &st);
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
        std::printf("SCALAR_PROJECT_US %.2f\n", us);
        CHECK(us < 100.0);  // was >= 100 us with two buffer creations, a cudaMalloc / cudaFree and three synchronisations per call
    }

    if (ndev < ndev_want || ndev_want < 2) {
        std::printf("MULTI_GPU_SKIPPED (%d device(s))\nMULTI_TEST_OK\n", ndev);
        acm_ctx_destroy(c0);
        return 0;
    }
    const int G = ndev_want;
    const size_t N = 400000 + 3;   // odd, not divisible by G
    const double cos_max = std::cos(85.0 * M_PI / 180.0);

    // ---- reference: everything on one GPU -----------------------------------------------------------------------
    acm_points *X = nullptr, *U = nullptr;
    OK(c0, acm_points_create(c0, 3, N, ACM_F64, &X));
    OK(c0, acm_points_create(c0, 2, N, ACM_F64, &U));
    OK(c0, acm_synth_points3(c0, 0xACE50004ULL, 0, cos_max, 0, X));
    OK(c0, acm_project(c0, &kb, X, U, nullptr));
    acm_camera ds0 = make(ACM_MODEL_DOUBLE_SPHERE, {KB[0], KB[1], KB[2], KB[3], 0.5, 0.1}, 512, 512);
    acm_camera ds1 = ds0;
    OK(c0, acm_linear_estimation(c0, &ds1, X, U));
    acm_normal_equations ne1;
    OK(c0, acm_linearize(c0, &ds1, ACM_RESIDUAL_ALGEBRAIC, X, U, &ne1));
    const double lo[6] = {1, 1, 0, 0, 1e-6, -5}, hi[6] = {2000, 2000, 2000, 2000, 1, 5};
    double par1[ACM_MAX_PARAMS];
    acm_lm_result r1;
    OK(c0, acm_lm_solve(c0, &ds1, ACM_RESIDUAL_ALGEBRAIC, X, U, lo, hi, nullptr, par1, &r1));
    acm_camera dsf1 = ds1;
    for (int i = 0; i < 6; ++i) dsf1.params[i] = par1[i];
    acm_projection_error pe1;
    OK(c0, acm_reprojection_error(c0, &dsf1, X, U, &pe1));
    acm_points *su1 = nullptr, *sx1 = nullptr;
    size_t kept1 = 0;
    OK(c0, acm_sample_points(c0, &kb, 100000, &su1, &sx1, &kept1));
    std::vector<double> sx1h(3 * kept1);
    OK(c0, acm_points_download_aos_f64(c0, sx1, sx1h.data(), kept1));
    acm_points_destroy(c0, su1); acm_points_destroy(c0, sx1);
    acm_points_destroy(c0, X); acm_points_destroy(c0, U);
    acm_ctx_destroy(c0);

    // ---- the same on G GPUs, driven from this one thread -----------------------------------------------------------
    std::vector<acm_ctx*> ctx(G, nullptr);
    for (int g = 0; g < G; ++g) OK(nullptr, acm_ctx_create(g, nullptr, &ctx[g]));
    OK(ctx[0], acm_comm_init_all(ctx.data(), G));
    CHECK(acm_comm_size(ctx[0]) == G);
    std::vector<acm_points*> Xs(G, nullptr), Us(G, nullptr);
    for (int g = 0; g < G; ++g) {
        const size_t a = N * g / G, b = N * (g + 1) / G;
        OK(ctx[g], acm_points_create(ctx[g], 3, b - a, ACM_F64, &Xs[g]));
        OK(ctx[g], acm_points_create(ctx[g], 2, b - a, ACM_F64, &Us[g]));
        OK(ctx[g], acm_synth_points3(ctx[g], 0xACE50004ULL, a, cos_max, 0, Xs[g]));
        OK(ctx[g], acm_project(ctx[g], &kb, Xs[g], Us[g], nullptr));
        OK(ctx[g], acm_ctx_sync(ctx[g]));
    }
    acm_camera dsG = ds0;
    OK(ctx[0], acm_linear_estimation_multi(ctx.data(), G, &dsG, Xs.data(), Us.data()));
    for (int i = 0; i < 6; ++i) CHECK(close_rel(dsG.params[i], ds1.params[i], 1e-13));
    acm_normal_equations neG;
    OK(ctx[0], acm_linearize_multi(ctx.data(), G, &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs.data(), Us.data(), &neG));
    CHECK(neG.n_valid == ne1.n_valid && close_rel(neG.cost, ne1.cost, 1e-12));
    for (int i = 0; i < 36; ++i) CHECK(close_rel(neG.H[i], ne1.H[i], 1e-11));
    double parG[ACM_MAX_PARAMS];
    acm_lm_result rG;
    for (int rep = 0; rep < 3; ++rep) {   // repeated solves: the exchange counters of the ranks must stay in step
        OK(ctx[0], acm_lm_solve_multi(ctx.data(), G, &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs.data(), Us.data(), lo, hi, nullptr, parG, &rG));
        CHECK(rG.status == r1.status && rG.iterations == r1.iterations && rG.passes == r1.passes && rG.n_valid == r1.n_valid);
        for (int i = 0; i < 6; ++i) CHECK(close_rel(parG[i], par1[i], 1e-11));
    }
    std::printf("LM 1 GPU %.3f ms (device %.3f) | %d GPUs %.3f ms (device %.3f), %d passes\n", r1.elapsed_ms, r1.device_ms, G, rG.elapsed_ms, rG.device_ms, rG.passes);
    acm_camera dsfG = ds1;
    for (int i = 0; i < 6; ++i) dsfG.params[i] = parG[i];
    acm_projection_error peG;
    OK(ctx[0], acm_reprojection_error_multi(ctx.data(), G, &dsfG, Xs.data(), Us.data(), &peG));
    CHECK(peG.count == pe1.count && close_rel(peG.mean, pe1.mean, 1e-9) && close_rel(peG.median, pe1.median, 1e-9) && close_rel(peG.max, pe1.max, 1e-9));
    // sample_points over the group: concatenated shards == the single-GPU output, bit for bit
    std::vector<acm_points*> su(G, nullptr), sx(G, nullptr);
    std::vector<size_t> kept(G, 0);
    OK(ctx[0], acm_sample_points_multi(ctx.data(), G, &kb, 100000, su.data(), sx.data(), kept.data()));
    size_t tot = 0, off = 0;
    for (int g = 0; g < G; ++g) tot += kept[g];
    CHECK(tot == kept1);
    std::vector<double> cat(3 * tot);
    for (int g = 0; g < G; ++g) {
        OK(ctx[g], acm_points_download_aos_f64(ctx[g], sx[g], cat.data() + 3 * off, kept[g]));
        off += kept[g];
        acm_points_destroy(ctx[g], su[g]); acm_points_destroy(ctx[g], sx[g]);
    }
    CHECK(std::memcmp(cat.data(), sx1h.data(), sizeof(double) * 3 * tot) == 0);
    // a group is not a bag of contexts: wrong arrays are rejected, not silently mis-sharded
    std::vector<acm_ctx*> swapped(ctx.rbegin(), ctx.rend());
    CHECK(acm_linearize_multi(swapped.data(), G, &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs.data(), Us.data(), &neG) == ACM_ERR_INVALID_ARG);
    // The minimum-point tests of linear_estimation look at the GLOBAL count (with local counts some ranks
    // returned InvalidParams before the collective and the others hung in it).  5 correspondences over G ranks: shards of
    // 0..3 points, KB needs 4 in total -> every rank succeeds and agrees with one GPU; 3 correspondences -> every rank fails.
    {
        acm_ctx* c1 = nullptr;
        OK(nullptr, acm_ctx_create(0, nullptr, &c1));   // a second, ungrouped context on device 0 for the single-GPU answer
        for (size_t total : {(size_t)5, (size_t)3}) {
            std::vector<double> hx(3 * total), hu(2 * total);
            for (size_t i = 0; i < total; ++i) { hx[3 * i] = 0.1 * (double)(i + 1); hx[3 * i + 1] = -0.07 * (double)(i + 1); hx[3 * i + 2] = 1.0 + 0.1 * (double)i; }
            std::vector<uint8_t> st(total);
            OK(c1, acm_project_host(c1, &kb, hx.data(), total, hu.data(), st.data()));
            acm_points *x1 = nullptr, *u1 = nullptr;
            OK(c1, acm_points_create(c1, 3, total, ACM_F64, &x1)); OK(c1, acm_points_create(c1, 2, total, ACM_F64, &u1));
            OK(c1, acm_points_upload_aos_f64(c1, x1, hx.data(), total)); OK(c1, acm_points_upload_aos_f64(c1, u1, hu.data(), total));
            acm_camera k1 = make(ACM_MODEL_KANNALA_BRANDT, {KB[0], KB[1], KB[2], KB[3], 0, 0, 0, 0}, 512, 512), kG = k1;
            const int32_t rc1 = acm_linear_estimation(c1, &k1, x1, u1);
            std::vector<acm_points*> xs(G, nullptr), us(G, nullptr);
            for (int g = 0; g < G; ++g) {
                const size_t a = total * g / G, b = total * (g + 1) / G;
                OK(ctx[g], acm_points_create(ctx[g], 3, b - a, ACM_F64, &xs[g])); OK(ctx[g], acm_points_create(ctx[g], 2, b - a, ACM_F64, &us[g]));
                if (b > a) { OK(ctx[g], acm_points_upload_aos_f64(ctx[g], xs[g], hx.data() + 3 * a, b - a)); OK(ctx[g], acm_points_upload_aos_f64(ctx[g], us[g], hu.data() + 2 * a, b - a)); }
                OK(ctx[g], acm_ctx_sync(ctx[g]));
            }
            const int32_t rcG = acm_linear_estimation_multi(ctx.data(), G, &kG, xs.data(), us.data());
            CHECK(rcG == rc1);
            CHECK(total >= 4 ? rc1 == ACM_OK : rc1 == ACM_ERR_INVALID_PARAMS);
            if (rc1 == ACM_OK) for (int i = 4; i < 8; ++i) CHECK(close_rel(kG.params[i], k1.params[i], 1e-9) || std::fabs(kG.params[i] - k1.params[i]) < 1e-12);
            for (int g = 0; g < G; ++g) { acm_points_destroy(ctx[g], xs[g]); acm_points_destroy(ctx[g], us[g]); }
            acm_points_destroy(c1, x1); acm_points_destroy(c1, u1);
        }
        acm_ctx_destroy(c1);
    }
    // A lost peer must end in a hard error on EVERY rank, not in ranks that drift apart: rank 0 solves alone,
    // nobody delivers the other ranks' sums, its kernel gives up after ~2 s, raises the sticky abort flag in every rank's
    // exchange buffer and the call returns ACM_ERR_PEER.  The next group call fails on all ranks at once (rank 0 on the
    // host, the others as soon as their kernels see the flag); re-creating the group clears it.
    {
        acm_lm_result rr;
        double pp[ACM_MAX_PARAMS];
        const auto t0 = std::chrono::steady_clock::now();
        CHECK(acm_lm_solve(ctx[0], &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs[0], Us[0], lo, hi, nullptr, pp, &rr) == ACM_ERR_PEER);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("PEER_TIMEOUT_S %.2f (%s)\n", sec, acm_last_error(ctx[0]));
        CHECK(sec > 0.5 && sec < 20.0);
        const auto t1 = std::chrono::steady_clock::now();
        CHECK(acm_lm_solve_multi(ctx.data(), G, &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs.data(), Us.data(), lo, hi, nullptr, pp, &rr) == ACM_ERR_PEER);
        CHECK(std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count() < 1.0);   // fails fast: the flag is sticky
        CHECK(acm_linearize_multi(ctx.data(), G, &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs.data(), Us.data(), &neG) == ACM_ERR_PEER);
        OK(ctx[0], acm_comm_destroy_all(ctx.data(), G));
        OK(ctx[0], acm_comm_init_all(ctx.data(), G));
        OK(ctx[0], acm_lm_solve_multi(ctx.data(), G, &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs.data(), Us.data(), lo, hi, nullptr, pp, &rr));
        CHECK(rr.status == r1.status && rr.iterations == r1.iterations);
        for (int i = 0; i < 6; ++i) CHECK(close_rel(pp[i], par1[i], 1e-11));
    }
    // Ranks whose call sequences drifted apart must not add the sums of different kernels: rank 0 linearises with the
    // algebraic residual while rank 1 linearises with the pixel residual under the same exchange number.  The call signature
    // travels in the cell tags, so neither accepts the other's cells; both give up after the time-out with ACM_ERR_PEER.
    {
        int32_t rc0 = 0, rc1 = 0;
        acm_normal_equations ne0, ne1;
        std::thread ta([&] { rc0 = acm_linearize(ctx[0], &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs[0], Us[0], &ne0); });
        std::thread tb([&] { rc1 = acm_linearize(ctx[1], &ds1, ACM_RESIDUAL_PIXEL, Xs[1], Us[1], &ne1); });
        ta.join(); tb.join();
        std::printf("MISMATCH_RC %d %d\n", rc0, rc1);
        CHECK(rc0 == ACM_ERR_PEER && rc1 == ACM_ERR_PEER);
        OK(ctx[0], acm_comm_destroy_all(ctx.data(), G));
        OK(ctx[0], acm_comm_init_all(ctx.data(), G));
        acm_lm_result rr;
        double pp[ACM_MAX_PARAMS];
        OK(ctx[0], acm_lm_solve_multi(ctx.data(), G, &ds1, ACM_RESIDUAL_ALGEBRAIC, Xs.data(), Us.data(), lo, hi, nullptr, pp, &rr));
        for (int i = 0; i < 6; ++i) CHECK(close_rel(pp[i], par1[i], 1e-11));
    }
    for (int g = 0; g < G; ++g) { acm_points_destroy(ctx[g], Xs[g]); acm_points_destroy(ctx[g], Us[g]); }
    OK(ctx[0], acm_comm_destroy_all(ctx.data(), G));
    for (int g = 0; g < G; ++g) acm_ctx_destroy(ctx[g]);
    std::printf("MULTI_TEST_OK\n");
    return 0;
}
