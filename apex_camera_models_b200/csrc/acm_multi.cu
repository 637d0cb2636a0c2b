// Multi-GPU from ONE host thread (SURVEY.md section 8b "Threading" / "Minimum C symbols").
//
// The reference's converter `main` (bin/camera_converter.rs:127-343) is single-threaded and has no
// launcher, so the drop-in offers a form that needs neither: acm_comm_init_all binds one context per
// GPU of this process into a group -- peer access between the devices, the NVLink exchange buffers of
// the fused kernel mapped directly (no IPC: the contexts share an address space), one NCCL communicator
// per context (ncclCommInitAll) for the small host-side gathers -- and starts one worker thread per
// context.  From then on every context IS a rank of the one-process-per-GPU form; the *_multi entry
// points hand each rank's share to its worker, block until all are done and return rank 0's result
// (every rank computes bit-identical results, and that is checked).
#include "acm_internal.cuh"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct acm_group {
    std::vector<acm_ctx*> ctxs;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::function<int32_t(int)> job;
    unsigned long long generation = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int32_t> rc;
};

static void worker_main(acm_group* g, int i) {
    cudaSetDevice(g->ctxs[i]->device);
    unsigned long long seen = 0;
    for (;;) {
        std::function<int32_t(int)> job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_work.wait(lk, [&] { return g->stop || g->generation != seen; });
            if (g->stop) return;
            seen = g->generation;
            job = g->job;
        }
        const int32_t r = job(i);
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->rc[i] = r;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// Run job(i) for every rank on its worker thread; returns the first non-zero status in rank order.
static int32_t run_on_group(acm_group* g, std::function<int32_t(int)> job) {
    std::unique_lock<std::mutex> lk(g->mu);
    g->job = std::move(job);
    g->pending = (int)g->ctxs.size();
    for (auto& r : g->rc) r = ACM_OK;
    ++g->generation;
    g->cv_work.notify_all();
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
    for (size_t i = 0; i < g->rc.size(); ++i) if (g->rc[i]) return g->rc[i];
    return ACM_OK;
}

static int32_t check_group(acm_ctx** ctxs, int32_t n, acm_group** out) {
    if (!ctxs || n < 1 || !ctxs[0]) return ACM_ERR_INVALID_ARG;
    acm_group* g = ctxs[0]->group;
    if (!g || (int32_t)g->ctxs.size() != n) return acm_fail(ctxs[0], ACM_ERR_INVALID_ARG, "these contexts are not a group: call acm_comm_init_all on exactly this array first");
    for (int i = 0; i < n; ++i)
        if (ctxs[i] != g->ctxs[i]) return acm_fail(ctxs[0], ACM_ERR_INVALID_ARG, "context %d differs from the group built by acm_comm_init_all", i);
    *out = g;
    return ACM_OK;
}

// NCCL in-process: ncclCommInitAll hands back one communicator per device (acm_core.cu owns the dlopen'ed API)
int32_t acm_nccl_init_all(acm_ctx** ctxs, int32_t n);

extern "C" int32_t acm_comm_init_all(acm_ctx** ctxs, int32_t n) {
    if (!ctxs || n < 1 || n > ACM_MAX_PEERS) return ACM_ERR_INVALID_ARG;
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return ACM_ERR_INVALID_ARG;
        if (ctxs[i]->group || ctxs[i]->comm || ctxs[i]->peer_n) return acm_fail(ctxs[i], ACM_ERR_INVALID_ARG, "comm_init_all: context %d already belongs to a group / communicator", i);
        for (int j = 0; j < i; ++j)
            if (ctxs[j]->device == ctxs[i]->device) return acm_fail(ctxs[i], ACM_ERR_INVALID_ARG, "comm_init_all: contexts %d and %d share device %d (one context per GPU)", j, i, ctxs[i]->device);
    }
    if (n == 1) return ACM_OK;  // a group of one is the plain single-GPU form
    // peer access between every pair, exchange buffer on every device
    for (int i = 0; i < n; ++i) {
        acm_ctx* c = ctxs[i];
        ACM_CUDA(c, cudaSetDevice(c->device));
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            int can = 0;
            ACM_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, ctxs[j]->device));
            if (!can) return acm_fail(c, ACM_ERR_CUDA, "device %d cannot access device %d as a peer", c->device, ctxs[j]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return acm_fail(c, ACM_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", c->device, ctxs[j]->device, cudaGetErrorString(e));
        }
        if (!c->peer_local) {
            ACM_CUDA(c, cudaMalloc(&c->peer_local, ACM_PEER_BUFFER_BYTES));
        }
        ACM_CUDA(c, cudaMemset(c->peer_local, 0, ACM_PEER_BUFFER_BYTES));
        ACM_CUDA(c, cudaDeviceSynchronize());
    }
    unsigned char* ptrs[ACM_MAX_PEERS];
    for (int i = 0; i < n; ++i) ptrs[i] = ctxs[i]->peer_local;   // one address space: a peer's buffer is its own pointer
    for (int i = 0; i < n; ++i) {
        ACM_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
        int32_t rc = acm_peer_setup_pointers(ctxs[i], n, i, ptrs);
        if (rc) return rc;
    }
    // NCCL for the host-side gathers of the initialisers / statistics (optional: without libnccl those entry points report it)
    int32_t rc = acm_nccl_init_all(ctxs, n);
    if (rc) return rc;
    acm_group* g = new acm_group();
    g->ctxs.assign(ctxs, ctxs + n);
    g->rc.assign(n, ACM_OK);
    for (int i = 0; i < n; ++i) ctxs[i]->group = g;
    for (int i = 0; i < n; ++i) g->workers.emplace_back(worker_main, g, i);
    return ACM_OK;
}

extern "C" int32_t acm_comm_destroy_all(acm_ctx** ctxs, int32_t n) {
    if (!ctxs || n < 1 || !ctxs[0]) return ACM_ERR_INVALID_ARG;
    acm_group* g = ctxs[0]->group;
    if (!g) return ACM_OK;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
        g->cv_work.notify_all();
    }
    for (auto& t : g->workers) t.join();
    for (acm_ctx* c : g->ctxs) {
        c->group = nullptr;
        acm_comm_destroy(c);
        acm_peer_detach(c);
    }
    delete g;
    return ACM_OK;
}

void acm_group_dissolve(acm_ctx* member) {
    acm_group* g = member->group;
    if (g) acm_comm_destroy_all(g->ctxs.data(), (int32_t)g->ctxs.size());
}

extern "C" int32_t acm_linearize_multi(acm_ctx** ctxs, int32_t n, const acm_camera* cam, int32_t residual_kind, acm_points* const* xyz,
                                       acm_points* const* uv, acm_normal_equations* out) {
    if (n == 1 && ctxs && ctxs[0] && !ctxs[0]->group) return acm_linearize(ctxs[0], cam, residual_kind, xyz ? xyz[0] : nullptr, uv ? uv[0] : nullptr, out);
    acm_group* g = nullptr;
    int32_t rc = check_group(ctxs, n, &g);
    if (rc) return rc;
    if (!xyz || !uv || !out) return acm_fail(ctxs[0], ACM_ERR_INVALID_ARG, "linearize_multi: null argument");
    std::vector<acm_normal_equations> res(n);
    rc = run_on_group(g, [&](int i) { return acm_linearize(ctxs[i], cam, residual_kind, xyz[i], uv[i], &res[i]); });
    if (rc) return rc;
    for (int i = 1; i < n; ++i)
        if (memcmp(&res[i], &res[0], sizeof(res[0])) != 0) return acm_fail(ctxs[0], ACM_ERR_NUMERICAL, "linearize_multi: rank %d holds different sums than rank 0", i);
    *out = res[0];
    return ACM_OK;
}

extern "C" int32_t acm_lm_solve_multi(acm_ctx** ctxs, int32_t n, const acm_camera* init, int32_t residual_kind, acm_points* const* xyz,
                                      acm_points* const* uv, const double* lower, const double* upper, const acm_lm_config* cfg,
                                      double* out_params, acm_lm_result* result) {
    if (n == 1 && ctxs && ctxs[0] && !ctxs[0]->group)
        return acm_lm_solve(ctxs[0], init, residual_kind, xyz ? xyz[0] : nullptr, uv ? uv[0] : nullptr, lower, upper, cfg, out_params, result);
    acm_group* g = nullptr;
    int32_t rc = check_group(ctxs, n, &g);
    if (rc) return rc;
    if (!init || !xyz || !uv || !out_params || !result) return acm_fail(ctxs[0], ACM_ERR_INVALID_ARG, "lm_solve_multi: null argument");
    std::vector<acm_lm_result> res(n);
    std::vector<double> par((size_t)n * ACM_MAX_PARAMS, 0.0);
    rc = run_on_group(g, [&](int i) {
        return acm_lm_solve(ctxs[i], init, residual_kind, xyz[i], uv[i], lower, upper, cfg, par.data() + (size_t)i * ACM_MAX_PARAMS, &res[i]);
    });
    if (rc) return rc;
    const int P = init->n_params;
    double wall = 0.0, dev = 0.0;
    for (int i = 0; i < n; ++i) {
        if (memcmp(par.data() + (size_t)i * ACM_MAX_PARAMS, par.data(), P * sizeof(double)) != 0 || res[i].iterations != res[0].iterations ||
            res[i].status != res[0].status)
            return acm_fail(ctxs[0], ACM_ERR_NUMERICAL, "lm_solve_multi: rank %d ended on a different trajectory than rank 0", i);
        wall = res[i].elapsed_ms > wall ? res[i].elapsed_ms : wall;
        dev = res[i].device_ms > dev ? res[i].device_ms : dev;
    }
    memcpy(out_params, par.data(), P * sizeof(double));
    *result = res[0];
    result->elapsed_ms = wall; result->device_ms = dev;
    return ACM_OK;
}

extern "C" int32_t acm_linear_estimation_multi(acm_ctx** ctxs, int32_t n, acm_camera* cam, acm_points* const* xyz, acm_points* const* uv) {
    if (n == 1 && ctxs && ctxs[0] && !ctxs[0]->group) return acm_linear_estimation(ctxs[0], cam, xyz ? xyz[0] : nullptr, uv ? uv[0] : nullptr);
    acm_group* g = nullptr;
    int32_t rc = check_group(ctxs, n, &g);
    if (rc) return rc;
    if (!cam || !xyz || !uv) return acm_fail(ctxs[0], ACM_ERR_INVALID_ARG, "linear_estimation_multi: null argument");
    std::vector<acm_camera> cams(n, *cam);
    rc = run_on_group(g, [&](int i) { return acm_linear_estimation(ctxs[i], &cams[i], xyz[i], uv[i]); });
    if (rc) return rc;
    for (int i = 1; i < n; ++i)
        if (memcmp(cams[i].params, cams[0].params, sizeof(cams[0].params)) != 0) return acm_fail(ctxs[0], ACM_ERR_NUMERICAL, "linear_estimation_multi: rank %d differs from rank 0", i);
    *cam = cams[0];
    return ACM_OK;
}

extern "C" int32_t acm_reprojection_error_multi(acm_ctx** ctxs, int32_t n, const acm_camera* cam, acm_points* const* xyz, acm_points* const* uv,
                                                acm_projection_error* out) {
    if (n == 1 && ctxs && ctxs[0] && !ctxs[0]->group) return acm_reprojection_error(ctxs[0], cam, xyz ? xyz[0] : nullptr, uv ? uv[0] : nullptr, out);
    acm_group* g = nullptr;
    int32_t rc = check_group(ctxs, n, &g);
    if (rc) return rc;
    if (!cam || !xyz || !uv || !out) return acm_fail(ctxs[0], ACM_ERR_INVALID_ARG, "reprojection_error_multi: null argument");
    std::vector<acm_projection_error> res(n);
    rc = run_on_group(g, [&](int i) { return acm_reprojection_error(ctxs[i], cam, xyz[i], uv[i], &res[i]); });
    if (rc) return rc;
    for (int i = 1; i < n; ++i)
        if (memcmp(&res[i], &res[0], sizeof(res[0])) != 0) return acm_fail(ctxs[0], ACM_ERR_NUMERICAL, "reprojection_error_multi: rank %d differs from rank 0", i);
    *out = res[0];
    return ACM_OK;
}

extern "C" int32_t acm_sample_points_multi(acm_ctx** ctxs, int32_t n, const acm_camera* cam, size_t n_requested, acm_points** uv_out,
                                           acm_points** xyz_out, size_t* n_kept) {
    if (!ctxs || n < 1 || !ctxs[0]) return ACM_ERR_INVALID_ARG;
    if (!cam || !uv_out || !xyz_out || !n_kept) return acm_fail(ctxs[0], ACM_ERR_INVALID_ARG, "sample_points_multi: null argument");
    if (n == 1 && !ctxs[0]->group) return acm_sample_points(ctxs[0], cam, n_requested, &uv_out[0], &xyz_out[0], &n_kept[0]);
    acm_group* g = nullptr;
    int32_t rc = check_group(ctxs, n, &g);
    if (rc) return rc;
    return run_on_group(g, [&](int i) { return acm_sample_points_shard(ctxs[i], cam, n_requested, i, n, &uv_out[i], &xyz_out[i], &n_kept[i]); });
}
