// The callers on either side of the LM in the converter:
//   util::sample_points            -> grid, unproject, order-preserving compaction
//   <Model>::linear_estimation     -> normal equations of the linear initialisers
//   util::compute_reprojection_error -> error statistics incl. exact median (radix select)
// Compiled with -fmad=false (same arithmetic as the reference for everything that feeds a
// mask or a kept/dropped decision).
#include <math.h>
#include <stdlib.h>

#include <new>

#include "acm_models.cuh"
#include "acm_reduce.cuh"

// =======================================================================================
// compute_reprojection_error (reference src/util/error_metrics.rs:62-121)
// =======================================================================================
// pass 1: e_i = |project(X_i) - uv_i| for Ok projections (trait project, i.e. with the image
// bounds test of Pinhole/RadTan), NaN otherwise; sums: count, sum e, sum e^2; max e, max -e.
template <int M>
__global__ void __launch_bounds__(256) reproj_pass1_kernel(const __grid_constant__ CamParams c, const double* __restrict__ X,
                                                           const double* __restrict__ Y, const double* __restrict__ Z,
                                                           const double* __restrict__ U, const double* __restrict__ V,
                                                           double* __restrict__ E, size_t n, double* partials, double* out,
                                                           unsigned int* ticket) {
    double acc[5] = {0.0, 0.0, 0.0, -INFINITY, -INFINITY};
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double pu, pv;
        int st = CamModel<M>::template project<true>(c, X[i], Y[i], Z[i], pu, pv);
        double e = acm_nan();
        if (st == ACM_POINT_OK) {
            double dx = pu - U[i], dy = pv - V[i];
            e = sqrt(dx * dx + dy * dy);
            acc[0] += 1.0; acc[1] += e; acc[2] += e * e;
            acc[3] = fmax(acc[3], e); acc[4] = fmax(acc[4], -e);
        }
        E[i] = e;
    }
    GridReduce<3, 2, 0>::run(acc, partials, out, ticket);
}

// pass 2: sum (e - mean)^2
__global__ void __launch_bounds__(256) reproj_pass2_kernel(const double* __restrict__ E, size_t n, double mean, double* partials, double* out,
                                                           unsigned int* ticket) {
    double acc[1] = {0.0};
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double e = E[i];
        if (!isnan(e)) { double d = e - mean; acc[0] += d * d; }
    }
    GridReduce<1, 0, 0>::run(acc, partials, out, ticket);
}

// Radix select over the bit patterns of the non-negative errors (monotone as unsigned ints).
struct SelectState {
    unsigned long long prefix;  // high bits fixed so far
    unsigned long long rank;    // rank still to find inside the prefix class
    unsigned long long hist[256];  // 64-bit so that the counts of all ranks can be summed in place
    int shift;                  // current digit position (56, 48, ..., 0)
};

__global__ void select_init_kernel(SelectState* s, unsigned long long rank) {
    if (threadIdx.x < 256) s->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s->prefix = 0ULL; s->rank = rank; s->shift = 56; }
}

__global__ void __launch_bounds__(256) select_hist_kernel(const double* __restrict__ E, size_t n, SelectState* s) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int shift = s->shift;
    const unsigned long long prefix = s->prefix;
    const unsigned long long mask = (shift == 56) ? 0ULL : (~0ULL << (shift + 8));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double e = E[i];
        if (isnan(e)) continue;
        unsigned long long k = (unsigned long long)__double_as_longlong(e);
        if ((k & mask) == prefix) atomicAdd(&sh[(k >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&s->hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__global__ void select_pick_kernel(SelectState* s) {
    if (threadIdx.x != 0) return;
    unsigned long long r = s->rank, cum = 0;
    int d = 0;
    for (; d < 256; ++d) {
        unsigned long long h = s->hist[d];
        if (cum + h > r) break;
        cum += h;
    }
    if (d > 255) d = 255;
    s->prefix |= ((unsigned long long)d) << s->shift;
    s->rank = r - cum;
    s->shift -= 8;
    for (int i = 0; i < 256; ++i) s->hist[i] = 0;
}

static int32_t select_kth(acm_ctx* ctx, const double* d_E, size_t n, unsigned long long rank, SelectState* d_state, double* h_out) {
    select_init_kernel<<<1, 256, 0, ctx->stream>>>(d_state, rank);
    ACM_CHECK_LAUNCH(ctx);
    int grid = grid_for(ctx, n, 256, 8);
    for (int pass = 0; pass < 8; ++pass) {
        select_hist_kernel<<<grid, 256, 0, ctx->stream>>>(d_E, n, d_state);
        ACM_CHECK_LAUNCH(ctx);
        if (ctx->n_ranks > 1) {  // histogram of the whole set: every rank then picks the same digit
            int32_t rc = acm_allreduce_sum_u64(ctx, d_state->hist, 256);
            if (rc) return rc;
        }
        select_pick_kernel<<<1, 32, 0, ctx->stream>>>(d_state);
        ACM_CHECK_LAUNCH(ctx);
    }
    ACM_CUDA(ctx, cudaMemcpyAsync(h_out, &d_state->prefix, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return ACM_OK;
}

extern "C" int32_t acm_reprojection_error(acm_ctx* ctx, const acm_camera* cam, const acm_points* xyz, const acm_points* uv,
                                          acm_projection_error* out) {
    if (!ctx || !out) return ACM_ERR_INVALID_ARG;
    acm_bind(ctx);
    ACM_REQUIRE(ctx, cam && xyz && uv, "reprojection_error: null argument");
    ACM_REQUIRE(ctx, xyz->dim == 3 && uv->dim == 2 && xyz->n == uv->n, "reprojection_error: shape mismatch");
    ACM_REQUIRE(ctx, xyz->dtype == ACM_F64 && uv->dtype == ACM_F64, "reprojection_error: f64 buffers required");
    memset(out, 0, sizeof(*out));
    const size_t n = xyz->n;
    // an empty shard still takes part in the collectives of a multi-rank call
    if (n == 0 && ctx->n_ranks == 1) return acm_fail(ctx, ACM_ERR_ZERO_PROJECTION_POINTS, "No valid projections");
    CamParams c;
    int32_t rc = acm_make_cam_params(ctx, cam, &c);
    if (rc) return rc;
    rc = acm_ensure_partials(ctx, (size_t)ctx->sm_count * 32 * 64);
    if (rc) return rc;
    // temporaries live in the context's grow-only arena (a cudaMalloc/cudaFree pair per call costs
    // more than the kernels at 10 M points)
    const size_t e_bytes = ((n * sizeof(double) + 255) / 256) * 256;
    rc = acm_ensure_scratch(ctx, e_bytes + sizeof(SelectState));
    if (rc) return rc;
    double* d_E = static_cast<double*>(ctx->d_scratch);
    SelectState* d_sel = reinterpret_cast<SelectState*>(static_cast<char*>(ctx->d_scratch) + e_bytes);
    int grid = grid_for(ctx, n, 256, 4);
    double* h = ctx->h_reduce;
    auto cleanup = [&]() { cudaStreamSynchronize(ctx->stream); };
#define UTIL_TRY(expr) do { int32_t _rc = (expr); if (_rc) { cleanup(); return _rc; } } while (0)
    auto launch1 = [&]() -> int32_t {
        ACM_DISPATCH_MODEL(cam->model, (reproj_pass1_kernel<M><<<grid, 256, 0, ctx->stream>>>(
            c, comp<double>(xyz, 0), comp<double>(xyz, 1), comp<double>(xyz, 2), comp<double>(uv, 0), comp<double>(uv, 1), d_E, n,
            ctx->d_partials, ctx->d_reduce, ctx->d_ticket)))
        ACM_CHECK_LAUNCH(ctx);
        return ACM_OK;
    };
    UTIL_TRY(launch1());
    // with a communicator attached the point buffers are this rank's shard and the statistics are
    // those of the whole set: plain sums are added in rank order, max / min combined, the radix
    // select works on the all-reduced histogram
    const int R = ctx->n_ranks;
    UTIL_TRY(acm_rank_gather_to_host(ctx, 5));
    double cnt = h[0], sum = h[1], sumsq = h[2], mx = h[3], mn = -h[4];
    for (int r = 1; r < R; ++r) {
        const double* v = h + 5 * r;
        cnt += v[0]; sum += v[1]; sumsq += v[2]; mx = fmax(mx, v[3]); mn = fmin(mn, -v[4]);
    }
    if (cnt == 0.0) { cleanup(); return acm_fail(ctx, ACM_ERR_ZERO_PROJECTION_POINTS, "No valid projections"); }
    const double mean = sum / cnt;
    reproj_pass2_kernel<<<grid, 256, 0, ctx->stream>>>(d_E, n, mean, ctx->d_partials, ctx->d_reduce, ctx->d_ticket);
    ctx->launches++;
    UTIL_TRY(acm_rank_gather_to_host(ctx, 1));
    double ssd = h[0];
    for (int r = 1; r < R; ++r) ssd += h[r];
    const unsigned long long m = (unsigned long long)cnt;
    UTIL_TRY(select_kth(ctx, d_E, n, m / 2, d_sel, h + 16));
    if (m % 2 == 0) UTIL_TRY(select_kth(ctx, d_E, n, m / 2 - 1, d_sel, h + 17));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cleanup(); return acm_fail(ctx, ACM_ERR_CUDA, "reprojection_error: %s", cudaGetErrorString(e)); }
#undef UTIL_TRY
    out->count = m;
    out->mean = mean;
    out->stddev = sqrt(ssd / cnt);
    out->rmse = sqrt(sumsq / cnt);
    out->min = mn; out->max = mx;
    out->median = (m % 2 == 0) ? (h[17] + h[16]) / 2.0 : h[16];
    cleanup();
    return ACM_OK;
}

// =======================================================================================
// sample_points (reference src/util/point_sampling.rs:46-120)
// =======================================================================================
template <int M>
__global__ void __launch_bounds__(256) sample_unproject_kernel(const __grid_constant__ CamParams c, int ncx, size_t first, size_t total,
                                                               double cell_w, double cell_h, double* __restrict__ RX,
                                                               double* __restrict__ RY, double* __restrict__ RZ, uint8_t* __restrict__ keep,
                                                               unsigned int* __restrict__ block_counts) {
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;  // local cell; `first + idx` is its row-major grid index
    int k = 0;
    if (idx < total) {
        const size_t i = (first + idx) / (size_t)ncx, j = (first + idx) % (size_t)ncx;
        const double x = ((double)j + 0.5) * cell_w, y = ((double)i + 0.5) * cell_h;
        double rx, ry, rz;
        int st = CamModel<M>::template unproject<true>(c, x, y, rx, ry, rz);  // IEEE tails: the solver's inputs stay bit-identical
        k = (st == ACM_POINT_OK && rz > 0.0) ? 1 : 0;
        RX[idx] = rx; RY[idx] = ry; RZ[idx] = rz;
        keep[idx] = (uint8_t)k;
    }
    int cnt = __syncthreads_count(k);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (unsigned int)cnt;
}

// exclusive scan of the per-block counts by one block (sequential over tiles of 1024)
__global__ void __launch_bounds__(1024) scan_block_counts_kernel(unsigned int* __restrict__ counts, size_t nblk, unsigned long long* __restrict__ offsets,
                                                                 unsigned long long* __restrict__ total) {
    __shared__ unsigned long long sh[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < nblk; base += 1024) {
        size_t i = base + threadIdx.x;
        unsigned long long v = (i < nblk) ? counts[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            unsigned long long t = (threadIdx.x >= (unsigned)o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblk) offsets[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) sample_scatter_kernel(int ncx, size_t first, size_t total, double cell_w, double cell_h, const double* __restrict__ RX,
                                                             const double* __restrict__ RY, const double* __restrict__ RZ,
                                                             const uint8_t* __restrict__ keep, const unsigned long long* __restrict__ offsets,
                                                             double* __restrict__ OU, double* __restrict__ OV, double* __restrict__ OX,
                                                             double* __restrict__ OY, double* __restrict__ OZ) {
    __shared__ unsigned int warp_cnt[8];
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    const int k = (idx < total) ? keep[idx] : 0;
    const unsigned int ballot = __ballot_sync(0xffffffffu, k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    if (k) {
        const size_t pos = (size_t)offsets[blockIdx.x] + before + __popc(ballot & ((1u << lane) - 1u));
        const size_t i = (first + idx) / (size_t)ncx, j = (first + idx) % (size_t)ncx;
        OU[pos] = ((double)j + 0.5) * cell_w; OV[pos] = ((double)i + 0.5) * cell_h;
        OX[pos] = RX[idx]; OY[pos] = RY[idx]; OZ[pos] = RZ[idx];
    }
}

extern "C" int32_t acm_sample_points(acm_ctx* ctx, const acm_camera* cam, size_t n_requested, acm_points** uv_out, acm_points** xyz_out,
                                     size_t* n_kept) {
    return acm_sample_points_shard(ctx, cam, n_requested, 0, 1, uv_out, xyz_out, n_kept);
}

extern "C" int32_t acm_sample_points_shard(acm_ctx* ctx, const acm_camera* cam, size_t n_requested, int32_t shard, int32_t n_shards,
                                           acm_points** uv_out, acm_points** xyz_out, size_t* n_kept) {
    if (!ctx || !uv_out || !xyz_out || !n_kept) return ACM_ERR_INVALID_ARG;
    acm_bind(ctx);
    *uv_out = nullptr; *xyz_out = nullptr; *n_kept = 0;
    ACM_REQUIRE(ctx, cam && cam->width > 0 && cam->height > 0, "sample_points: camera resolution must be set");
    ACM_REQUIRE(ctx, n_shards >= 1 && shard >= 0 && shard < n_shards, "sample_points: shard index out of range");
    CamParams c;
    int32_t rc = acm_make_cam_params(ctx, cam, &c);
    if (rc) return rc;
    // point_sampling.rs:56-62 -- f64::round is half-away-from-zero like C round()
    const double width = (double)cam->width, height = (double)cam->height;
    const int ncx = (int)round(sqrt((double)n_requested * (width / height)));
    const int ncy = (int)round(sqrt((double)n_requested * (height / width)));
    const size_t cells = (size_t)((long long)ncx * (long long)ncy);
    // shard s owns the contiguous row-major cell range [s*cells/S, (s+1)*cells/S): concatenating the
    // shards in order gives exactly the single-shard output
    const size_t first = (size_t)(((unsigned __int128)cells * (unsigned)shard) / (unsigned)n_shards);
    const size_t total = (size_t)(((unsigned __int128)cells * (unsigned)(shard + 1)) / (unsigned)n_shards) - first;
    const double cell_w = width / (double)ncx, cell_h = height / (double)ncy;
    acm_points *uvp = nullptr, *xyzp = nullptr;
    rc = acm_points_create(ctx, 2, total, ACM_F64, &uvp);
    if (!rc) rc = acm_points_create(ctx, 3, total, ACM_F64, &xyzp);
    const size_t nblk = (total + 255) / 256;
    // temporaries (rays before compaction, keep flags, block counts / offsets) in the context's arena
    auto up256 = [](size_t b) { return ((b + 255) / 256) * 256; };
    const size_t ray_bytes = up256(total * sizeof(double)), keep_bytes = up256(total), cnt_bytes = up256(nblk * sizeof(unsigned int)),
                 off_bytes = up256((nblk + 1) * sizeof(unsigned long long));
    if (!rc) rc = acm_ensure_scratch(ctx, 3 * ray_bytes + keep_bytes + cnt_bytes + off_bytes);
    unsigned long long kept = 0;
    if (!rc && total > 0) {
        char* base = static_cast<char*>(ctx->d_scratch);
        double* RX = reinterpret_cast<double*>(base);
        double* RY = reinterpret_cast<double*>(base + ray_bytes);
        double* RZ = reinterpret_cast<double*>(base + 2 * ray_bytes);
        uint8_t* d_keep = reinterpret_cast<uint8_t*>(base + 3 * ray_bytes);
        unsigned int* d_cnt = reinterpret_cast<unsigned int*>(base + 3 * ray_bytes + keep_bytes);
        unsigned long long* d_off = reinterpret_cast<unsigned long long*>(base + 3 * ray_bytes + keep_bytes + cnt_bytes);
        auto run = [&]() -> int32_t {
            ACM_DISPATCH_MODEL(cam->model, (sample_unproject_kernel<M><<<(unsigned)nblk, 256, 0, ctx->stream>>>(
                c, ncx, first, total, cell_w, cell_h, RX, RY, RZ, d_keep, d_cnt)))
            ACM_CHECK_LAUNCH(ctx);
            scan_block_counts_kernel<<<1, 1024, 0, ctx->stream>>>(d_cnt, nblk, d_off, d_off + nblk);
            ACM_CHECK_LAUNCH(ctx);
            sample_scatter_kernel<<<(unsigned)nblk, 256, 0, ctx->stream>>>(ncx, first, total, cell_w, cell_h, RX, RY, RZ, d_keep, d_off,
                                                                          comp<double>(uvp, 0), comp<double>(uvp, 1), comp<double>(xyzp, 0),
                                                                          comp<double>(xyzp, 1), comp<double>(xyzp, 2));
            ACM_CHECK_LAUNCH(ctx);
            ACM_CUDA(ctx, cudaMemcpyAsync(ctx->h_reduce, d_off + nblk, sizeof(kept), cudaMemcpyDeviceToHost, ctx->stream));
            ACM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            memcpy(&kept, ctx->h_reduce, sizeof(kept));
            return ACM_OK;
        };
        rc = run();
    }
    if (rc) { acm_points_destroy(ctx, uvp); acm_points_destroy(ctx, xyzp); return rc; }
    uvp->n = (size_t)kept; xyzp->n = (size_t)kept;  // capacity stays `total`
    *uv_out = uvp; *xyz_out = xyzp; *n_kept = (size_t)kept;
    return ACM_OK;
}

// =======================================================================================
// linear_estimation
// =======================================================================================
// --- UCM / EUCM / Double Sphere: 2N x 1 system, a = (u-cx)(d-z), b = fx*x - (u-cx)z ----------
// (double_sphere.rs:242-258, ucm.rs:213-231, eucm.rs:239-257)
__global__ void __launch_bounds__(256) linest_unified_kernel(double fx, double fy, double cx, double cy, const double* __restrict__ X,
                                                             const double* __restrict__ Y, const double* __restrict__ Z,
                                                             const double* __restrict__ U, const double* __restrict__ V, size_t n,
                                                             double* partials, double* out, unsigned int* ticket) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // dd(sum a*a), dd(sum a*b)
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double x = X[i], y = Y[i], z = Z[i], u = U[i], v = V[i];
        double d = sqrt(x * x + y * y + z * z);
        double u_cx = u - cx, v_cy = v - cy;
        double a0 = u_cx * (d - z), a1 = v_cy * (d - z);
        double b0 = (fx * x) - (u_cx * z), b1 = (fy * y) - (v_cy * z);
        dd_add_prod(acc[0], acc[1], a0, a0); dd_add_prod(acc[0], acc[1], a1, a1);
        dd_add_prod(acc[2], acc[3], a0, b0); dd_add_prod(acc[2], acc[3], a1, b1);
    }
    GridReduce<0, 0, 2>::run(acc, partials, out, ticket);
}

// --- Kannala-Brandt: rows [t^3,t^5,t^7,t^9] twice per point (kannala_brandt.rs:193-259) --------
// Gram entries depend on j+k only: S_m = sum t^(2m+6)... but to reproduce the rounded matrix
// entries of the reference we multiply the *rounded* powers pairwise (10 products).
__global__ void __launch_bounds__(256) linest_kb_kernel(double fx, double fy, double cx, double cy, const double* __restrict__ X,
                                                        const double* __restrict__ Y, const double* __restrict__ Z,
                                                        const double* __restrict__ U, const double* __restrict__ V, size_t n,
                                                        double* partials, double* out, unsigned int* ticket) {
    // pairs: G (10 upper-tri entries, each already doubled for the two identical rows), c (4); + 1 plain flag slot
    double acc[1 + 28];
#pragma unroll
    for (int i = 0; i < 29; ++i) acc[i] = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double xw = X[i], yw = Y[i], zw = Z[i], u = U[i], v = V[i];
        if (zw <= ACM_EPS) continue;
        double rw = sqrt(xw * xw + yw * yw);
        double th = acm_atan2_q1(rw, zw);
        double t2 = th * th, t3 = t2 * th, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
        double x_r = (rw < ACM_EPS) ? 0.0 : xw / rw, y_r = (rw < ACM_EPS) ? 0.0 : yw / rw;
        if ((fabs(fx * x_r) < ACM_EPS && fabs(x_r) > ACM_EPS) || (fabs(fy * y_r) < ACM_EPS && fabs(y_r) > ACM_EPS)) acc[0] += 1.0;
        double bx, by;
        if (fabs(x_r) > ACM_EPS) bx = (u - cx) / (fx * x_r) - th; else bx = (fabs(u - cx) < ACM_EPS) ? -th : 0.0;
        if (fabs(y_r) > ACM_EPS) by = (v - cy) / (fy * y_r) - th; else by = (fabs(v - cy) < ACM_EPS) ? -th : 0.0;
        const double t[4] = {t3, t5, t7, t9};
        int q = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = j; k < 4; ++k) {
                dd_add_prod(acc[1 + 2 * q], acc[2 + 2 * q], t[j], t[k]);
                dd_add_prod(acc[1 + 2 * q], acc[2 + 2 * q], t[j], t[k]);
                ++q;
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dd_add_prod(acc[21 + 2 * j], acc[22 + 2 * j], t[j], bx);
            dd_add_prod(acc[21 + 2 * j], acc[22 + 2 * j], t[j], by);
        }
    }
    GridReduce<1, 0, 14>::run(acc, partials, out, ticket);
}

// --- RadTan: rows fx*xn*[r2,r4,r6], fy*yn*[..] (rad_tan.rs:179-210) ------------------------------
__global__ void __launch_bounds__(256) linest_radtan_kernel(double fx, double fy, double cx, double cy, const double* __restrict__ X,
                                                            const double* __restrict__ Y, const double* __restrict__ Z,
                                                            const double* __restrict__ U, const double* __restrict__ V, size_t n,
                                                            double* partials, double* out, unsigned int* ticket) {
    double acc[18];  // G upper-tri (6 pairs), c (3 pairs)
#pragma unroll
    for (int i = 0; i < 18; ++i) acc[i] = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double x = X[i], y = Y[i], z = Z[i], u = U[i], v = V[i];
        double xn = x / z, yn = y / z;
        double r2 = xn * xn + yn * yn, r4 = r2 * r2, r6 = r4 * r2;
        double uu = fx * xn + cx, vu = fy * yn + cy;
        const double a0[3] = {fx * xn * r2, fx * xn * r4, fx * xn * r6};
        const double a1[3] = {fy * yn * r2, fy * yn * r4, fy * yn * r6};
        const double b0 = u - uu, b1 = v - vu;
        int q = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = j; k < 3; ++k) {
                dd_add_prod(acc[2 * q], acc[2 * q + 1], a0[j], a0[k]);
                dd_add_prod(acc[2 * q], acc[2 * q + 1], a1[j], a1[k]);
                ++q;
            }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            dd_add_prod(acc[12 + 2 * j], acc[13 + 2 * j], a0[j], b0);
            dd_add_prod(acc[12 + 2 * j], acc[13 + 2 * j], a1[j], b1);
        }
    }
    GridReduce<0, 0, 9>::run(acc, partials, out, ticket);
}

// --- FOV: grid search over w = i/100, i in [10, 300) (fov.rs:175-228) ----------------------------
// Two stages for large inputs (290 x N atan2 evaluations in f64 cost 20 ms at N = 10 M):
//   1. `linest_fov_prefilter_kernel`: every candidate in float with cheap branch-free math; its mean
//      errors, together with a rounding bound, shortlist the candidates that can still be the arg-min;
//   2. `linest_fov_exact_kernel`: the reference arithmetic in f64 on the shortlist (or on all 290
//      candidates for small inputs / whenever stage 1 saw a non-finite value).
// Both carry C candidates per thread so that the points are swept 290 / C times, not 290 times.
// block_out[(slot * gridDim.x + block) * 3 + {0: error sum, 1: finite count, 2: sum |u-cx| + |v-cy|}].
constexpr int FOV_NV = 3;

// exact stage: `cand` (nullptr = identity) lists indices iw into the w grid; blockIdx.y = group of C
// consecutive list entries (the tail repeats the last one and is not stored)
template <int C>
__global__ void __launch_bounds__(256) linest_fov_exact_kernel(double fx, double fy, double cx, double cy, const double* __restrict__ tan_half,
                                                               const int* __restrict__ cand, int n_cand, const double* __restrict__ X,
                                                               const double* __restrict__ Y, const double* __restrict__ Z,
                                                               const double* __restrict__ U, const double* __restrict__ V, size_t n,
                                                               double* __restrict__ block_out) {
    double t[C], w[C], sum[C], cnt[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int slot = min((int)blockIdx.y * C + k, n_cand - 1);
        const int iw = cand ? cand[slot] : slot;
        w[k] = (double)(iw + 10) / 100.0;
        t[k] = tan_half[iw];
        sum[k] = cnt[k] = 0.0;
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = X[i], y = Y[i], z = Z[i], u = U[i], v = V[i];
        const double r2 = x * x + y * y, r = sqrt(r2);
#pragma unroll
        for (int k = 0; k < C; ++k) {
            double a = (z > 0.0 && r >= 0.0) ? acm_atan2_q1(2.0 * t[k] * r, z) : atan2(2.0 * t[k] * r, z);  // the grid search has no z guard
            double rd = (r2 < ACM_SQRT_EPS) ? 2.0 * t[k] / w[k] : a / (r * w[k]);
            double mx = x * rd, my = y * rd;
            double up = fx * mx + cx, vp = fy * my + cy;
            double du = up - u, dv = vp - v;
            double e = sqrt(du * du + dv * dv);
            if (isfinite(e)) { sum[k] += e; cnt[k] += 1.0; }
        }
    }
    __shared__ double sm[2][C][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        double s_ = sum[k], c_ = cnt[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s_ += __shfl_down_sync(0xffffffffu, s_, o); c_ += __shfl_down_sync(0xffffffffu, c_, o); }
        if (lane == 0) { sm[0][k][warp] = s_; sm[1][k][warp] = c_; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * C) {
        const int which = threadIdx.x / C, k = threadIdx.x % C;
        const int slot = blockIdx.y * C + k;
        if (slot < n_cand) {
            double acc = 0.0;
            for (int q = 0; q < 8; ++q) acc += sm[which][k][q];
            block_out[((size_t)slot * gridDim.x + blockIdx.x) * FOV_NV + which] = acc;
        }
    }
}

// float pre-filter: candidates blockIdx.y * C .. + C of the full grid.  Non-finite errors are added
// unconditionally, so they surface as a non-finite sum (the host then falls back to the exact search).
// Per evaluation: branch-free atan2 (degree-7 polynomial in q^2 on [0, 1], 1.7e-7 absolute), MUFU
// reciprocal / rsqrt; everything that does not depend on the candidate is hoisted.  Float partial sums
// are flushed into doubles every 16 points, which bounds the accumulation error by 8 ulp.
template <int C>
__global__ void __launch_bounds__(256, 2) linest_fov_prefilter_kernel(double fx_, double fy_, double cx_, double cy_,
                                                                      const double* __restrict__ tan_half, const double* __restrict__ X,
                                                                      const double* __restrict__ Y, const double* __restrict__ Z,
                                                                      const double* __restrict__ U, const double* __restrict__ V, size_t n,
                                                                      double* __restrict__ block_out) {
    float T2[C], inv_w[C], sum[C];
    double dsum[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int iw = blockIdx.y * C + k;
        T2[k] = (float)(2.0 * tan_half[iw]);
        inv_w[k] = (float)(100.0 / (double)(iw + 10));
        sum[k] = 0.0f; dsum[k] = 0.0;
    }
    float mag = 0.0f;
    double dmag = 0.0, dcnt = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int since_flush = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double xd = X[i], yd = Y[i], zd = Z[i], ud = U[i], vd = V[i];
        const float x = (float)xd, y = (float)yd, z = (float)zd;
        const float cxu = (float)(cx_ - ud), cyv = (float)(cy_ - vd);  // centred in f64 first
        const float r2 = x * x + y * y;
        const bool small = r2 < (float)ACM_SQRT_EPS;
        const float r = sqrtf(r2);
        const float inv_r = small ? 1.0f : __fdividef(1.0f, r);
        const float gx = (float)fx_ * x * inv_r, gy = (float)fy_ * y * inv_r;
        const float az = fabsf(z);
        mag += fabsf(cxu) + fabsf(cyv);
        dcnt += 1.0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const float yv = T2[k] * r;                      // >= 0
            const float mx = fmaxf(az, yv), mn = fminf(az, yv);
            const float q = __fdividef(mn, mx);               // NaN for 0/0 -> reported as non-finite
            const float s2 = q * q;
            float p = -4.668773307e-03f;
            p = fmaf(p, s2, 2.416618952e-02f);
            p = fmaf(p, s2, -5.936710079e-02f);
            p = fmaf(p, s2, 9.906096896e-02f);
            p = fmaf(p, s2, -1.401658504e-01f);
            p = fmaf(p, s2, 1.996923539e-01f);
            p = fmaf(p, s2, -3.333195972e-01f);
            p = fmaf(p, s2, 9.999998978e-01f);
            float a = p * q;
            a = (yv > az) ? 1.5707963267948966f - a : a;
            a = (z < 0.0f) ? 3.14159265358979f - a : a;
            const float ak = (small ? T2[k] : a) * inv_w[k];   // r2 < sqrt(EPS): rd = 2 t / w (fov.rs:205)
            const float du = fmaf(gx, ak, cxu), dv = fmaf(gy, ak, cyv);
            const float d2 = fmaf(du, du, dv * dv);
            const float e = d2 > 0.0f ? d2 * rsqrtf(d2) : d2;  // 0 stays 0, NaN stays NaN, inf -> NaN
            sum[k] += e;
        }
        if (++since_flush == 16) {
            since_flush = 0;
#pragma unroll
            for (int k = 0; k < C; ++k) { dsum[k] += (double)sum[k]; sum[k] = 0.0f; }
            dmag += (double)mag; mag = 0.0f;
        }
    }
    __shared__ double sm[C + 2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < C + 2; ++k) {
        double s_ = k < C ? dsum[k < C ? k : 0] + (double)sum[k < C ? k : 0] : (k == C ? dcnt : dmag + (double)mag);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s_ += __shfl_down_sync(0xffffffffu, s_, o);
        if (lane == 0) sm[k][warp] = s_;
    }
    __syncthreads();
    if (threadIdx.x < 3 * C) {
        const int which = threadIdx.x / C, k = threadIdx.x % C;
        const int row = which == 0 ? k : (which == 1 ? C : C + 1);
        double acc = 0.0;
        for (int q = 0; q < 8; ++q) acc += sm[row][q];
        block_out[((size_t)(blockIdx.y * C + k) * gridDim.x + blockIdx.x) * FOV_NV + which] = acc;
    }
}

// sums the per-block (sum, count) pairs of every candidate in block order
__global__ void linest_fov_sum_kernel(const double* __restrict__ block_out, int gx, int nvals, double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;  // v = 3*slot + {0,1,2}
    if (v >= nvals) return;
    const int slot = v / 3, which = v % 3;
    double s = 0.0;
    for (int b = 0; b < gx; ++b) s += block_out[((size_t)slot * gx + b) * 3 + which];
    out[v] = s;
}

// ---- host-side double-double arithmetic for the tiny dense solves -------------------------------
namespace {
struct dd { double hi, lo; };
inline dd dd_make(double h, double l = 0.0) { return {h, l}; }
inline dd two_sum(double a, double b) { volatile double s = a + b; volatile double bb = s - a; volatile double e = (a - (s - bb)) + (b - bb); return {s, e}; }
inline dd quick_two_sum(double a, double b) { volatile double s = a + b; volatile double e = b - (s - a); return {s, e}; }
inline dd two_prod(double a, double b) { double p = a * b; double e = fma(a, b, -p); return {p, e}; }
inline dd operator+(dd a, dd b) { dd s = two_sum(a.hi, b.hi); dd t = two_sum(a.lo, b.lo); s.lo += t.hi; s = quick_two_sum(s.hi, s.lo); s.lo += t.lo; return quick_two_sum(s.hi, s.lo); }
inline dd operator-(dd a) { return {-a.hi, -a.lo}; }
inline dd operator-(dd a, dd b) { return a + (-b); }
inline dd operator*(dd a, dd b) { dd p = two_prod(a.hi, b.hi); p.lo += a.hi * b.lo + a.lo * b.hi; return quick_two_sum(p.hi, p.lo); }
inline dd operator/(dd a, dd b) {
    double q1 = a.hi / b.hi; dd r = a - b * dd_make(q1);
    double q2 = r.hi / b.hi; r = r - b * dd_make(q2);
    double q3 = r.hi / b.hi;
    dd q = quick_two_sum(q1, q2); return q + dd_make(q3);
}
inline dd dd_sqrt(dd a) {
    if (a.hi <= 0.0) return dd_make(0.0);
    double x = 1.0 / sqrt(a.hi); double ax = a.hi * x;
    dd d = a - two_prod(ax, ax);
    return two_sum(ax, d.hi * (x * 0.5));
}
inline double dd_abs(dd a) { return fabs(a.hi); }

// Symmetric eigen-decomposition (cyclic Jacobi) of a k x k matrix in double-double, then the
// truncated pseudo-inverse solve that nalgebra's SVD::solve(b, eps) performs on A with
// A^T A = G: singular values sigma_i = sqrt(lambda_i), components with sigma_i <= eps dropped.
void solve_gram_svd(int k, dd* G /* k*k */, const dd* c, double eps, double* x) {
    dd V[16];
    for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) V[i * k + j] = dd_make(i == j ? 1.0 : 0.0);
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) (i == j ? diag : off) += G[i * k + j].hi * G[i * k + j].hi;
        if (off <= 1e-60 * diag || off == 0.0) break;  // |off| / |diag| below double-double resolution
        for (int p = 0; p < k - 1; ++p) for (int q = p + 1; q < k; ++q) {
            dd apq = G[p * k + q];
            // already zero to double-double precision relative to the diagonal: rotating would only divide by noise
            if (apq.hi == 0.0 || fabs(apq.hi) <= 1e-31 * sqrt(fabs(G[p * k + p].hi * G[q * k + q].hi))) continue;
            dd theta = (G[q * k + q] - G[p * k + p]) / (apq + apq);
            dd t;
            dd root = dd_sqrt(theta * theta + dd_make(1.0));
            if (theta.hi >= 0.0) t = dd_make(1.0) / (theta + root); else t = dd_make(-1.0) / (root - theta);
            dd cs = dd_make(1.0) / dd_sqrt(t * t + dd_make(1.0)), sn = t * cs;
            for (int r = 0; r < k; ++r) {  // columns p,q
                dd grp = G[r * k + p], grq = G[r * k + q];
                G[r * k + p] = cs * grp - sn * grq; G[r * k + q] = sn * grp + cs * grq;
            }
            for (int r = 0; r < k; ++r) {  // rows p,q
                dd gpr = G[p * k + r], gqr = G[q * k + r];
                G[p * k + r] = cs * gpr - sn * gqr; G[q * k + r] = sn * gpr + cs * gqr;
            }
            for (int r = 0; r < k; ++r) {
                dd vrp = V[r * k + p], vrq = V[r * k + q];
                V[r * k + p] = cs * vrp - sn * vrq; V[r * k + q] = sn * vrp + cs * vrq;
            }
        }
    }
    dd xs[4] = {dd_make(0), dd_make(0), dd_make(0), dd_make(0)};
    for (int i = 0; i < k; ++i) {
        dd lam = G[i * k + i];
        if (!(lam.hi > 0.0)) continue;
        double sigma = sqrt(lam.hi);
        if (!(sigma > eps)) continue;
        dd proj = dd_make(0.0);
        for (int r = 0; r < k; ++r) proj = proj + V[r * k + i] * c[r];
        dd coef = proj / lam;
        for (int r = 0; r < k; ++r) xs[r] = xs[r] + V[r * k + i] * coef;
    }
    for (int r = 0; r < k; ++r) x[r] = xs[r].hi + xs[r].lo;
    (void)dd_abs;
}
}  // namespace

extern "C" int32_t acm_linear_estimation(acm_ctx* ctx, acm_camera* cam, const acm_points* xyz, const acm_points* uv) {
    ACM_ENTER(ctx);
    ACM_REQUIRE(ctx, cam && xyz && uv, "linear_estimation: null argument");
    ACM_REQUIRE(ctx, xyz->dim == 3 && uv->dim == 2, "linear_estimation: xyz must have dim 3 and uv dim 2");
    ACM_REQUIRE(ctx, xyz->dtype == ACM_F64 && uv->dtype == ACM_F64, "linear_estimation: f64 buffers required");
    if (xyz->n != uv->n) return acm_fail(ctx, ACM_ERR_INVALID_PARAMS, "Number of 2D and 3D points must match");
    ACM_REQUIRE(ctx, cam->model >= 0 && cam->model <= 6 && cam->n_params == acm_n_params(cam->model), "linear_estimation: bad camera block");
    const size_t n = xyz->n;
    const double fx = cam->params[0], fy = cam->params[1], cx = cam->params[2], cy = cam->params[3];
    int32_t rc = acm_ensure_partials(ctx, (size_t)ctx->sm_count * 32 * 64);
    if (rc) return rc;
    const double *X = comp<double>(xyz, 0), *Y = comp<double>(xyz, 1), *Z = comp<double>(xyz, 2), *U = comp<double>(uv, 0), *V = comp<double>(uv, 1);
    int grid = grid_for(ctx, n, 256, 2);
    double* h = ctx->h_reduce;
    // Cross-rank combination keeps the double-double pairs exact: every rank writes its vector into
    // its own slot of a zero-padded buffer, the all-reduce then acts as an all-gather, and the
    // host adds the slots in rank order (n_plain leading plain sums, then (hi, lo) pairs).
    auto fetch = [&](int count, int n_plain) -> int32_t {
        const int R = ctx->n_ranks;
        int32_t r2 = acm_rank_gather_to_host(ctx, count);
        if (r2) return r2;
        for (int r = 1; r < R; ++r) {
            const double* v = h + (size_t)r * count;
            for (int i = 0; i < n_plain; ++i) h[i] += v[i];
            for (int i = n_plain; i + 1 < count; i += 2) {
                dd a = dd_make(h[i], h[i + 1]) + dd_make(v[i], v[i + 1]);
                h[i] = a.hi; h[i + 1] = a.lo;
            }
        }
        return ACM_OK;
    };
    // The minimum-point tests of the reference look at the whole correspondence set; with a communicator attached `n` is
    // this rank's shard (possibly empty), so the ranks first agree on the global count -- every rank then takes the same
    // branch and none is left waiting in a collective.
    double n_global = (double)n;
    if (ctx->n_ranks > 1) {
        if (cam->model == ACM_MODEL_PINHOLE) return acm_fail(ctx, ACM_ERR_INVALID_ARG, "the pinhole model has no linear_estimation");
        h[0] = (double)n;
        ACM_CUDA(ctx, cudaMemcpyAsync(ctx->d_reduce, h, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        rc = acm_allreduce_sum_f64(ctx, ctx->d_reduce, 1);
        if (rc) return rc;
        ACM_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_reduce, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ACM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        n_global = h[0];
    }
    switch (cam->model) {
        case ACM_MODEL_UCM: case ACM_MODEL_EUCM: case ACM_MODEL_DOUBLE_SPHERE: {
            if (cam->model == ACM_MODEL_EUCM) {
                if (n_global < 1) return acm_fail(ctx, ACM_ERR_INVALID_PARAMS, "Need at least 1 point for EUCM linear estimation");
                cam->params[5] = 1.0;  // eucm.rs:236
            }
            double alpha = 0.0;
            {
                linest_unified_kernel<<<grid, 256, 0, ctx->stream>>>(fx, fy, cx, cy, X, Y, Z, U, V, n, ctx->d_partials, ctx->d_reduce, ctx->d_ticket);
                ACM_CHECK_LAUNCH(ctx);
                rc = fetch(4, 0);
                if (rc) return rc;
                dd G[1] = {dd_make(h[0], h[1])}, c[1] = {dd_make(h[2], h[3])};
                solve_gram_svd(1, G, c, 1e-10, &alpha);
            }
            if (cam->model == ACM_MODEL_DOUBLE_SPHERE) {  // double_sphere.rs:269-286
                cam->params[5] = 0.0;
                if (alpha <= 0.0) alpha = 0.01; else if (alpha > 1.0) alpha = 1.0;
            } else if (cam->model == ACM_MODEL_UCM) {     // ucm.rs:246-250
                if (alpha <= 0.0) alpha = 0.01;
            } else {                                      // eucm.rs:273-280
                if (alpha <= 0.0) alpha = 0.01; else if (alpha > 2.0) alpha = 2.0;
            }
            cam->params[4] = alpha;
            break;
        }
        case ACM_MODEL_KANNALA_BRANDT: {
            if (n_global < 4) return acm_fail(ctx, ACM_ERR_INVALID_PARAMS, "Not enough points for linear estimation (need at least 4)");
            linest_kb_kernel<<<grid, 256, 0, ctx->stream>>>(fx, fy, cx, cy, X, Y, Z, U, V, n, ctx->d_partials, ctx->d_reduce, ctx->d_ticket);
            ACM_CHECK_LAUNCH(ctx);
            rc = fetch(29, 1);
            if (rc) return rc;
            if (h[0] > 0.0) return acm_fail(ctx, ACM_ERR_NUMERICAL, "fx * x_r is zero in linear estimation");
            dd G[16], c[4];
            int q = 0;
            for (int j = 0; j < 4; ++j) for (int k = j; k < 4; ++k) { G[j * 4 + k] = G[k * 4 + j] = dd_make(h[1 + 2 * q], h[2 + 2 * q]); ++q; }
            for (int j = 0; j < 4; ++j) c[j] = dd_make(h[21 + 2 * j], h[22 + 2 * j]);
            double kk[4];
            solve_gram_svd(4, G, c, 2.220446049250313e-16, kk);
            for (int j = 0; j < 4; ++j) cam->params[4 + j] = kk[j];
            break;
        }
        case ACM_MODEL_RADTAN: {
            if (n_global < 3) return acm_fail(ctx, ACM_ERR_INVALID_PARAMS, "Need at least 3 points for RadTan linear estimation");
            linest_radtan_kernel<<<grid, 256, 0, ctx->stream>>>(fx, fy, cx, cy, X, Y, Z, U, V, n, ctx->d_partials, ctx->d_reduce, ctx->d_ticket);
            ACM_CHECK_LAUNCH(ctx);
            rc = fetch(18, 0);
            if (rc) return rc;
            dd G[9], c[3];
            int q = 0;
            for (int j = 0; j < 3; ++j) for (int k = j; k < 3; ++k) { G[j * 3 + k] = G[k * 3 + j] = dd_make(h[2 * q], h[2 * q + 1]); ++q; }
            for (int j = 0; j < 3; ++j) c[j] = dd_make(h[12 + 2 * j], h[13 + 2 * j]);
            double kk[3];
            solve_gram_svd(3, G, c, 1e-10, kk);
            cam->params[4] = kk[0]; cam->params[5] = kk[1]; cam->params[6] = 0.0; cam->params[7] = 0.0; cam->params[8] = kk[2];
            return ACM_OK;  // rad_tan.rs:221-233: no validate_params
        }
        case ACM_MODEL_FOV: {
            if (n_global < 2) return acm_fail(ctx, ACM_ERR_INVALID_PARAMS, "Need at least 2 point correspondences for linear estimation");
            constexpr int NW = 290, C32 = 10, C64 = 2, MAX_SHORT = 32;
            static_assert(NW % C32 == 0, "the pre-filter has no tail group");
            int gx = grid_for(ctx, n, 256, 1);
            if (gx > 64) gx = 64;
            const int gx_short = grid_for(ctx, n, 256, 4);  // the shortlist has few groups: more blocks along the points
            // arena: tan table | candidate list | per-block (sum, count, magnitude) of every candidate
            const size_t tan_bytes = 2560, cand_bytes = 1280;
            const size_t blk_doubles = (size_t)FOV_NV * ((size_t)NW * gx > (size_t)MAX_SHORT * gx_short ? (size_t)NW * gx : (size_t)MAX_SHORT * gx_short);
            rc = acm_ensure_scratch(ctx, tan_bytes + cand_bytes + blk_doubles * sizeof(double));
            if (rc) return rc;
            double* d_tan = static_cast<double*>(ctx->d_scratch);
            int* d_cand = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + tan_bytes);
            double* d_blk = reinterpret_cast<double*>(static_cast<char*>(ctx->d_scratch) + tan_bytes + cand_bytes);
            double* h_tan = h + 1024;  // pinned (2048 doubles); h[0 .. 870) receives the reduced sums
            int* h_cand = reinterpret_cast<int*>(h + 1024 + 320);
            for (int i = 0; i < NW; ++i) h_tan[i] = tan(((double)(i + 10) / 100.0) / 2.0);  // host libm, as the reference
            ACM_CUDA(ctx, cudaMemcpyAsync(d_tan, h_tan, NW * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            // reduced (error sum, finite count, magnitude sum) of `n_cand` candidates -> h[3*slot + {0,1,2}]
            auto evaluate = [&](bool f32, const int* cand_dev, int n_cand) -> int32_t {
                int g = gx;
                if (f32) {
                    linest_fov_prefilter_kernel<C32><<<dim3(gx, NW / C32), 256, 0, ctx->stream>>>(fx, fy, cx, cy, d_tan, X, Y, Z, U, V, n, d_blk);
                } else {
                    if (n_cand <= MAX_SHORT) g = gx_short;
                    linest_fov_exact_kernel<C64><<<dim3(g, (n_cand + C64 - 1) / C64), 256, 0, ctx->stream>>>(
                        fx, fy, cx, cy, d_tan, cand_dev, n_cand, X, Y, Z, U, V, n, d_blk);
                }
                ACM_CHECK_LAUNCH(ctx);
                linest_fov_sum_kernel<<<(3 * n_cand + 255) / 256, 256, 0, ctx->stream>>>(d_blk, g, 3 * n_cand, ctx->d_reduce);
                ACM_CHECK_LAUNCH(ctx);
                int32_t r2 = acm_allreduce_sum_f64(ctx, ctx->d_reduce, 3 * (size_t)n_cand);
                if (r2) return r2;
                ACM_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_reduce, 3 * (size_t)n_cand * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                ACM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                return ACM_OK;
            };
            // Stage 1 (large inputs only): float pre-filter over all 290 candidates.  A candidate stays
            // on the shortlist when its float mean minus its rounding bound does not exceed the
            // smallest (float mean + bound).  Any non-finite float error, an empty or an oversized
            // shortlist falls back to the exact search over every candidate.  The decision uses
            // all-reduced sums, so every rank takes the same path.
            int n_short = 0;
            // the pre-filter is only worth its extra launches on large inputs; the global size decides so that all ranks agree
            const double n_total = n_global;
            if (n_total >= 200000.0) {
                rc = evaluate(true, nullptr, NW);
                if (rc) return rc;
                bool usable = true;
                double lo[NW], best_hi = INFINITY;
                for (int i = 0; i < NW && usable; ++i) {
                    const double s_ = h[3 * i], cnt = h[3 * i + 1], mg = h[3 * i + 2];
                    if (cnt != n_total || !isfinite(s_) || !isfinite(mg)) { usable = false; break; }
                    // float error of one evaluation: <= ~17 ulp of (|u-cx| + |v-cy|) + ~20 ulp of e, the flushed
                    // accumulation adds <= 8 ulp of the sum; 64 ulp of both leaves a factor 2-3 of margin
                    const double mean = s_ / cnt, bound = 64.0 * 5.9604644775390625e-08 * (mg / cnt + mean);
                    lo[i] = mean - bound;
                    if (mean + bound < best_hi) best_hi = mean + bound;
                }
                if (usable) {
                    for (int i = 0; i < NW; ++i)
                        if (lo[i] <= best_hi) { if (n_short < MAX_SHORT) h_cand[n_short] = i; ++n_short; }
                    if (n_short == 0 || n_short > MAX_SHORT) n_short = 0;
                }
            }
            double best_w = 1.0, best_err = INFINITY;
            if (n_short > 0) {
                ACM_CUDA(ctx, cudaMemcpyAsync(d_cand, h_cand, n_short * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                int short_list[MAX_SHORT];
                memcpy(short_list, h_cand, n_short * sizeof(int));
                rc = evaluate(false, d_cand, n_short);
                if (rc) return rc;
                for (int k = 0; k < n_short; ++k) {  // ascending candidate order, first strict minimum wins (fov.rs:218-221)
                    const double s_ = h[3 * k], cnt = h[3 * k + 1];
                    if (cnt > 0.0) { double avg = s_ / cnt; if (avg < best_err) { best_err = avg; best_w = (double)(short_list[k] + 10) / 100.0; } }
                }
            } else {
                rc = evaluate(false, nullptr, NW);
                if (rc) return rc;
                for (int i = 0; i < NW; ++i) {
                    const double s_ = h[3 * i], cnt = h[3 * i + 1];
                    if (cnt > 0.0) { double avg = s_ / cnt; if (avg < best_err) { best_err = avg; best_w = (double)(i + 10) / 100.0; } }
                }
            }
            double w = best_w;
            if (w <= 2.220446049250313e-16) w = 0.01; else if (w > 3.0) w = 3.0;
            cam->params[4] = w;
            break;
        }
        default:
            return acm_fail(ctx, ACM_ERR_INVALID_ARG, "the pinhole model has no linear_estimation");
    }
    char msg[128];
    rc = acm_validate_params(cam, msg, sizeof(msg));  // every model but RadTan validates afterwards
    if (rc) return acm_fail(ctx, rc, "%s", msg);
    return ACM_OK;
}
