// Image-quality diagnostics of the converter, on RGB8 images in HBM:
//   util::calculate_psnr                   reference src/util/image_quality.rs:45-89
//   util::calculate_ssim (+ rgb_to_grayscale)                                  :108-210
//   the radius-2 disc every drawing routine uses (create_projection_image :338-373,
//   create_combined_projection_image[_on_reference] :389-505, model_projection_visualization :553-616)
//   util::compute_image_quality_metrics                                        :254-324
// Compiled with -fmad=false: the per-window SSIM term is evaluated with the reference's separately
// rounded operations in the reference's order (bit-identical per pixel; only the order of the final
// sum over pixels differs, deterministically), PSNR is integer arithmetic up to the last division.
#include <math.h>
#include <stdlib.h>

#include "acm_internal.cuh"
#include "acm_math.cuh"
#include "acm_reduce.cuh"

// ---------------------------------------------------------------------------------------
// PSNR: sum of squared channel differences and the number of channels of the non-black pixels.
// Integers (u64 atomics): exact and order-independent, like the reference's f64 sums below 2^53.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void psnr_pixel(uint32_t r1, uint32_t g1, uint32_t b1, uint32_t r2, uint32_t g2, uint32_t b2,
                                           unsigned long long& sse, unsigned long long& valid) {
    if ((r1 | g1 | b1 | r2 | g2 | b2) != 0u) {   // image_quality.rs:60-66: skip pixels that are black in both images
        const int dr = (int)r1 - (int)r2, dg = (int)g1 - (int)g2, db = (int)b1 - (int)b2;
        sse += (unsigned long long)(dr * dr + dg * dg + db * db);
        valid += 3ull;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256) psnr_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, size_t npix,
                                                   unsigned long long* __restrict__ out /* [2] */) {
    unsigned long long sse = 0ull, valid = 0ull;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        // 4 pixels = 12 bytes = three 32-bit words per image (both bases 4-byte aligned)
        const size_t ngroups = npix / 4;
        const uint32_t* A4 = reinterpret_cast<const uint32_t*>(A);
        const uint32_t* B4 = reinterpret_cast<const uint32_t*>(B);
        for (size_t g = tid; g < ngroups; g += stride) {
            const uint32_t a0 = __ldcs(A4 + 3 * g), a1 = __ldcs(A4 + 3 * g + 1), a2 = __ldcs(A4 + 3 * g + 2);
            const uint32_t b0 = __ldcs(B4 + 3 * g), b1 = __ldcs(B4 + 3 * g + 1), b2 = __ldcs(B4 + 3 * g + 2);
            psnr_pixel(a0 & 255u, (a0 >> 8) & 255u, (a0 >> 16) & 255u, b0 & 255u, (b0 >> 8) & 255u, (b0 >> 16) & 255u, sse, valid);
            psnr_pixel(a0 >> 24, a1 & 255u, (a1 >> 8) & 255u, b0 >> 24, b1 & 255u, (b1 >> 8) & 255u, sse, valid);
            psnr_pixel((a1 >> 16) & 255u, a1 >> 24, a2 & 255u, (b1 >> 16) & 255u, b1 >> 24, b2 & 255u, sse, valid);
            psnr_pixel((a2 >> 8) & 255u, (a2 >> 16) & 255u, a2 >> 24, (b2 >> 8) & 255u, (b2 >> 16) & 255u, b2 >> 24, sse, valid);
        }
        for (size_t p = ngroups * 4 + tid; p < npix; p += stride)
            psnr_pixel(A[3 * p], A[3 * p + 1], A[3 * p + 2], B[3 * p], B[3 * p + 1], B[3 * p + 2], sse, valid);
    } else {
        for (size_t p = tid; p < npix; p += stride)
            psnr_pixel(A[3 * p], A[3 * p + 1], A[3 * p + 2], B[3 * p], B[3 * p + 1], B[3 * p + 2], sse, valid);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sse += __shfl_down_sync(0xffffffffu, sse, o);
        valid += __shfl_down_sync(0xffffffffu, valid, o);
    }
    if ((threadIdx.x & 31) == 0 && (sse | valid)) { atomicAdd(out, sse); atomicAdd(out + 1, valid); }
}

// ---------------------------------------------------------------------------------------
// SSIM: 3x3 windows over the interior of the truncated-luma grey images.
// A block owns 32 x 8 window centres per tile and stages the (34 x 10) grey values of both images
// in shared memory (the grey conversion happens once per tile pixel, not nine times).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t luma_u8(uint32_t rgb) {
    // image_quality.rs:203-204: (0.299 r + 0.587 g + 0.114 b) as u8 -- truncating, saturating
    const double g = 0.299 * (double)(rgb & 255u) + 0.587 * (double)((rgb >> 8) & 255u) + 0.114 * (double)((rgb >> 16) & 255u);
    return (uint8_t)min(255, max(0, __double2int_rz(g)));
}

#define SSIM_TX 32
#define SSIM_TY 8
#define SSIM_TILE_PX ((SSIM_TY + 2) * (SSIM_TX + 2))
__global__ void __launch_bounds__(256) ssim_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, uint32_t W, uint32_t H,
                                                   double c1, double c2, double* partials, double* out, unsigned int* ticket) {
    __shared__ uint8_t t1[SSIM_TY + 2][SSIM_TX + 2 + 2], t2[SSIM_TY + 2][SSIM_TX + 2 + 2];
    double acc[2] = {0.0, 0.0};  // sum of the window terms, number of windows
    const uint32_t iw = W - 2, ih = H - 2;                     // interior size (W, H >= 3)
    const uint32_t tiles_x = (iw + SSIM_TX - 1) / SSIM_TX, tiles_y = (ih + SSIM_TY - 1) / SSIM_TY;
    const size_t ntiles = (size_t)tiles_x * tiles_y;
    const int lx = threadIdx.x & (SSIM_TX - 1), ly = threadIdx.x / SSIM_TX;
    // the RGB bytes of the NEXT tile are fetched into registers while the windows of the current one are evaluated
    // (ncu: the load and compute phases of a block used to alternate, FP64 pipe 39 % busy); a thread owns tile pixels
    // k = tid and, for tid < SSIM_TILE_PX - 256, k = tid + 256
    auto fetch = [&](size_t tile, int k, uint32_t& pa, uint32_t& pb) {
        pa = pb = 0u;
        if (tile < ntiles && k < SSIM_TILE_PX) {
            const int ty = k / (SSIM_TX + 2), tx = k - ty * (SSIM_TX + 2);
            const uint32_t x = (uint32_t)(tile % tiles_x) * SSIM_TX + tx, y = (uint32_t)(tile / tiles_x) * SSIM_TY + ty;
            if (x < W && y < H) {
                const size_t off = 3 * ((size_t)y * W + x);
                pa = (uint32_t)A[off] | ((uint32_t)A[off + 1] << 8) | ((uint32_t)A[off + 2] << 16);
                pb = (uint32_t)B[off] | ((uint32_t)B[off + 1] << 8) | ((uint32_t)B[off + 2] << 16);
            }
        }
    };
    uint32_t pa0, pb0, pa1, pb1;
    fetch(blockIdx.x, threadIdx.x, pa0, pb0);
    fetch(blockIdx.x, threadIdx.x + 256, pa1, pb1);
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t x0 = (uint32_t)(tile % tiles_x) * SSIM_TX, y0 = (uint32_t)(tile / tiles_x) * SSIM_TY;  // tile origin in image coords (halo corner)
        __syncthreads();
        {   // pixels outside the image were fetched as 0 and are never part of a window
            const int k0 = threadIdx.x, k1 = threadIdx.x + 256;
            t1[k0 / (SSIM_TX + 2)][k0 % (SSIM_TX + 2)] = luma_u8(pa0); t2[k0 / (SSIM_TX + 2)][k0 % (SSIM_TX + 2)] = luma_u8(pb0);
            if (k1 < SSIM_TILE_PX) { t1[k1 / (SSIM_TX + 2)][k1 % (SSIM_TX + 2)] = luma_u8(pa1); t2[k1 / (SSIM_TX + 2)][k1 % (SSIM_TX + 2)] = luma_u8(pb1); }
        }
        __syncthreads();
        fetch(tile + gridDim.x, threadIdx.x, pa0, pb0);
        fetch(tile + gridDim.x, threadIdx.x + 256, pa1, pb1);
        const uint32_t x = x0 + 1 + lx, y = y0 + 1 + ly;   // window centre
        if (x < W - 1 && y < H - 1) {
            // the window sums run over integers <= 9 * 255: exact in any arithmetic, so they are taken in integer
            // registers (the reference's sequential f64 sum gives the same value); each grey value is converted once
            int is1 = 0, is2 = 0;
            double v1[9], v2[9];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int a = t1[ly + dy][lx + dx], b = t2[ly + dy][lx + dx];
                    is1 += a; is2 += b;
                    v1[dy * 3 + dx] = (double)a; v2[dy * 3 + dx] = (double)b;
                }
            const double ls1 = (double)is1, ls2 = (double)is2;
            // x / 9.0 through the correctly rounded reciprocal (bit-identical, acm_div_by); x / 8.0 == x * 0.125 exactly
            const double mu1 = acm_div_by(ls1, 9.0, 1.0 / 9.0), mu2 = acm_div_by(ls2, 9.0, 1.0 / 9.0);
            double s1 = 0.0, s2 = 0.0, s12 = 0.0;
#pragma unroll
            for (int k = 0; k < 9; ++k) {   // dy outer, dx inner: the reference's order
                s1 += (v1[k] - mu1) * (v1[k] - mu1);
                s2 += (v2[k] - mu2) * (v2[k] - mu2);
                s12 += (v1[k] - mu1) * (v2[k] - mu2);
            }
            s1 *= 0.125; s2 *= 0.125; s12 *= 0.125;
            const double numerator = (2.0 * mu1 * mu2 + c1) * (2.0 * s12 + c2);
            const double denominator = (mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2);
            if (denominator > 0.0) { acc[0] += numerator / denominator; acc[1] += 1.0; }
        }
    }
    GridReduce<2, 0, 0>::run(acc, partials, out, ticket);
}

// ---------------------------------------------------------------------------------------
// drawing: one thread per point, the 13 pixels of the radius-2 disc, clipped to the image.
// All points of a call carry one colour, so concurrent writers of a pixel store the same bytes.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long round_i32(double v) {
    // `x.round() as i32` (image_quality.rs:353-354): half away from zero, saturating, NaN -> 0
    // (the hardware conversion saturates, but does not map NaN to 0 on sm_100: handled explicitly)
    return v != v ? 0ll : (long long)__double2int_rz(round(v));
}

__global__ void __launch_bounds__(256) draw_points_kernel(const double* __restrict__ U, const double* __restrict__ V,
                                                          const uint8_t* __restrict__ keep, size_t n, uint8_t r, uint8_t g, uint8_t b,
                                                          uint8_t* __restrict__ img, uint32_t W, uint32_t H) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (keep && !keep[i]) continue;
        const long long cx = round_i32(U[i]), cy = round_i32(V[i]);
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx)
                if (dx * dx + dy * dy <= 4) {
                    const long long x = cx + dx, y = cy + dy;
                    if (x >= 0 && x < (long long)W && y >= 0 && y < (long long)H) {
                        uint8_t* p = img + 3 * ((size_t)y * W + (size_t)x);
                        p[0] = r; p[1] = g; p[2] = b;
                    }
                }
    }
}

// keep[i] = both projections Ok and the OUTPUT projection inside [0,W) x [0,H)  (image_quality.rs:283-302)
__global__ void __launch_bounds__(256) iq_keep_kernel(const uint8_t* __restrict__ st_in, const uint8_t* __restrict__ st_out,
                                                      const double* __restrict__ UO, const double* __restrict__ VO, size_t n, double W,
                                                      double H, uint8_t* __restrict__ keep, unsigned long long* __restrict__ count) {
    unsigned int mine = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double u = UO[i], v = VO[i];
        const bool k = st_in[i] == ACM_POINT_OK && st_out[i] == ACM_POINT_OK && u >= 0.0 && u < W && v >= 0.0 && v < H;
        keep[i] = k ? 1 : 0;
        mine += k ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, (unsigned long long)mine);
}

// display image = what drawing green input discs, then magenta output discs over the reference (or black)
// leaves behind: magenta where an output disc covers the pixel, else green where an input disc does
__global__ void __launch_bounds__(256) iq_compose_kernel(const uint8_t* __restrict__ img_in, const uint8_t* __restrict__ img_out,
                                                         const uint8_t* __restrict__ reference, uint8_t* __restrict__ combined, size_t npix) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
        uint8_t r = 0, g = 0, b = 0;
        if (reference) { r = reference[3 * p]; g = reference[3 * p + 1]; b = reference[3 * p + 2]; }
        if (img_out[3 * p]) { r = 255; g = 0; b = 255; }
        else if (img_in[3 * p]) { r = 0; g = 255; b = 0; }
        combined[3 * p] = r; combined[3 * p + 1] = g; combined[3 * p + 2] = b;
    }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static int32_t psnr_sums(acm_ctx* ctx, const uint8_t* a, const uint8_t* b, size_t npix, unsigned long long* d_sums, unsigned long long h[2]) {
    ACM_CUDA(ctx, cudaMemsetAsync(d_sums, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (npix > 0) {
        const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 3u) == 0;
        const int grid = grid_for(ctx, npix / 4 + 1, 256, 8);
        if (vec) psnr_kernel<true><<<grid, 256, 0, ctx->stream>>>(a, b, npix, d_sums);
        else psnr_kernel<false><<<grid, 256, 0, ctx->stream>>>(a, b, npix, d_sums);
        ACM_CHECK_LAUNCH(ctx);
    }
    ACM_CUDA(ctx, cudaMemcpyAsync(h, d_sums, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    ACM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACM_OK;
}

static double psnr_from_sums(unsigned long long sse, unsigned long long valid) {
    if (valid == 0) return INFINITY;                 // image_quality.rs:77-79
    volatile double mse = (double)sse / (double)valid;
    if (mse <= 1e-10) return INFINITY;               // :83-84
    volatile double q = 255.0 * 255.0 / mse;
    return 10.0 * log10(q);                           // :86
}

extern "C" int32_t acm_image_psnr(acm_ctx* ctx, const uint8_t* d_img1, const uint8_t* d_img2, uint32_t width, uint32_t height, double* psnr) {
    if (!ctx || !psnr) return ACM_ERR_INVALID_ARG;
    acm_bind(ctx);
    const size_t npix = (size_t)width * height;
    ACM_REQUIRE(ctx, npix == 0 || (d_img1 && d_img2), "image_psnr: null image");
    int32_t rc = acm_ensure_scratch(ctx, 256);
    if (rc) return rc;
    unsigned long long h[2];
    rc = psnr_sums(ctx, d_img1, d_img2, npix, static_cast<unsigned long long*>(ctx->d_scratch), h);
    if (rc) return rc;
    *psnr = psnr_from_sums(h[0], h[1]);
    return ACM_OK;
}

extern "C" int32_t acm_image_ssim(acm_ctx* ctx, const uint8_t* d_img1, const uint8_t* d_img2, uint32_t width, uint32_t height, double* ssim) {
    if (!ctx || !ssim) return ACM_ERR_INVALID_ARG;
    acm_bind(ctx);
    if (width < 3 || height < 3) { *ssim = 1.0; return ACM_OK; }   // no interior window: count == 0 (image_quality.rs:184-188)
    ACM_REQUIRE(ctx, d_img1 && d_img2, "image_ssim: null image");
    int32_t rc = acm_ensure_partials(ctx, (size_t)ctx->sm_count * 32 * 64);
    if (rc) return rc;
    volatile double t1 = 0.01 * 255.0, t2 = 0.03 * 255.0;          // :118-119, powi(2) == x * x
    const double c1 = t1 * t1, c2 = t2 * t2;
    const size_t ntiles = (size_t)((width - 2 + SSIM_TX - 1) / SSIM_TX) * ((height - 2 + SSIM_TY - 1) / SSIM_TY);
    const size_t cap = (size_t)ctx->sm_count * 8;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    ssim_kernel<<<grid, 256, 0, ctx->stream>>>(d_img1, d_img2, width, height, c1, c2, ctx->d_partials, ctx->d_reduce, ctx->d_ticket);
    ACM_CHECK_LAUNCH(ctx);
    ACM_CUDA(ctx, cudaMemcpyAsync(ctx->h_reduce, ctx->d_reduce, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ACM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *ssim = ctx->h_reduce[1] > 0.0 ? ctx->h_reduce[0] / ctx->h_reduce[1] : 1.0;
    return ACM_OK;
}

extern "C" int32_t acm_draw_points_rgb8(acm_ctx* ctx, const acm_points* uv, const uint8_t* d_keep, uint8_t r, uint8_t g, uint8_t b,
                                        uint8_t* d_image, uint32_t width, uint32_t height) {
    ACM_ENTER(ctx);
    ACM_REQUIRE(ctx, uv && d_image, "draw_points: null argument");
    ACM_REQUIRE(ctx, uv->dim == 2 && uv->dtype == ACM_F64, "draw_points: uv must be an f64 buffer of dim 2");
    if (uv->n == 0 || width == 0 || height == 0) return ACM_OK;
    draw_points_kernel<<<grid_for(ctx, uv->n, 256, 8), 256, 0, ctx->stream>>>(comp<double>(uv, 0), comp<double>(uv, 1), d_keep, uv->n, r, g, b,
                                                                              d_image, width, height);
    ACM_CHECK_LAUNCH(ctx);
    return ACM_OK;
}

extern "C" int32_t acm_image_quality_metrics(acm_ctx* ctx, const acm_camera* input_model, const acm_camera* output_model, const acm_points* xyz,
                                             uint32_t width, uint32_t height, const uint8_t* d_reference, uint8_t* d_combined,
                                             acm_image_quality* out) {
    if (!ctx || !out) return ACM_ERR_INVALID_ARG;
    acm_bind(ctx);
    ACM_REQUIRE(ctx, input_model && output_model && xyz, "image_quality_metrics: null argument");
    ACM_REQUIRE(ctx, xyz->dim == 3 && xyz->dtype == ACM_F64, "image_quality_metrics: xyz must be an f64 buffer of dim 3");
    memset(out, 0, sizeof(*out));
    out->psnr = out->ssim = NAN;
    const size_t n = xyz->n, npix = (size_t)width * height, img_bytes = ((npix * 3 + 255) / 256) * 256;
    // scratch: [counters 256 B][status_in n][status_out n][keep n][img_in][img_out]
    const size_t st_bytes = ((n + 255) / 256) * 256;
    int32_t rc = acm_ensure_scratch(ctx, 256 + 3 * st_bytes + 2 * img_bytes);
    if (rc) return rc;
    char* base = static_cast<char*>(ctx->d_scratch);
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(base);
    uint8_t* st_in = reinterpret_cast<uint8_t*>(base + 256);
    uint8_t* st_out = st_in + st_bytes;
    uint8_t* keep = st_out + st_bytes;
    uint8_t* img_in = keep + st_bytes;
    uint8_t* img_out = img_in + img_bytes;
    acm_points *uv_in = nullptr, *uv_out = nullptr;
    auto done = [&](int32_t code) { cudaStreamSynchronize(ctx->stream); acm_points_destroy(ctx, uv_in); acm_points_destroy(ctx, uv_out); return code; };
    if ((rc = acm_points_create(ctx, 2, n, ACM_F64, &uv_in))) return done(rc);
    if ((rc = acm_points_create(ctx, 2, n, ACM_F64, &uv_out))) return done(rc);
    // the trait's project of both models (with the image-bounds test of Pinhole / RadTan)
    if (n > 0) {
        if ((rc = acm_project(ctx, input_model, xyz, uv_in, st_in))) return done(rc);
        if ((rc = acm_project(ctx, output_model, xyz, uv_out, st_out))) return done(rc);
    }
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, 256, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(img_in, 0, 2 * img_bytes, ctx->stream);
    if (e != cudaSuccess) return done(acm_fail(ctx, ACM_ERR_CUDA, "image_quality_metrics: %s", cudaGetErrorString(e)));
    if (n > 0) {
        const int grid = grid_for(ctx, n, 256, 8);
        iq_keep_kernel<<<grid, 256, 0, ctx->stream>>>(st_in, st_out, comp<double>(uv_out, 0), comp<double>(uv_out, 1), n, (double)width,
                                                      (double)height, keep, d_cnt);
        ctx->launches++;
        if (npix > 0) {
            draw_points_kernel<<<grid, 256, 0, ctx->stream>>>(comp<double>(uv_in, 0), comp<double>(uv_in, 1), keep, n, 255, 255, 255, img_in, width, height);
            draw_points_kernel<<<grid, 256, 0, ctx->stream>>>(comp<double>(uv_out, 0), comp<double>(uv_out, 1), keep, n, 255, 255, 255, img_out, width, height);
            ctx->launches += 2;
        }
    }
    // sharded points (one rank per GPU): the images of the whole set are the OR of the ranks' images
    // (white on black: a byte-wise max), the kept count their sum
    if (ctx->n_ranks > 1) {
        if ((rc = acm_allreduce_sum_u64(ctx, d_cnt, 1))) return done(rc);
        if (npix > 0 && (rc = acm_allreduce_max_u8(ctx, img_in, 2 * img_bytes))) return done(rc);
    }
    unsigned long long kept = 0;
    e = cudaMemcpyAsync(&kept, d_cnt, sizeof(kept), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return done(acm_fail(ctx, ACM_ERR_CUDA, "image_quality_metrics: %s", cudaGetErrorString(e)));
    out->n_points = kept;
    if (kept == 0) return done(acm_fail(ctx, ACM_ERR_ZERO_PROJECTION_POINTS, "No valid projections"));  // image_quality.rs:306-308
    if (d_combined && npix > 0) {
        iq_compose_kernel<<<grid_for(ctx, npix, 256, 8), 256, 0, ctx->stream>>>(img_in, img_out, d_reference, d_combined, npix);
        ctx->launches++;
    }
    unsigned long long h[2];
    if ((rc = psnr_sums(ctx, img_in, img_out, npix, d_cnt + 2, h))) return done(rc);
    out->psnr = psnr_from_sums(h[0], h[1]);
    if ((rc = acm_image_ssim(ctx, img_in, img_out, width, height, &out->ssim))) return done(rc);
    return done(ACM_OK);
}
