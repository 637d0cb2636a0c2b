// Internal declarations shared by the translation units of libacm.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/acm.h"

// ---------------------------------------------------------------------------------------
// Camera parameter block passed by value to every kernel (<= 160 B, lives in the kernel's
// constant bank).  k0..k2 are per-model constants the reference recomputes per point; they
// are evaluated once on the host with the reference's own operation order (and glibc's
// tan for FOV) so hoisting them does not change a single bit.
// ---------------------------------------------------------------------------------------
struct CamParams {
    double fx, fy, cx, cy;
    double d[5];
    double W, H;     // resolution as f64 (`width as f64`)
    double k0, k1, k2;
    double ifx, ify;  // RN(1/fx), RN(1/fy) from the host's IEEE division: (u - cx) / fx on the device is then acm_div_by()
    int32_t model;
    int32_t has_resolution;  // width > 0 && height > 0 (kannala_brandt.rs:447-448)
    int32_t fast_div;        // 2^-100 <= |fx|, |fy| <= 2^100: acm_div_by() is bit-identical to the division
    int32_t kb_fast;         // Kannala-Brandt: the host verified that Newton's method converges for every pixel (acm_make_cam_params)
};

struct acm_points {
    int32_t dim;
    int32_t dtype;
    size_t n;
    size_t stride_bytes;  // distance between components, multiple of 256
    void* base;
    size_t alloc_bytes;   // size of the allocation behind `base` (>= dim * stride_bytes)
};

#define ACM_FREE_LIST 8

#define ACM_MAX_PEERS 8

struct NcclApi;  // dlopen'ed NCCL entry points (acm_core.cu)

struct acm_ctx {
    int device;
    cudaStream_t stream;
    bool owns_stream;
    cudaStream_t copy_stream;  // second stream for chunked host pipelines
    cudaEvent_t t0, t1;
    cudaEvent_t chunk_ev[4];
    int sm_count;
    size_t l2_bytes;
    int cc;
    uint64_t launches;
    std::string err;
    // scratch owned by the context
    double* d_partials;     // per-block partial sums of the reduction kernels
    size_t partials_cap;    // in doubles
    double* d_reduce;       // final reduced vector (2048 doubles; gathers use the first 1024)
    unsigned int* d_ticket; // last-block-done counters
    double* h_reduce;       // pinned mirror of d_reduce (2048 doubles)
    void* d_lm;             // LmState
    void* h_lm;             // pinned mirror
    void* d_stage[2];       // staging for host pipelines
    size_t stage_cap;
    void* h_stage;          // pinned staging for pageable host buffers
    size_t h_stage_cap;
    void* d_scratch;        // grow-only arena for the temporaries of the util entry points
    size_t scratch_cap;
    // destroyed point buffers are kept for the next acm_points_create of a similar size (a
    // cudaMalloc / cudaFree pair costs milliseconds at 10 M points, the kernels around it microseconds)
    void* free_ptr[ACM_FREE_LIST];
    size_t free_bytes[ACM_FREE_LIST];
    acm_points* cache3;     // device buffers kept between *_host calls (grow-only)
    acm_points* cache2;
    size_t cache_cap;
    // NCCL
    void* comm;
    int n_ranks, rank;
    // NVLink peer exchange (acm_peer_*)
    double* peer_local;          // this rank's exchange buffer (exported through CUDA IPC)
    void* peer_mapped[ACM_MAX_PEERS];  // peers' buffers opened in this process (nullptr for self)
    double** d_peer_ptrs;        // device array [n_peers] of every rank's buffer, indexed by rank
    int peer_n, peer_rank;
    unsigned long long peer_seq; // exchange counter, advances identically on every rank
};

// exchange buffer: 2 alternating sets x ACM_MAX_PEERS rank slots x (64 values + flag + padding)
#define ACM_PEER_SLOT_DOUBLES 72
#define ACM_PEER_BUFFER_DOUBLES (2 * ACM_MAX_PEERS * ACM_PEER_SLOT_DOUBLES)

struct PeerArgs {
    double* const* bufs;         // nullptr: exchange disabled
    int n_ranks, rank;
    unsigned long long seq;
};

int32_t acm_fail(acm_ctx* ctx, int32_t code, const char* fmt, ...);
void acm_set_global_error(const char* msg);

#define ACM_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return acm_fail((ctx), ACM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define ACM_CHECK_LAUNCH(ctx)                                                                 \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return acm_fail((ctx), ACM_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define ACM_REQUIRE(ctx, cond, msg)                                                           \
    do {                                                                                      \
        if (!(cond)) return acm_fail((ctx), ACM_ERR_INVALID_ARG, "%s", (msg));                \
    } while (0)

// host helpers implemented in acm_core.cu
int32_t acm_make_cam_params(acm_ctx* ctx, const acm_camera* cam, CamParams* out);
int32_t acm_device_malloc(acm_ctx* ctx, void** out, size_t bytes);  // cudaMalloc; empties the free list and retries when out of memory
int32_t acm_ensure_stage(acm_ctx* ctx, size_t bytes);
int32_t acm_ensure_host_stage(acm_ctx* ctx, size_t bytes);
int32_t acm_ensure_partials(acm_ctx* ctx, size_t doubles);
int32_t acm_ensure_scratch(acm_ctx* ctx, size_t bytes);  // ctx->d_scratch holds >= bytes afterwards (256-byte aligned)
int32_t acm_allreduce_sum_f64(acm_ctx* ctx, double* d_buf, size_t count);
int32_t acm_allreduce_sum_u64(acm_ctx* ctx, unsigned long long* d_buf, size_t count);
int32_t acm_allreduce_max_u8(acm_ctx* ctx, uint8_t* d_buf, size_t count);
int32_t acm_rank_gather_to_host(acm_ctx* ctx, int count);  // d_reduce[0..count) of every rank -> h_reduce[rank][count]
int32_t acm_points_upload_any(acm_ctx* ctx, acm_points* p, const double* host_aos, size_t n, size_t dst_offset);

template <typename T>
static inline T* comp(const acm_points* p, int c) {
    return reinterpret_cast<T*>(static_cast<char*>(p->base) + (size_t)c * p->stride_bytes);
}

static inline int grid_for(const acm_ctx* ctx, size_t work_items, int block, int blocks_per_sm) {
    size_t need = (work_items + (size_t)block - 1) / (size_t)block;
    size_t cap = (size_t)ctx->sm_count * (size_t)blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
