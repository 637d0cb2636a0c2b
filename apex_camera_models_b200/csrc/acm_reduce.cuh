// Deterministic grid-wide reductions of small per-thread vectors.
//
//   per-thread registers -> fixed warp shuffle tree -> shared memory -> one partial per block in
//   global memory -> the last block to finish (ticket counter) combines the partials in block
//   order and leaves the result in `out`.
//
// No floating-point atomics anywhere, so the result depends only on the launch geometry.
// Three combine rules are supported, selected per slot range at compile time:
//   [0, NS)            plain sum
//   [NS, NS+NM)        max (min is carried as max of the negated value)
//   [NS+NM, +2*NP)     double-double pairs (hi, lo) combined with an error-free two-sum
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ void dd_two_sum(double a, double b, double& s, double& e) {
    s = __dadd_rn(a, b);
    double bb = __dsub_rn(s, a);
    e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}
// (hi, lo) += (bhi, blo)
__device__ __forceinline__ void dd_add(double& hi, double& lo, double bhi, double blo) {
    double s, e;
    dd_two_sum(hi, bhi, s, e);
    e = __dadd_rn(e, __dadd_rn(lo, blo));
    hi = __dadd_rn(s, e);
    lo = __dsub_rn(e, __dsub_rn(hi, s));
}
// (hi, lo) += a*b exactly (error-free product through an explicit fma)
__device__ __forceinline__ void dd_add_prod(double& hi, double& lo, double a, double b) {
    double p = __dmul_rn(a, b);
    double pe = __fma_rn(a, b, -p);
    dd_add(hi, lo, p, pe);
}

template <int NS, int NM, int NP, int BS = 256>
struct GridReduce {
    static_assert(BS == 128 || BS == 256, "block size must be 128 or 256");
    static constexpr int NW = BS / 32, NSUB = BS / 64;
    static constexpr int N = NS + NM + 2 * NP;
    static constexpr int NSLOT = NS + NM + NP;  // one thread owns one slot (a pair counts once)

    static __device__ __forceinline__ void combine_slot(double* a, const double* b, int slot) {
        if (slot < NS) a[0] += b[0];
        else if (slot < NS + NM) a[0] = fmax(a[0], b[0]);
        else dd_add(a[0], a[1], b[0], b[1]);
    }
    static __device__ __forceinline__ int slot_offset(int slot) { return slot < NS + NM ? slot : NS + NM + 2 * (slot - NS - NM); }
    static __device__ __forceinline__ int slot_width(int slot) { return slot < NS + NM ? 1 : 2; }

    // acc: this thread's N values.  partials: gridDim.x * N doubles.  Returns true in every thread
    // of the last block, after `out[0..N)` holds the final values (visible to that block).
    // Block size must be BS.
    static __device__ bool run(double* acc, double* __restrict__ partials, double* __restrict__ out, unsigned int* __restrict__ ticket) {
        __shared__ double sm[NW][N > 0 ? N : 1];
        __shared__ double fin4[NSUB][N > 0 ? N : 1];
        __shared__ bool is_last;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double v = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) sm[warp][i] = v;
        }
#pragma unroll
        for (int i = NS; i < NS + NM; ++i) {
            double v = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
            if (lane == 0) sm[warp][i] = v;
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            double hi = acc[NS + NM + 2 * i], lo = acc[NS + NM + 2 * i + 1];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double bh = __shfl_down_sync(0xffffffffu, hi, o), bl = __shfl_down_sync(0xffffffffu, lo, o);
                dd_add(hi, lo, bh, bl);
            }
            if (lane == 0) { sm[warp][NS + NM + 2 * i] = hi; sm[warp][NS + NM + 2 * i + 1] = lo; }
        }
        __syncthreads();
        if (threadIdx.x < NSLOT) {
            const int off = slot_offset(threadIdx.x), w = slot_width(threadIdx.x);
            double a[2] = {sm[0][off], w == 2 ? sm[0][off + 1] : 0.0};
#pragma unroll
            for (int k = 1; k < NW; ++k) {
                double b[2] = {sm[k][off], w == 2 ? sm[k][off + 1] : 0.0};
                combine_slot(a, b, threadIdx.x);
            }
            partials[(size_t)blockIdx.x * N + off] = a[0];
            if (w == 2) partials[(size_t)blockIdx.x * N + off + 1] = a[1];
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (!is_last) return false;
        __threadfence();
        // final pass: NSUB contiguous block ranges per slot, then combined in fixed order
        static_assert(NSLOT <= 64, "final pass assumes <= 64 slots");
        const int slot = threadIdx.x & 63, sub = threadIdx.x >> 6;
        const unsigned int nb = gridDim.x;
        const unsigned int b0 = (unsigned int)(((size_t)nb * sub) / NSUB), b1 = (unsigned int)(((size_t)nb * (sub + 1)) / NSUB);
        if (slot < NSLOT) {
            const int off = slot_offset(slot), w = slot_width(slot);
            double a[2];
            a[0] = (slot >= NS && slot < NS + NM) ? -INFINITY : 0.0; a[1] = 0.0;
            for (unsigned int b = b0; b < b1; ++b) {
                double v[2] = {__ldcg(partials + (size_t)b * N + off), w == 2 ? __ldcg(partials + (size_t)b * N + off + 1) : 0.0};
                combine_slot(a, v, slot);
            }
            fin4[sub][off] = a[0];
            if (w == 2) fin4[sub][off + 1] = a[1];
        }
        __syncthreads();
        if (threadIdx.x < NSLOT) {
            const int off = slot_offset(threadIdx.x), w = slot_width(threadIdx.x);
            double a[2] = {fin4[0][off], w == 2 ? fin4[0][off + 1] : 0.0};
#pragma unroll
            for (int k = 1; k < NSUB; ++k) {
                double b[2] = {fin4[k][off], w == 2 ? fin4[k][off + 1] : 0.0};
                combine_slot(a, b, threadIdx.x);
            }
            out[off] = a[0];
            if (w == 2) out[off + 1] = a[1];
        }
        if (threadIdx.x == 0) *ticket = 0;  // re-arm for the next launch on this stream
        __syncthreads();
        return true;
    }
};
