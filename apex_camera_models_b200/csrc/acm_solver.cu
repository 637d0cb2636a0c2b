// Fused linearisation kernel, deterministic reduction, device-resident Levenberg-Marquardt.
//
// One LM iteration = one streaming pass (linearize_kernel at the trial point: cost, H and g
// together, "speculative linearisation") + an all-reduce of <= 54 doubles when a communicator
// is attached + one single-thread step kernel (accept / reject, damping update, P x P
// Cholesky with Jacobi scaling, next trial point).  The host only polls a done flag every
// `check_every` iterations; kernels launched after convergence exit on their first load.
//
// Determinism: fixed grid, per-thread sequential accumulation over a grid-stride range,
// fixed shuffle tree, per-block partials, the last block to finish sums the partials in
// block order.  No floating-point atomics.  The result depends on the grid size only.
#include "acm_linearize.cuh"
#include "acm_reduce.cuh"
#include "acm_models.cuh"
#include "acm_pointjac.cuh"

#include <stdlib.h>

#include <chrono>

struct LmState {
    double x[ACM_MAX_PARAMS], xt[ACM_MAX_PARAMS];
    double H[ACM_MAX_PARAMS * ACM_MAX_PARAMS], g[ACM_MAX_PARAMS];
    double lower[ACM_MAX_PARAMS], upper[ACM_MAX_PARAMS];
    double cost, lambda, nu;
    double pred, dnorm, xnorm;
    double cost_tol, param_tol, grad_tol;
    double initial_cost, n_valid;
    int32_t max_iter, iterations, passes, status, done, first, P, pad;
};
static_assert(sizeof(LmState) <= 4096, "LmState must fit the context's 4 KiB slot");

// ---------------------------------------------------------------------------------------
// LM step: serial algebra on P <= 9 unknowns, run by one thread out of shared memory.
// Written for a small register footprint (arrays in the shared LmWork, rolled loops, not
// inlined) because it is also called from the tail of the streaming kernel.
// Divisions are hoisted into reciprocals (1/D_i, 1/L_ii): a serial f64 division costs ~250
// cycles and the first version spent most of its 12 us on ~50 of them.
// ---------------------------------------------------------------------------------------
struct LmWork {
    double Ht[ACM_MAX_PARAMS * ACM_MAX_PARAMS], A[ACM_MAX_PARAMS * ACM_MAX_PARAMS], L[ACM_MAX_PARAMS * ACM_MAX_PARAMS];
    double gt[ACM_MAX_PARAMS], invD[ACM_MAX_PARAMS], gs[ACM_MAX_PARAMS], st[ACM_MAX_PARAMS], dx[ACM_MAX_PARAMS], yv[ACM_MAX_PARAMS],
        invdiag[ACM_MAX_PARAMS];
    double red[64];
};

// Cholesky solve of the damped, Jacobi-scaled system (P <= 9) by one thread: operands are pulled
// from the shared LmWork into registers once, the factorisation is fully unrolled (compile-time P),
// the result goes back to w->st.  Reciprocals of the diagonal replace the divisions of the
// substitutions (a dependent f64 division costs ~150 cycles).
template <int P>
__device__ __noinline__ bool chol_solve_work(LmWork* w) {
    double L[P * (P + 1) / 2], inv[P], yv[P], x[P];
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) L[i * (i + 1) / 2 + j] = w->A[i * P + j];
#pragma unroll
    for (int i = 0; i < P; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double sum = L[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) sum -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            if (i == j) {
                if (!(sum > 0.0)) return false;
                const double d = sqrt(sum);
                L[i * (i + 1) / 2 + i] = d;
                inv[i] = 1.0 / d;
            } else {
                L[i * (i + 1) / 2 + j] = sum * inv[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
        double sum = w->gs[i];
#pragma unroll
        for (int k = 0; k < i; ++k) sum -= L[i * (i + 1) / 2 + k] * yv[k];
        yv[i] = sum * inv[i];
    }
#pragma unroll
    for (int i = P - 1; i >= 0; --i) {
        double sum = yv[i];
#pragma unroll
        for (int k = i + 1; k < P; ++k) sum -= L[k * (k + 1) / 2 + i] * x[k];
        x[i] = sum * inv[i];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) w->st[i] = x[i];
    return true;
}

__device__ inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Decision part of the step (thread 0): accept / reject the trial point, update the damping,
// test convergence.  Returns true when the trial point was accepted.
__device__ __noinline__ bool lm_decide(int P, LmState* s, const LmWork* w, double cost_t) {
    s->passes++;
    if (!(cost_t == cost_t)) { s->status = 4; s->done = 1; return false; }  // NaN sums (poisoned exchange or NaN observations): stop
    if (s->first) {
        s->first = 0;
        s->initial_cost = cost_t;
        return true;
    }
    const bool small_step = s->dnorm <= s->param_tol * (s->xnorm + s->param_tol);
    if (s->pred > 0.0 && cost_t < s->cost) {
        const double rho = (s->cost - cost_t) / s->pred;
        const double dcost = s->cost - cost_t, cost_old = s->cost;
        const double q = 2.0 * rho - 1.0, f = 1.0 - q * q * q;
        s->lambda *= (f > 1.0 / 3.0) ? f : 1.0 / 3.0;
        s->nu = 2.0;
        if (s->lambda < 1e-15) s->lambda = 1e-15;
        double gmax = 0.0;
#pragma unroll 1
        for (int i = 0; i < P; ++i) gmax = fmax(gmax, fabs(w->gt[i]));
        if (dcost <= s->cost_tol * cost_old) { s->status = 0; s->done = 1; }
        else if (small_step) { s->status = 1; s->done = 1; }
        else if (gmax <= s->grad_tol) { s->status = 2; s->done = 1; }
        return true;
    }
    if (small_step) { s->status = 1; s->done = 1; }
    else {
        s->lambda *= s->nu; s->nu *= 2.0;
        if (s->lambda > 1e30) { s->status = 4; s->done = 1; }
    }
    return false;
}

// Cooperative step: `nthreads` (>= 81) threads copy the 1.1 KB state and the reduced accumulators
// into shared memory; thread 0 decides and factors, the element-wise parts (copies, scaled matrix,
// trial point, quadratic-model rows) are spread over the threads; the state is copied back.
// The arithmetic of every scalar is the same as in the serial reference (oracle/acm_oracle_solver.c),
// including the order of the few sums, so both walk the same trajectory.
template <int M, int KIND>
__device__ __forceinline__ void lm_step_block(LmState* __restrict__ s, const double* __restrict__ red, LmState* sh, LmWork* w,
                                              int tid, int nthreads) {
    static_assert(sizeof(LmState) % sizeof(double) == 0, "LmState is copied as doubles");
    constexpr int NW = sizeof(LmState) / sizeof(double);
    constexpr int P = LinOps<M, KIND>::P;
    __shared__ int flag_accept, flag_ok;
    __shared__ double s_cost_t, s_cnt;
    double* shw = reinterpret_cast<double*>(sh);
    const double* gw = reinterpret_cast<const double*>(s);
    for (int i = tid; i < NW; i += nthreads) shw[i] = __ldcg(gw + i);
    for (int i = tid; i < LinOps<M, KIND>::NACC; i += nthreads) w->red[i] = __ldcg(red + i);
    __syncthreads();
    if (sh->done) return;
    if (tid == 0) {
        double cost_t, cnt;
        LinOps<M, KIND>::unpack(w->red, sh->xt, w->Ht, w->gt, &cost_t, &cnt);  // the pass ran at the trial point xt
        s_cost_t = cost_t; s_cnt = cnt;
        flag_accept = lm_decide(P, sh, w, cost_t) ? 1 : 0;
    }
    __syncthreads();
    if (flag_accept) {
        if (tid < P) { sh->x[tid] = sh->xt[tid]; sh->g[tid] = w->gt[tid]; }
        if (tid < P * P) sh->H[tid] = w->Ht[tid];
        if (tid == 0) { sh->cost = s_cost_t; sh->n_valid = s_cnt; }
    }
    __syncthreads();
    if (!sh->done) {
        // next trial point from (H, g, lambda) at the accepted x
        for (;;) {
            if (tid == 0) {
                flag_ok = 1;
                if (sh->iterations >= sh->max_iter) { sh->status = 3; sh->done = 1; flag_ok = -1; }
                else sh->iterations++;
            }
            if (tid < P) { const double d = sqrt(sh->H[tid * P + tid]); w->invD[tid] = (d > 1e-300) ? 1.0 / d : 1.0; }
            __syncthreads();
            if (flag_ok < 0) break;
            if (tid < P * P) {
                const int i = tid / P, j = tid - i * P;
                double a = sh->H[tid] * w->invD[i] * w->invD[j];
                if (i == j) a += sh->lambda;
                w->A[tid] = a;
            }
            if (tid < P) w->gs[tid] = -sh->g[tid] * w->invD[tid];
            __syncthreads();
            if (tid == 0) {
                if (!chol_solve_work<P>(w)) {
                    sh->lambda *= sh->nu; sh->nu *= 2.0;
                    flag_ok = 0;
                    if (sh->lambda > 1e30) { sh->status = 4; sh->done = 1; flag_ok = -1; }
                }
            }
            __syncthreads();
            if (flag_ok != 0) break;
        }
        if (flag_ok > 0) {
            if (tid < P) {
                sh->xt[tid] = clampd(sh->x[tid] + w->st[tid] * w->invD[tid], sh->lower[tid], sh->upper[tid]);
                w->dx[tid] = sh->xt[tid] - sh->x[tid];
            }
            __syncthreads();
            if (tid < P) {
                double hd = 0.0;
#pragma unroll 1
                for (int j = 0; j < P; ++j) hd += sh->H[tid * P + j] * w->dx[j];
                w->yv[tid] = w->dx[tid] * (sh->g[tid] + 0.5 * hd);   // row of the quadratic model
            }
            __syncthreads();
            if (tid == 0) {
                double xnorm = 0.0, dnorm = 0.0, pred = 0.0;
#pragma unroll 1
                for (int i = 0; i < P; ++i) { xnorm += sh->x[i] * sh->x[i]; dnorm += w->dx[i] * w->dx[i]; pred -= w->yv[i]; }
                sh->xnorm = sqrt(xnorm); sh->dnorm = sqrt(dnorm); sh->pred = pred;
            }
        }
    }
    __syncthreads();
    double* gout = reinterpret_cast<double*>(s);
    for (int i = tid; i < NW; i += nthreads) gout[i] = shw[i];
}

// Stand-alone step kernel: used when an all-reduce sits between the pass and the step (N > 1).
template <int M, int KIND>
__global__ void __launch_bounds__(128) lm_step_kernel(LmState* __restrict__ s, const double* __restrict__ red) {
    __shared__ LmState sh;
    __shared__ LmWork work;
    lm_step_block<M, KIND>(s, red, &sh, &work, threadIdx.x, blockDim.x);
}

// ---------------------------------------------------------------------------------------
// All-reduce over NVLink peer memory, fused into the tail of the streaming kernel.
// Executed by the last block of every rank's kernel (the ranks run on different GPUs, so they are
// all resident).  Rank r stores its NACC sums into slot [set][r] of EVERY rank's exchange buffer
// (plain stores to peer-mapped addresses travel over NVLink), publishes them with a release store
// of the exchange counter, waits (acquire loads of its own buffer) until all slots carry that
// counter and adds the slots in rank order: the totals are bit-identical on every rank, which
// keeps the ranks' LM decisions -- and therefore their kernel sequences -- in lock step.
// Two slot sets alternate so that a rank that runs ahead by one exchange cannot overwrite data a
// slower rank still has to read.  The spin is bounded: on time-out the sums are poisoned with NaN
// (the solve then stops with status 4) instead of hanging the GPU.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int NACC>
__device__ __forceinline__ void peer_exchange(const PeerArgs& peer, double* __restrict__ out) {
    const int t = threadIdx.x;
    const size_t set_off = (size_t)(peer.seq & 1ULL) * ACM_MAX_PEERS * ACM_PEER_SLOT_DOUBLES;
    const size_t my_slot = set_off + (size_t)peer.rank * ACM_PEER_SLOT_DOUBLES;
    __shared__ int timed_out;
    if (t == 0) timed_out = 0;
    // 1. my sums -> my slot in every rank's buffer
    for (int i = t; i < NACC * peer.n_ranks; i += blockDim.x) {
        const int r = i / NACC, k = i - r * NACC;
        peer.bufs[r][my_slot + k] = __ldcg(out + k);
    }
    __threadfence_system();
    __syncthreads();
    if (t < peer.n_ranks) st_release_sys(reinterpret_cast<unsigned long long*>(peer.bufs[t] + my_slot + 64), peer.seq);
    // 2. wait for every rank's slot in my own buffer
    if (t < peer.n_ranks) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(peer.bufs[peer.rank] + set_off + (size_t)t * ACM_PEER_SLOT_DOUBLES + 64);
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) != peer.seq) {
            if (clock64() - t0 > 4000000000LL) { timed_out = 1; break; }  // ~2 s
            __nanosleep(64);
        }
    }
    __syncthreads();
    // 3. add the slots in rank order
    if (t < NACC) {
        const double* mine = peer.bufs[peer.rank] + set_off;
        double tot = 0.0;
        for (int r = 0; r < peer.n_ranks; ++r) tot += __ldcv(mine + (size_t)r * ACM_PEER_SLOT_DOUBLES + t);
        out[t] = timed_out ? __longlong_as_double(0x7ff8000000000000LL) : tot;
    }
    __threadfence();
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// linearize kernel
// ---------------------------------------------------------------------------------------
// Per-model streaming configuration, measured on 100 M points (scripts/lin_bench.py; GB/s register
// prefetch -> cp.async ring): DEPTH > 0 = every thread keeps that many packets in flight in a
// shared-memory ring fed by cp.async; 0 = the next packet is prefetched into registers.
//   DS 5.88 -> 6.24 TB/s (3 deep), EUCM 5.63 -> 5.79, UCM 5.83 -> 6.47 (2 deep), FOV 4.50 -> 4.97
//   (3 deep, 128 threads), RadTan 4.09 -> 4.25 (2 deep; 5.1 with the structured accumulation; 128-thread blocks capped at 168
//   registers for 12 warps/SM measured 4.2-4.8); Pinhole (already at 7.1 TB/s) is faster without the ring.
// MIN_BLOCKS > 1 caps the registers through __launch_bounds__.
template <int M> struct LinStreamDefault { static constexpr int DEPTH = 0, BLOCK = 256, MIN_BLOCKS = 0; };
#ifndef ACM_LIN_NO_RING  // A/B aid: -DACM_LIN_NO_RING builds every model with the register prefetch
template <> struct LinStreamDefault<ACM_MODEL_DOUBLE_SPHERE> { static constexpr int DEPTH = 3, BLOCK = 256, MIN_BLOCKS = 0; };
template <> struct LinStreamDefault<ACM_MODEL_EUCM> { static constexpr int DEPTH = 3, BLOCK = 256, MIN_BLOCKS = 0; };
template <> struct LinStreamDefault<ACM_MODEL_UCM> { static constexpr int DEPTH = 2, BLOCK = 256, MIN_BLOCKS = 0; };
template <> struct LinStreamDefault<ACM_MODEL_FOV> { static constexpr int DEPTH = 3, BLOCK = 128, MIN_BLOCKS = 0; };
template <> struct LinStreamDefault<ACM_MODEL_RADTAN> { static constexpr int DEPTH = 2, BLOCK = 256, MIN_BLOCKS = 0; };
#endif
// KB (37 accumulators, 166 registers): 3 blocks of 128 threads.  Same-box A/B (scripts/ab_lin.sh; boxes of the pool differ by
// up to 40 % on this FP64-bound kernel, so only same-box comparisons count): register prefetch 4.42 TB/s, ring 1 deep 3.95,
// 2 deep 4.51, 3 deep 4.59; capped at 128 registers (MIN_BLOCKS = 4, 36-byte spill) 4.19.
template <> struct LinStreamDefault<ACM_MODEL_KANNALA_BRANDT> { static constexpr int DEPTH = 3, BLOCK = 128, MIN_BLOCKS = 0; };
template <int M> struct LinStream : LinStreamDefault<M> {};
#ifdef ACM_EXP_MODEL  // tuning aid: -DACM_EXP_MODEL=<id> -DACM_EXP_DEPTH= -DACM_EXP_BLOCK= -DACM_EXP_MINB= overrides one model
template <> struct LinStream<ACM_EXP_MODEL> { static constexpr int DEPTH = ACM_EXP_DEPTH, BLOCK = ACM_EXP_BLOCK, MIN_BLOCKS = ACM_EXP_MINB; };
#endif

template <int M, int KIND, int BS>
__global__ void __launch_bounds__(BS, (BS == LinStream<M>::BLOCK ? LinStream<M>::MIN_BLOCKS : 0)) linearize_kernel(LinParams hp, LmState* __restrict__ lm, int fuse_step, PeerArgs peer,
                                                       const double2* __restrict__ X,
                                                        const double2* __restrict__ Y, const double2* __restrict__ Z,
                                                        const double2* __restrict__ U, const double2* __restrict__ V, size_t n,
                                                        double pen2x2, double* __restrict__ partials, double* __restrict__ out,
                                                        unsigned int* __restrict__ ticket) {
    using LM_ = LinOps<M, KIND>;
    constexpr int ND = LM_::ND;
    constexpr int NACC = LM_::NACC;
    static_assert(NACC <= 64, "final-pass layout assumes <= 64 accumulators");

    LinParams p = hp;
    if (lm) {
        // one round trip: the done flag and the trial parameters are fetched together
        const int done = lm->done;
        p.fx = lm->xt[0]; p.fy = lm->xt[1]; p.cx = lm->xt[2]; p.cy = lm->xt[3];
#pragma unroll
        for (int k = 0; k < ND; ++k) p.d[k] = lm->xt[4 + k];
        if (done) return;  // converged earlier in this enqueue batch
        lin_derive(M, p);
    }

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

    const size_t npairs = n >> 1;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int DEPTH = LinStream<M>::DEPTH;
    if constexpr (DEPTH > 0) {
    // cp.async ring: every thread keeps DEPTH packets (5 x 16 B each) in flight in its own
    // shared-memory slots -- deeper than a register prefetch could afford -- and reads them back with
    // five conflict-free LDS.128.  Only the owning thread touches a slot, so wait_group is all the
    // synchronisation needed.  ring[stage][array][thread].
    extern __shared__ double2 lin_ring[];
    const double2* const src[5] = {X, Y, Z, U, V};
    auto issue = [&](int stage, size_t idx) {
        if (idx < npairs) {
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&lin_ring[(stage * 5 + a) * BS + threadIdx.x]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src[a] + idx) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < DEPTH; ++s) issue(s, i + (size_t)s * stride);
    int stage = 0;
#pragma unroll 1
    while (i < npairs) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH > 0 ? DEPTH - 1 : 0) : "memory");
        const double2 x = lin_ring[(stage * 5 + 0) * BS + threadIdx.x], y = lin_ring[(stage * 5 + 1) * BS + threadIdx.x],
                      z = lin_ring[(stage * 5 + 2) * BS + threadIdx.x], u = lin_ring[(stage * 5 + 3) * BS + threadIdx.x],
                      v = lin_ring[(stage * 5 + 4) * BS + threadIdx.x];
        LM_::point(acc, p, x.x, y.x, z.x, u.x, v.x);
        LM_::point(acc, p, x.y, y.y, z.y, u.y, v.y);
        issue(stage, i + (size_t)DEPTH * stride);  // after the packet has been consumed
        stage = (stage + 1 == DEPTH) ? 0 : stage + 1;
        i += stride;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
    // Software-pipelined stream: the five 16-byte loads of the next pair of points are in flight
    // while the current pair is evaluated, so that the few resident warps (accumulators cost
    // registers) still keep enough bytes in flight to cover the HBM latency.
    double2 x, y, z, u, v;
    if (i < npairs) { x = __ldcs(X + i); y = __ldcs(Y + i); z = __ldcs(Z + i); u = __ldcs(U + i); v = __ldcs(V + i); }
#pragma unroll 1
    while (i < npairs) {
        const size_t nx = i + stride;
        const size_t j = nx < npairs ? nx : i;  // clamp: the tail re-reads its own (cached) packet
        const double2 x2 = __ldcs(X + j), y2 = __ldcs(Y + j), z2 = __ldcs(Z + j), u2 = __ldcs(U + j), v2 = __ldcs(V + j);
        LM_::point(acc, p, x.x, y.x, z.x, u.x, v.x);
        LM_::point(acc, p, x.y, y.y, z.y, u.y, v.y);
        x = x2; y = y2; z = z2; u = u2; v = v2;
        i = nx;
    }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const size_t t = n - 1;
        const double* Xs = reinterpret_cast<const double*>(X); const double* Ys = reinterpret_cast<const double*>(Y);
        const double* Zs = reinterpret_cast<const double*>(Z); const double* Us = reinterpret_cast<const double*>(U);
        const double* Vs = reinterpret_cast<const double*>(V);
        LM_::point(acc, p, Xs[t], Ys[t], Zs[t], Us[t], Vs[t]);
    }

    if (GridReduce<NACC, 0, 0, BS>::run(acc, partials, out, ticket)) {
        // invalid points carry the residual (pen, pen): cost += pen^2 per invalid point
        if (threadIdx.x == 0 && pen2x2 != 0.0) out[LM_::COST] += pen2x2 * ((double)n - out[LM_::COUNT]);
        if (peer.bufs) {
            __threadfence();
            __syncthreads();
            peer_exchange<NACC>(peer, out);
        }
        if (lm && fuse_step) {
            // single GPU: no all-reduce between the pass and the step, so the last block takes the
            // LM step right here (saves a launch and the global round trip of the sums)
            __shared__ LmState sh;
            __shared__ LmWork work;
            __threadfence();
            __syncthreads();
            lm_step_block<M, KIND>(lm, out, &sh, &work, threadIdx.x, BS);
        }
    }
}

template <int M, int KIND, int BS>
static int32_t launch_linearize_bs(acm_ctx* ctx, const LinParams& hp, LmState* d_lm, int fuse_step, const PeerArgs& peer,
                                   const acm_points* xyz, const acm_points* uv, double invalid_penalty) {
    static int blocks_per_sm = 0;
    constexpr size_t ring_bytes = (size_t)LinStream<M>::DEPTH * 5 * BS * sizeof(double2);
    if (!blocks_per_sm) {
        int b = 0;
        if (ring_bytes > 0) ACM_CUDA(ctx, cudaFuncSetAttribute(linearize_kernel<M, KIND, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
        ACM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, linearize_kernel<M, KIND, BS>, BS, ring_bytes));
        blocks_per_sm = b > 0 ? b : 1;
    }
    const size_t n = xyz->n;
    int grid = grid_for(ctx, (n >> 1) + 1, BS, blocks_per_sm);
    int32_t rc = acm_ensure_partials(ctx, (size_t)ctx->sm_count * 32 * 64);
    if (rc) return rc;
    linearize_kernel<M, KIND, BS><<<grid, BS, ring_bytes, ctx->stream>>>(
        hp, d_lm, fuse_step, peer, comp<double2>(xyz, 0), comp<double2>(xyz, 1), comp<double2>(xyz, 2), comp<double2>(uv, 0), comp<double2>(uv, 1), n,
        2.0 * invalid_penalty * invalid_penalty, ctx->d_partials, ctx->d_reduce, ctx->d_ticket);
    ACM_CHECK_LAUNCH(ctx);
    return ACM_OK;
}

// Block size per model: the accumulators of the wide models (KB: 37 doubles) push the
// kernel past 128 registers/thread; 128-thread blocks then pack one more block per SM.
// ACM_LIN_BLOCK=128|256 overrides (tuning aid).
template <int M, int KIND>
static int32_t launch_linearize(acm_ctx* ctx, const LinParams& hp, LmState* d_lm, int fuse_step, const PeerArgs& peer,
                                const acm_points* xyz, const acm_points* uv, double invalid_penalty) {
    static int bs = 0;
    if (!bs) {
        bs = LinStream<M>::BLOCK;
        const char* e = getenv("ACM_LIN_BLOCK");
        if (e && (atoi(e) == 128 || atoi(e) == 256)) bs = atoi(e);
    }
    if (bs == 128) return launch_linearize_bs<M, KIND, 128>(ctx, hp, d_lm, fuse_step, peer, xyz, uv, invalid_penalty);
    return launch_linearize_bs<M, KIND, 256>(ctx, hp, d_lm, fuse_step, peer, xyz, uv, invalid_penalty);
}

#define ACM_DISPATCH_LIN(model, kind, ...)                                                                       \
    do {                                                                                                         \
        if ((kind) == ACM_RESIDUAL_ALGEBRAIC) {                                                                  \
            constexpr int KIND = ACM_RESIDUAL_ALGEBRAIC;                                                         \
            switch (model) {                                                                                     \
                case ACM_MODEL_UCM: { constexpr int M = ACM_MODEL_UCM; __VA_ARGS__; break; }                     \
                case ACM_MODEL_EUCM: { constexpr int M = ACM_MODEL_EUCM; __VA_ARGS__; break; }                   \
                case ACM_MODEL_DOUBLE_SPHERE: { constexpr int M = ACM_MODEL_DOUBLE_SPHERE; __VA_ARGS__; break; } \
                default: return acm_fail(ctx, ACM_ERR_INVALID_ARG, "the algebraic residual exists only for UCM, EUCM and Double Sphere"); \
            }                                                                                                    \
        } else if ((kind) == ACM_RESIDUAL_PIXEL) {                                                               \
            constexpr int KIND = ACM_RESIDUAL_PIXEL;                                                             \
            switch (model) {                                                                                     \
                case ACM_MODEL_PINHOLE: { constexpr int M = ACM_MODEL_PINHOLE; __VA_ARGS__; break; }             \
                case ACM_MODEL_RADTAN: { constexpr int M = ACM_MODEL_RADTAN; __VA_ARGS__; break; }               \
                case ACM_MODEL_KANNALA_BRANDT: { constexpr int M = ACM_MODEL_KANNALA_BRANDT; __VA_ARGS__; break; } \
                case ACM_MODEL_UCM: { constexpr int M = ACM_MODEL_UCM; __VA_ARGS__; break; }                     \
                case ACM_MODEL_EUCM: { constexpr int M = ACM_MODEL_EUCM; __VA_ARGS__; break; }                   \
                case ACM_MODEL_DOUBLE_SPHERE: { constexpr int M = ACM_MODEL_DOUBLE_SPHERE; __VA_ARGS__; break; } \
                case ACM_MODEL_FOV: { constexpr int M = ACM_MODEL_FOV; __VA_ARGS__; break; }                     \
                default: return acm_fail(ctx, ACM_ERR_INVALID_ARG, "unknown camera model id %d", (int)(model));  \
            }                                                                                                    \
        } else return acm_fail(ctx, ACM_ERR_INVALID_ARG, "unknown residual kind %d", (int)(kind));              \
    } while (0)

static int32_t check_lin_args(acm_ctx* ctx, const acm_camera* cam, const acm_points* xyz, const acm_points* uv) {
    ACM_REQUIRE(ctx, cam && xyz && uv, "linearize: null argument");
    ACM_REQUIRE(ctx, xyz->dim == 3 && uv->dim == 2, "linearize: xyz must have dim 3 and uv dim 2");
    ACM_REQUIRE(ctx, xyz->dtype == ACM_F64 && uv->dtype == ACM_F64, "linearize: f64 point buffers required");
    if (xyz->n != uv->n) return acm_fail(ctx, ACM_ERR_INVALID_PARAMS, "Number of 2D and 3D points must match");
    ACM_REQUIRE(ctx, cam->model >= 0 && cam->model <= 6 && cam->n_params == acm_n_params(cam->model), "linearize: bad camera block");
    return ACM_OK;
}

static void make_lin_params(const acm_camera* cam, LinParams* p) {
    memset(p, 0, sizeof(*p));
    p->fx = cam->params[0]; p->fy = cam->params[1]; p->cx = cam->params[2]; p->cy = cam->params[3];
    for (int i = 4; i < cam->n_params; ++i) p->d[i - 4] = cam->params[i];
    lin_derive(cam->model, *p);
}

// Enqueue one pass (+ the cross-rank sum).  With peers attached the sum happens inside the kernel
// (and so can the LM step); otherwise an NCCL all-reduce follows when a communicator is attached.
// *fused_step tells the caller whether the kernel also takes the LM step.
static int32_t enqueue_linearize(acm_ctx* ctx, const acm_camera* cam, int32_t kind, LmState* d_lm, const acm_points* xyz,
                                 const acm_points* uv, double invalid_penalty, int* nacc, int* fused_step) {
    LinParams hp;
    make_lin_params(cam, &hp);
    PeerArgs peer{nullptr, 1, 0, 0ULL};
    const bool use_peer = ctx->peer_n > 1 && !getenv("ACM_NO_PEER_EXCHANGE");
    if (use_peer) {
        peer.bufs = ctx->d_peer_ptrs; peer.n_ranks = ctx->peer_n; peer.rank = ctx->peer_rank;
        peer.seq = ++ctx->peer_seq;
    }
    const bool need_nccl = !use_peer && ctx->comm && ctx->n_ranks > 1;
    const int fuse = (d_lm && !need_nccl && !getenv("ACM_LM_NO_FUSE")) ? 1 : 0;
    if (fused_step) *fused_step = fuse;
    ACM_DISPATCH_LIN(cam->model, kind, {
        int32_t rc = launch_linearize<M, KIND>(ctx, hp, d_lm, fuse, peer, xyz, uv, invalid_penalty);
        if (rc) return rc;
        *nacc = LinOps<M, KIND>::NACC;
    });
    if (need_nccl) return acm_allreduce_sum_f64(ctx, ctx->d_reduce, (size_t)*nacc);
    return ACM_OK;
}

static int32_t unpack_host(acm_ctx* ctx, const acm_camera* cam, int32_t kind, const double* r, acm_normal_equations* out) {
    memset(out, 0, sizeof(*out));
    out->n_params = cam->n_params;
    double cnt = 0.0;
    ACM_DISPATCH_LIN(cam->model, kind, (LinOps<M, KIND>::unpack(r, cam->params, out->H, out->g, &out->cost, &cnt)));
    out->n_valid = (uint64_t)cnt;
    return ACM_OK;
}

extern "C" int32_t acm_linearize_async(acm_ctx* ctx, const acm_camera* cam, int32_t residual_kind, const acm_points* xyz, const acm_points* uv) {
    if (!ctx) return ACM_ERR_INVALID_ARG;
    int32_t rc = check_lin_args(ctx, cam, xyz, uv);
    if (rc) return rc;
    int nacc = 0;
    return enqueue_linearize(ctx, cam, residual_kind, nullptr, xyz, uv, 0.0, &nacc, nullptr);
}

extern "C" int32_t acm_linearize(acm_ctx* ctx, const acm_camera* cam, int32_t residual_kind, const acm_points* xyz, const acm_points* uv,
                                 acm_normal_equations* out) {
    if (!ctx || !out) return ACM_ERR_INVALID_ARG;
    int32_t rc = check_lin_args(ctx, cam, xyz, uv);
    if (rc) return rc;
    int nacc = 0;
    rc = enqueue_linearize(ctx, cam, residual_kind, nullptr, xyz, uv, 0.0, &nacc, nullptr);
    if (rc) return rc;
    ACM_CUDA(ctx, cudaMemcpyAsync(ctx->h_reduce, ctx->d_reduce, nacc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ACM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return unpack_host(ctx, cam, residual_kind, ctx->h_reduce, out);
}

extern "C" int32_t acm_linearize_host(acm_ctx* ctx, const acm_camera* cam, int32_t residual_kind, const double* xyz_aos, const double* uv_aos,
                                      size_t n, acm_normal_equations* out) {
    if (!ctx || !out) return ACM_ERR_INVALID_ARG;
    ACM_REQUIRE(ctx, cam && (n == 0 || (xyz_aos && uv_aos)), "linearize_host: null argument");
    // device buffers are kept between calls (grow-only): a 4 GB cudaMalloc/cudaFree pair per call
    // would cost more than the kernel
    if (ctx->cache_cap < n || !ctx->cache3) {
        cudaStreamSynchronize(ctx->stream);
        acm_points_destroy(ctx, ctx->cache3); acm_points_destroy(ctx, ctx->cache2);
        ctx->cache3 = ctx->cache2 = nullptr; ctx->cache_cap = 0;
        int32_t rc0 = acm_points_create(ctx, 3, n, ACM_F64, &ctx->cache3);
        if (!rc0) rc0 = acm_points_create(ctx, 2, n, ACM_F64, &ctx->cache2);
        if (rc0) { acm_points_destroy(ctx, ctx->cache3); ctx->cache3 = nullptr; return rc0; }
        ctx->cache_cap = n;
    }
    acm_points* xyz = ctx->cache3; acm_points* uv = ctx->cache2;
    xyz->n = n; uv->n = n;  // views of the first n points (the component stride is unchanged)
    int32_t rc = acm_points_upload_any(ctx, xyz, xyz_aos, n, 0);
    if (!rc) rc = acm_points_upload_any(ctx, uv, uv_aos, n, 0);
    if (!rc) rc = acm_linearize(ctx, cam, residual_kind, xyz, uv, out);
    cudaStreamSynchronize(ctx->stream);
    xyz->n = ctx->cache_cap; uv->n = ctx->cache_cap;
    return rc;
}

extern "C" int32_t acm_lm_default_config(acm_lm_config* cfg) {
    if (!cfg) return ACM_ERR_INVALID_ARG;
    cfg->max_iterations = 100;        // bin/camera_converter.rs:411
    cfg->cost_tolerance = 1e-6;       // :412
    cfg->parameter_tolerance = 1e-8;  // :413
    cfg->gradient_tolerance = 1e-6;   // :414
    cfg->lambda0 = 1e-3;
    cfg->invalid_penalty = 0.0;
    cfg->check_every = 4;
    return ACM_OK;
}

extern "C" int32_t acm_lm_solve(acm_ctx* ctx, const acm_camera* init, int32_t residual_kind, const acm_points* xyz, const acm_points* uv,
                                const double* lower, const double* upper, const acm_lm_config* cfg_in, double* out_params,
                                acm_lm_result* result) {
    if (!ctx || !out_params || !result) return ACM_ERR_INVALID_ARG;
    int32_t rc = check_lin_args(ctx, init, xyz, uv);
    if (rc) return rc;
    acm_lm_config cfg;
    if (cfg_in) cfg = *cfg_in; else acm_lm_default_config(&cfg);
    if (cfg.check_every < 1) cfg.check_every = 1;
    const auto t_start = std::chrono::steady_clock::now();
    const int P = init->n_params;
    LmState* h = static_cast<LmState*>(ctx->h_lm);
    LmState* d = static_cast<LmState*>(ctx->d_lm);
    memset(h, 0, sizeof(LmState));
    for (int i = 0; i < P; ++i) {
        h->lower[i] = lower ? lower[i] : -INFINITY;
        h->upper[i] = upper ? upper[i] : INFINITY;
        double v = init->params[i];
        v = v < h->lower[i] ? h->lower[i] : (v > h->upper[i] ? h->upper[i] : v);
        h->x[i] = v; h->xt[i] = v;
    }
    h->lambda = cfg.lambda0; h->nu = 2.0;
    h->cost_tol = cfg.cost_tolerance; h->param_tol = cfg.parameter_tolerance; h->grad_tol = cfg.gradient_tolerance;
    h->max_iter = cfg.max_iterations; h->first = 1; h->P = P; h->status = 3;
    ACM_CUDA(ctx, cudaMemcpyAsync(d, h, sizeof(LmState), cudaMemcpyHostToDevice, ctx->stream));
    // the host copy doubles as the read-back buffer: wait until the upload has consumed it
    ACM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

    int nacc = 0;
    auto one_iteration = [&]() -> int32_t {
        int fuse = 0;
        int32_t r = enqueue_linearize(ctx, init, residual_kind, d, xyz, uv, cfg.invalid_penalty, &nacc, &fuse);
        if (r) return r;
        if (!fuse) {
            ACM_DISPATCH_LIN(init->model, residual_kind, (lm_step_kernel<M, KIND><<<1, 128, 0, ctx->stream>>>(d, ctx->d_reduce)));
            ACM_CHECK_LAUNCH(ctx);
        }
        return ACM_OK;
    };
    const int max_passes = cfg.max_iterations + 2;
    int enq = 0;
    bool done = false;
    while (!done && enq < max_passes) {
        for (int k = 0; k < cfg.check_every && enq < max_passes; ++k, ++enq) {
            rc = one_iteration();
            if (rc) return rc;
        }
        ACM_CUDA(ctx, cudaMemcpyAsync(h, d, sizeof(LmState), cudaMemcpyDeviceToHost, ctx->stream));
        ACM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        done = h->done != 0;
    }
    for (int i = 0; i < P; ++i) out_params[i] = h->x[i];
    result->status = h->status; result->iterations = h->iterations; result->passes = h->passes;
    result->initial_cost = h->initial_cost; result->final_cost = h->cost; result->n_valid = (uint64_t)h->n_valid;
    result->elapsed_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    return ACM_OK;
}

// ---------------------------------------------------------------------------------------
// project with `compute_jacobian = true` (README-era trait surface, reference README.md:119-126;
// per-model doc-comments double_sphere.rs:326-332 "2x6", kannala_brandt.rs:309-313 "2x8"):
// uv + the 2xP Jacobian w.r.t. [fx,fy,cx,cy,dist..], written as 2P rows of n doubles.
// ---------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256) project_jacobian_kernel(const __grid_constant__ CamParams c, LinParams p, const double* __restrict__ X,
                                                               const double* __restrict__ Y, const double* __restrict__ Z,
                                                               double* __restrict__ U, double* __restrict__ V, double* __restrict__ J,
                                                               uint8_t* __restrict__ S, size_t n) {
    using LM_ = Lin<M, ACM_RESIDUAL_PIXEL>;
    constexpr int ND = LM_::ND, P = 4 + ND;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = X[i], y = Y[i], z = Z[i];
        double u, v;
        int st = CamModel<M>::template project<false>(c, x, y, z, u, v);
        double ru, rv, au[2 + ND], av[2 + ND];
        double row_u[P], row_v[P];
#pragma unroll
        for (int k = 0; k < P; ++k) row_u[k] = row_v[k] = 0.0;
        if (st == ACM_POINT_OK && LM_::eval(p, x, y, z, 0.0, 0.0, ru, rv, au, av)) {
            row_u[0] = au[0]; row_u[2] = 1.0; row_v[1] = av[0]; row_v[3] = 1.0;
#pragma unroll
            for (int k = 0; k < ND; ++k) { row_u[4 + k] = au[2 + k]; row_v[4 + k] = av[2 + k]; }
        } else {
            if (st == ACM_POINT_OK) st = ACM_POINT_NUMERICAL_ERROR;
            u = v = acm_nan();
        }
        U[i] = u; V[i] = v;
        if (S) S[i] = (uint8_t)st;
#pragma unroll
        for (int k = 0; k < P; ++k) { J[(size_t)k * n + i] = row_u[k]; J[(size_t)(P + k) * n + i] = row_v[k]; }
    }
}

extern "C" int32_t acm_project_jacobian(acm_ctx* ctx, const acm_camera* cam, const acm_points* xyz, acm_points* uv, double* d_jac,
                                        uint8_t* d_status) {
    if (!ctx) return ACM_ERR_INVALID_ARG;
    ACM_REQUIRE(ctx, cam && xyz && uv && d_jac, "project_jacobian: null argument");
    ACM_REQUIRE(ctx, xyz->dim == 3 && uv->dim == 2 && xyz->n == uv->n, "project_jacobian: shape mismatch");
    ACM_REQUIRE(ctx, xyz->dtype == ACM_F64 && uv->dtype == ACM_F64, "project_jacobian: f64 buffers required");
    CamParams c;
    int32_t rc = acm_make_cam_params(ctx, cam, &c);
    if (rc) return rc;
    LinParams p;
    make_lin_params(cam, &p);
    const size_t n = xyz->n;
    if (n == 0) return ACM_OK;
    int grid = grid_for(ctx, n, 256, 4);
    ACM_DISPATCH_MODEL(cam->model, (project_jacobian_kernel<M><<<grid, 256, 0, ctx->stream>>>(
        c, p, comp<double>(xyz, 0), comp<double>(xyz, 1), comp<double>(xyz, 2), comp<double>(uv, 0), comp<double>(uv, 1), d_jac, d_status, n)))
    ACM_CHECK_LAUNCH(ctx);
    return ACM_OK;
}

// ---------------------------------------------------------------------------------------
// project with the 2x3 Jacobian w.r.t. the 3-D point (the other reading of the README-era
// `compute_jacobian` flag: trait doc reference src/camera/mod.rs:246-252 "Jacobian matrix (2x3)").
// uv + six rows of n doubles: du/dx, du/dy, du/dz, dv/dx, dv/dy, dv/dz.
// ---------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256) project_point_jacobian_kernel(const __grid_constant__ CamParams c, LinParams p, const double* __restrict__ X,
                                                                     const double* __restrict__ Y, const double* __restrict__ Z,
                                                                     double* __restrict__ U, double* __restrict__ V, double* __restrict__ J,
                                                                     uint8_t* __restrict__ S, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = X[i], y = Y[i], z = Z[i];
        double u, v, ju[3], jv[3];
        const int st = CamModel<M>::template project<false>(c, x, y, z, u, v);
        if (st == ACM_POINT_OK) PointJac<M>::eval(p, x, y, z, ju, jv);
        else { u = v = acm_nan(); ju[0] = ju[1] = ju[2] = jv[0] = jv[1] = jv[2] = 0.0; }
        U[i] = u; V[i] = v;
        if (S) S[i] = (uint8_t)st;
#pragma unroll
        for (int k = 0; k < 3; ++k) { __stcs(J + (size_t)k * n + i, ju[k]); __stcs(J + (size_t)(3 + k) * n + i, jv[k]); }
    }
}

extern "C" int32_t acm_project_point_jacobian(acm_ctx* ctx, const acm_camera* cam, const acm_points* xyz, acm_points* uv, double* d_jac,
                                              uint8_t* d_status) {
    if (!ctx) return ACM_ERR_INVALID_ARG;
    ACM_REQUIRE(ctx, cam && xyz && uv && d_jac, "project_point_jacobian: null argument");
    ACM_REQUIRE(ctx, xyz->dim == 3 && uv->dim == 2 && xyz->n == uv->n, "project_point_jacobian: shape mismatch");
    ACM_REQUIRE(ctx, xyz->dtype == ACM_F64 && uv->dtype == ACM_F64, "project_point_jacobian: f64 buffers required");
    CamParams c;
    int32_t rc = acm_make_cam_params(ctx, cam, &c);
    if (rc) return rc;
    LinParams p;
    make_lin_params(cam, &p);
    const size_t n = xyz->n;
    if (n == 0) return ACM_OK;
    int grid = grid_for(ctx, n, 256, 4);
    ACM_DISPATCH_MODEL(cam->model, (project_point_jacobian_kernel<M><<<grid, 256, 0, ctx->stream>>>(
        c, p, comp<double>(xyz, 0), comp<double>(xyz, 1), comp<double>(xyz, 2), comp<double>(uv, 0), comp<double>(uv, 1), d_jac, d_status, n)))
    ACM_CHECK_LAUNCH(ctx);
    return ACM_OK;
}
