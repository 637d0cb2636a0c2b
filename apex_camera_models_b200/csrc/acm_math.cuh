// Branch-free f64 helpers shared by every translation unit.  Written with explicit __fma_rn /
// __dmul_rn / __dadd_rn intrinsics so that they compile to the same instructions with and without
// -fmad=false.
#pragma once
#include <cuda_runtime.h>

// 1/a for normal finite a: MUFU seed (~2^-23) + two Newton steps => <= 1 ulp.
__device__ __forceinline__ double acm_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = __fma_rn(-a, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-a, r, 1.0);
    return __fma_rn(r, e, r);
}

// 1/sqrt(a) for normal finite a > 0: MUFU seed + two Newton steps => <= 2 ulp.
__device__ __forceinline__ double acm_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = __dmul_rn(0.5, a);
    y = __dmul_rn(y, __fma_rn(-h, __dmul_rn(y, y), 1.5));
    return __dmul_rn(y, __fma_rn(-h, __dmul_rn(y, y), 1.5));
}

// sqrt(a) (<= 1 ulp) and 1/sqrt(a) (<= 2 ulp) together: two coupled Goldschmidt iterations on
// g ~ sqrt(a), h ~ 1/(2 sqrt(a)):  r = 1/2 - g h,  g += g r,  h += h r.  2 DMUL + 6 DFMA + 1 DADD.
// a = 0 returns 0 (inv = NaN is never used then); a < 0 and NaN give NaN like sqrt().
__device__ __forceinline__ double acm_sqrt_inv(double a, double& inv) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = __dmul_rn(a, y), h = __dmul_rn(0.5, y);
    double r = __fma_rn(-g, h, 0.5);
    g = __fma_rn(g, r, g); h = __fma_rn(h, r, h);
    r = __fma_rn(-g, h, 0.5);
    g = __fma_rn(g, r, g); h = __fma_rn(h, r, h);
    inv = __dadd_rn(h, h);
    return a == 0.0 ? 0.0 : g;
}

// Correctly rounded a / b from the correctly rounded reciprocal ib = RN(1/b) (Markstein: q = RN(a ib)
// is within 1 ulp of a/b, the fma residual r = a - q b is exact, and RN(q + r ib) is then the correctly
// rounded quotient): 3 FP64 instructions and no slow-path call instead of ~10 + a branch.  The result is
// BIT-IDENTICAL to a / b, so it may feed the bit-exact validity tests.  Valid while nothing under- or
// overflows: the caller guarantees 2^-100 <= |b| <= 2^100 (host flag CamParams::fast_div, or
// acm_exp_ok(b)); numerators outside [2^-895, 2^897] -- including 0, whose sign the fma chain would
// lose -- take the IEEE division (an integer test on the exponent field, off the FP64 pipe).
// Checked against a / b on 1e9 random and adversarial (near-tie, all-ones mantissa) operand pairs on the CPU.
__device__ __forceinline__ bool acm_exp_ok(double b) {  // 2^-100 <= |b| < 2^101
    const unsigned e = ((unsigned)__double2hiint(b) >> 20) & 0x7ffu;
    return e - 923u <= 200u;
}
__device__ __forceinline__ double acm_div_by(double a, double b, double ib) {
    const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
    if (e - 128u >= 1792u) return a / b;
    const double q = __dmul_rn(a, ib);
    const double r = __fma_rn(-q, b, a);
    return __fma_rn(r, ib, q);
}

// sin and cos for 0 <= x <= ~1.8 (the Newton root of the Kannala-Brandt unprojection): x > pi/4 is reflected to
// y = pi/2 - x (two-term pi/2), then the classic minimax kernels on [-pi/4, pi/4] (degree 13 / 14, < 1 ulp).
// ~24 FP64 instructions, no branch, no slow path; used only behind the last validity test.
static __constant__ double ACM_SINCOS_C[12] = {
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
__device__ __forceinline__ void acm_sincos_small(double x, double& s, double& c) {
    const bool refl = x > 0.78539816339744828;
    const double y = refl ? __dadd_rn(__dsub_rn(1.5707963267948966, x), 6.123233995736766e-17) : x;
    const double z = __dmul_rn(y, y);
    double ps = ACM_SINCOS_C[5], pc = ACM_SINCOS_C[11];
#pragma unroll
    for (int k = 4; k >= 0; --k) { ps = __fma_rn(ps, z, ACM_SINCOS_C[k]); pc = __fma_rn(pc, z, ACM_SINCOS_C[6 + k]); }
    const double sy = __fma_rn(__dmul_rn(y, z), ps, y);                                   // y + y^3 S(z)
    const double cy = __fma_rn(__dmul_rn(z, z), pc, __fma_rn(-0.5, z, 1.0));              // 1 - z/2 + z^2 C(z)
    s = refl ? cy : sy;
    c = refl ? sy : cy;
}

// atan2(a, b) for a >= 0, b > 0 (first quadrant: all that the fisheye models need).  Two argument
// reductions share ONE reciprocal -- swap so that t = num/den <= 1, then
// atan(t) = pi/4 + atan((num-den)/(num+den)) above tan(pi/8) -- leaving |t| <= sqrt(2)-1, where a
// degree-10 polynomial in t^2 (Chebyshev-node fit computed with mpmath at 60 digits, approximation
// error 6.9e-17 relative) is evaluated by Horner.  Total error <= ~1.5 ulp, like the CUDA library's
// atan2 (2 ulp), in ~25 FP64 instructions with no branch and no slow-path call.
// Operands whose magnitude would push the MUFU seed into its flush-to-zero range take the library
// function (a branch that is practically never taken; GUARD = false drops it for the solver kernels,
// whose streaming loop must stay one basic block and whose inputs are sane by construction).
// The coefficients live in the constant bank: DFMA then reads them as c[bank][offset] operands,
// whereas 64-bit literals are re-materialised (two UMOV / IMAD.MOV each) inside the streaming loops.
static __constant__ double ACM_ATAN_C[14] = {
    2.11353731576932463e-02, -4.34805221571646222e-02, 5.68834922680901064e-02, -6.64023393042940807e-02,
    7.68995349630685748e-02, -9.09077307480841423e-02, 1.11111061804559458e-01, -1.42857141809764665e-01,
    1.99999999988551114e-01, -3.33333333333284410e-01,
    0.41421356237309503 /* tan(pi/8) */, 0.78539816339744828 /* pi/4 */, 1.5707963267948966 /* pi/2 */, 0.0};

template <bool GUARD = true>
__device__ __forceinline__ double acm_atan2_q1(double a, double b) {
    const bool swap = a > b;
    const double num = swap ? b : a, den = swap ? a : b;
    if (GUARD && !(den > 1e-280 && den < 1e280)) return atan2(a, b);
    const bool hi = num > __dmul_rn(ACM_ATAN_C[10], den);
    const double n2 = hi ? __dsub_rn(num, den) : num;
    const double d2 = hi ? __dadd_rn(num, den) : den;
    const double t = __dmul_rn(n2, acm_rcp(d2));
    const double s = __dmul_rn(t, t);
    // Horner: an Estrin split (depth 4 instead of 9, +2 instructions) measured 4 % SLOWER in the KB / FOV linearize
    // kernels (same-box A/B): with two points per thread and 12 warps per SM the chain is already hidden.
    double q = ACM_ATAN_C[0];
#pragma unroll
    for (int k = 1; k < 10; ++k) q = __fma_rn(q, s, ACM_ATAN_C[k]);
    double at = __fma_rn(__dmul_rn(t, s), q, t);
    at = hi ? __dadd_rn(ACM_ATAN_C[11], at) : at;
    return swap ? __dsub_rn(ACM_ATAN_C[12], at) : at;
}

// ---------------------------------------------------------------------------------------
// Table-driven atan2 for the solver kernels (first quadrant, a >= 0, b > 0), 16 FP64 instructions instead of 25 and a
// dependent chain of ~12 instead of ~20:  t = min/max in [0, 1];  with c = i/64 the table point nearest to t,
//   atan(t) = atan(c) + atan(x),   x = (t - c) / (1 + t c) = (min - c max) / (max + c min),   |x| <= 1/128 (+ index slack),
// so ONE reciprocal serves the division and the reduction, and x - x^3/3 + x^5/5 - x^7/7 is exact to 1.6e-18 relative
// (checked with mpmath over 2e5 random t).  The index comes from a float estimate of t built from the operands' high
// words (ALU / FP32 / MUFU work, nothing on the FP64 pipe); a wrong neighbour only widens |x| slightly, and i = 0 is exact
// for tiny t (x = min / max).  The table {c_i, atan(c_i)} (65 x 16 B) sits in shared memory: per-lane indices would
// serialise in the constant cache.  Total error <= ~1.5 ulp like acm_atan2_q1.
// ---------------------------------------------------------------------------------------
static __device__ const double ACM_ATAN_TAB64[65] = {
    0x0.0p+0, 0x1.fff555bbb729bp-7, 0x1.ffd55bba97625p-6, 0x1.7fb818430da2ap-5, 0x1.ff55bb72cfdeap-5, 0x1.3f59f0e7c559dp-4, 0x1.7ee182602f10fp-4, 0x1.be39ebe6f07c3p-4, 0x1.fd5ba9aac2f6ep-4, 0x1.1e1fafb043727p-3, 0x1.3d6eee8c6626cp-3, 0x1.5c9811e3ec26ap-3, 0x1.7b97b4bce5b02p-3, 0x1.9a6a8e96c8626p-3, 0x1.b90d7529260a2p-3, 0x1.d77d5df205736p-3, 0x1.f5b75f92c80ddp-3, 0x1.09dc597d86362p-2, 0x1.18bf5a30bf178p-2, 0x1.278372057ef46p-2, 0x1.362773707ebccp-2, 0x1.44aa436c2af0ap-2, 0x1.530ad9951cd4ap-2, 0x1.614840309cfe2p-2, 0x1.6f61941e4def1p-2, 0x1.7d5604b63b3f7p-2, 0x1.8b24d394a1b25p-2, 0x1.98cd5454d6b18p-2, 0x1.a64eec3cc23fdp-2, 0x1.b3a911da65c6cp-2, 0x1.c0db4c94ec9f0p-2, 0x1.cde53432c1351p-2, 0x1.dac670561bb4fp-2, 0x1.e77eb7f175a34p-2, 0x1.f40dd0b541418p-2, 0x1.0039c73c1a40cp-1, 0x1.0657e94db30d0p-1, 0x1.0c6145b5b43dap-1, 0x1.1255d9bfbd2a9p-1, 0x1.1835a88be7c13p-1, 0x1.1e00babdefeb4p-1, 0x1.23b71e2cc9e6ap-1, 0x1.2958e59308e31p-1, 0x1.2ee628406cbcap-1, 0x1.345f01cce37bbp-1, 0x1.39c391cd4171ap-1, 0x1.3f13fb89e96f4p-1, 0x1.445065b795b56p-1, 0x1.4978fa3269ee1p-1, 0x1.4e8de5bb6ec04p-1, 0x1.538f57b89061fp-1, 0x1.587d81f732fbbp-1, 0x1.5d58987169b18p-1, 0x1.6220d115d7b8ep-1, 0x1.66d663923e087p-1, 0x1.6b798920b3d99p-1, 0x1.700a7c5784634p-1, 0x1.748978fba8e0fp-1, 0x1.78f6bbd5d315ep-1, 0x1.7d528289fa093p-1, 0x1.819d0b7158a4dp-1, 0x1.85d69576cc2c5p-1, 0x1.89ff5ff57f1f8p-1, 0x1.8e17aa99cc05ep-1, 0x1.921fb54442d18p-1};
#define ACM_ATAN_TAB_BYTES (65 * 16)

// fill the shared-memory table (call with all threads of the block, then __syncthreads())
__device__ __forceinline__ void acm_atan_tab_init(double* tab_smem) {
    for (int i = threadIdx.x; i < 65; i += blockDim.x) { tab_smem[2 * i] = (double)i * 0.015625; tab_smem[2 * i + 1] = ACM_ATAN_TAB64[i]; }
}

// float with ~20 correct bits from the high word of a non-negative double (0 below 2^-126, 3e38 above 2^127)
__device__ __forceinline__ float acm_approx_f32(double v) {
    const unsigned hi = (unsigned)__double2hiint(v);
    const float f = __uint_as_float((hi << 3) - 0x38000000u);
    return hi < 0x38100000u ? 0.0f : (hi >= 0x47f00000u ? 3.0e38f : f);
}

__device__ __forceinline__ double acm_atan2_q1_tab(double a, double b, unsigned tab_smem_addr) {
    const bool swap = a > b;
    const double mn = swap ? b : a, mx = swap ? a : b;
    int i = __float2int_rn(__fdividef(acm_approx_f32(mn), acm_approx_f32(mx)) * 64.0f);
    i = min(max(i, 0), 64);
    double c, ac;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c), "=d"(ac) : "r"(tab_smem_addr + 16u * (unsigned)i));
    const double num = __fma_rn(-c, mx, mn), den = __fma_rn(c, mn, mx);
    const double x = __dmul_rn(num, acm_rcp(den));
    const double x2 = __dmul_rn(x, x);
    double p = __fma_rn(x2, -1.0 / 7.0, 0.2);
    p = __fma_rn(x2, p, -1.0 / 3.0);
    const double at = __dadd_rn(ac, __fma_rn(__dmul_rn(x, x2), p, x));
    return swap ? __dsub_rn(1.5707963267948966, at) : at;
}
