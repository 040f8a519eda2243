// Lean double-precision log / exp for the FP64-issue-bound thermodynamic kernels (a12).
//
// CUDA's log/exp cost ~85 / ~65 SASS instructions each on sm_100a, more than half of them moves of
// 64-bit polynomial constants (UMOV pairs: DFMA takes no 64-bit immediate) and special-case
// handling.  These versions are table driven (2.5 KB of shared memory), ~20 instructions each:
//   log x : x = 2^e m, i = top 7 mantissa bits, r_i ~ 1/m from the table, u = fma(m, r_i, -1)
//           (|u| < 2^-8, exact to one rounding), log x = e ln2 - log r_i + log1p(u), degree-6 series;
//   exp x : k = rint(64 x / ln2), r = x - k ln2/64 (two-term Cody-Waite), exp x = 2^(k>>6) *
//           2^((k&63)/64) * (1 + expm1 series of degree 5).
// Accuracy (measured against NumPy, profiles/experiments_r01.md): log <= 2 ulp for arguments away
// from 1 (absolute error <= 3.6e-15 everywhere), exp <= 1.3e-15 relative on [-30, 10]: the outputs
// that use them are tolerance-class (1e-10, north star).  NaN in -> NaN out WITHOUT a branch
// (terrain-masked points are common); zero, negative, denormal, infinite or |x| >= 690 arguments
// take CUDA's own function on a rarely taken branch.
#pragma once
#include "common.cuh"
#include "fastmath_tables.inc"

namespace mt {
namespace fm {

struct Tables {
    double2 lg[128];   // (r_i, -log r_i)
    double ex[64];     // 2^(j/64)
};

// 64-bit constants that are not encodable as an immediate (non-zero low word) cost two UMOVs per
// use on sm_100a -- a quarter of the fused cyl_td kernel's instructions were such moves (SASS of
// k_td_main).  Read from constant memory they are one LDC.64 / half an LDCU.128 each, and the
// compiler may keep them in uniform registers across the loop.
struct Consts {
    double m16, p02, p13, ln2;                       // log1p series -1/6, 1/5, 1/3; ln 2
    double r64ln2, ln2_64_hi, ln2_64_lo;             // 64/ln2, ln2/64 split (Cody-Waite)
    double e120, e24, e6;                            // expm1 series 1/120, 1/24, 1/6
};
static __constant__ Consts kc = {-1.0 / 6,       0.2,  1.0 / 3, MT_FM_LN2, MT_FM_64_LN2, MT_FM_LN2_64_HI,
                                 MT_FM_LN2_64_LO, 1.0 / 120, 1.0 / 24, 1.0 / 6};

static __device__ const double2 g_log_table[128] = {MT_FM_LOG_TABLE};
static __device__ const double g_exp_table[64] = {MT_FM_EXP_TABLE};

// cooperative copy into shared memory; the caller synchronises the block afterwards
__device__ __forceinline__ void load_tables(Tables *s)
{
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s->lg[i] = g_log_table[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s->ex[i] = g_exp_table[i];
}

// out-of-line slow paths: taken by (almost) no thread, they must not bloat the hot loop
static __device__ __noinline__ double slow_log(double x) { return log(x); }
static __device__ __noinline__ double slow_exp(double x) { return exp(x); }

// Two forms of each function.  The plain form handles its rare arguments itself (a branch to
// CUDA's function / the IEEE division).  The `_f` form never branches: it ORs "this argument needs
// the slow path" into `bad` and returns the fast-path value; a caller that evaluates a whole chain
// of them (the fused cyl_td pass: ~24 calls per point) tests `bad` ONCE and redoes the point with
// the plain forms -- one branch per point instead of one branch region per call.
template <bool FLAG>
__device__ __forceinline__ double flog_t(double x0, const Tables *t, bool &bad)
{
    // NaN lanes (terrain) compute on 1.0 and get their NaN back at the end.  Without this the
    // compiler folds "x < DBL_MIN || x == inf" and the final NaN select into ONE integer range test
    // that is also true for NaN, and every warp holding a masked point runs the slow path
    // (measured: 19 % of the kernel's instructions).
    const bool isn = x0 != x0;
    const double x = isn ? 1.0 : x0;
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double2 rt = t->lg[(hi >> 13) & 127];
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double e = (double)((hi >> 20) - 1023);
    const double u = fma(m, rt.x, -1.0);
    double q = fma(u, kc.m16, kc.p02);
    q = fma(u, q, -0.25);
    q = fma(u, q, kc.p13);
    q = fma(u, q, -0.5);
    double res = fma(e, kc.ln2, rt.y) + fma(u * u, q, u);
    const bool rare = (unsigned)(hi - 0x00100000) >= 0x7fe00000u;   // <= 0, denormal, +inf
    if (FLAG) bad = bad || rare;
    else if (rare) res = slow_log(x);
    return isn ? x0 : res;
}

template <bool FLAG>
__device__ __forceinline__ double fexp_t(double x, const Tables *t, bool &bad)
{
    const double kd0 = fma(x, kc.r64ln2, 6755399441055744.0);   // 1.5 * 2^52: rint in the low word
    const int k = __double2loint(kd0);
    const double kd = kd0 - 6755399441055744.0;
    double r = fma(-kd, kc.ln2_64_hi, x);
    r = fma(-kd, kc.ln2_64_lo, r);
    const double ej = t->ex[k & 63];
    double p = fma(r, kc.e120, kc.e24);
    p = fma(r, p, kc.e6);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    double res = fma(ej, p * r, ej);
    res = __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
    const bool rare = fabs(x) >= 690.0;                            // overflow / underflow range
    if (FLAG) bad = bad || rare;
    else if (rare) res = slow_exp(x);
    return (x != x) ? x : res;
}

// a / b for tolerance-class quotients: MUFU.RCP64H seed (~20 bits) + two Newton steps + one multiply
// (~8 instructions instead of ~20 for the IEEE division sequence), <= 2 ulp.  Denominators that are
// zero, denormal, huge or infinite take the IEEE division on a rarely taken branch; NaN propagates.
static __device__ __noinline__ double slow_div(double a, double b) { return a / b; }

template <bool FLAG>
__device__ __forceinline__ double fdiv_t(double a, double b, bool &bad)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    double res = a * y;
    // |b| outside [2^-1000, 2^992) incl. 0 and inf -- but not (quiet) NaN, which flows through the
    // fast path as NaN: terrain-masked lanes must not drag their warp into the slow path
    const unsigned hb = (unsigned)__double2hiint(b) & 0x7fffffffu;
    const bool rare = hb - 0x01700000u >= 0x7c800000u && hb <= 0x7ff00000u;
    if (FLAG) bad = bad || rare;
    else if (rare) res = slow_div(a, b);
    return res;
}

__device__ __forceinline__ double flog(double x, const Tables *t)
{
    bool b = false;
    return flog_t<false>(x, t, b);
}
__device__ __forceinline__ double fexp(double x, const Tables *t)
{
    bool b = false;
    return fexp_t<false>(x, t, b);
}
__device__ __forceinline__ double fdiv(double a, double b)
{
    bool f = false;
    return fdiv_t<false>(a, b, f);
}

}  // namespace fm
}  // namespace mt
