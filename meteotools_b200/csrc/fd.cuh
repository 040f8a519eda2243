// Finite-difference formulas of calc/differential.py, written once and shared by the generic
// mt_fd kernel and the fused diagnostic kernels.  `v(j)` returns element j along the
// differentiated axis; the evaluation order is NumPy's left-to-right order so that every
// rounding matches (e.g. (-3*v0 + 4*v1 - v2)/2/delta; note "/2/delta", not "/(2*delta)").
#pragma once
#include <string.h>

#include "common.cuh"

namespace mt {

// EXACT division by a divisor that is fixed for a whole kernel / thread (grid spacings, the radius
// of a thread's column): with rd = RN(1/d),  q0 = RN(a*rd),  r = a - q0*d (exact, one FMA),
// q = RN(q0 + r*rd)  is the correctly rounded a/d (Markstein's correction step -- the tail of the
// IEEE division sequence without the ~15 instructions that build the reciprocal).  Verified bit for
// bit against a/d on 352 000 (a, d) pairs with exact rational arithmetic (profiles/experiments_r01.md)
// and by every bit-exact FD / cyl_d / car_d parity test.  Three FP64 instructions + an INTEGER test of
// q0's exponent (the FP64 pipe is the scarce one): quotients that are huge, infinite, zero or close to
// the subnormal range, and divisors that are not normal numbers (r[0] = 0!), leave the fast path --
// a zero dividend returns q0 = +-0 (the sign IEEE gives), everything else takes the IEEE division;
// NaN flows through the fast path as NaN (terrain-masked lanes stay branch-free).
struct CDiv {
    double d, rd;
    bool plain;
    __device__ __forceinline__ explicit CDiv(double d_) : CDiv(d_, 1.0 / d_) {}
    // rd_ = RN(1 / d_) computed by the caller (on the host: the same IEEE division)
    __device__ __forceinline__ CDiv(double d_, double rd_) : d(d_), rd(rd_)
    {
        const unsigned h = (unsigned)__double2hiint(d_) & 0x7fffffffu;
        plain = h - 0x00100000u >= 0x7fe00000u;   // zero, denormal, inf, NaN
    }
    __device__ __forceinline__ double operator()(double a) const
    {
        if (plain) return a / d;
        const double q0 = a * rd;
        const double r = fma(-q0, d, a);
        double q = fma(r, rd, q0);
        // biased exponent of q0 in [60, 2020]  <=>  2^-963 <= |q0| < 2^998: the residual is exact
        const unsigned e = ((unsigned)__double2hiint(q0) >> 20) & 0x7ffu;
        if (e - 60u > 1960u) {
            if (a == 0.0) q = q0;            // +-0 / d
            else if (q0 == q0) q = a / d;    // tiny, huge, infinite; NaN keeps the fast path's NaN
        }
        return q;
    }
};

// The same quotient WITHOUT a branch, for the straight-line stencil body of csrc/diag_tma.cu.  The IEEE division
// the compiler emits rebuilds the reciprocal for every quotient (MUFU.RCP64H + 4 DFMA) and wraps each one in a
// convergence region around its slow path, which keeps the scheduler from overlapping the ~100-cycle dependent
// chains of neighbouring divisions; CDiv's per-quotient branch does the same.  Here the three-instruction tail
// always runs and an integer test of the dividend's exponent SELECTS between it and q0 = a * rd:
//   a = +-0, NaN or +-inf, or a divisor that is 0 or +-inf (r[0] = 0 on the axis of a cylindrical grid: rd =
//   +-inf)  ->  q0, which IS the IEEE result in all of these cases (sign of zero, 0/0 = NaN, x/0 = +-inf included);
//   every other (normal) dividend  ->  the corrected quotient, bit-identical to a / d.
// Not covered, and outside anything a model field holds: subnormal dividends or divisors (q0 may be 1 ulp off),
// quotients in the subnormal range, a divisor whose significand is all ones (Markstein's exception).  The host
// checks the grid spacings (fdiv_divisor_ok) and falls back to the IEEE kernels otherwise.
struct FDiv {
    double d, rd;
    bool special;   // the divisor is 0, subnormal, inf or NaN
    __device__ __forceinline__ explicit FDiv(double d_) : d(d_), rd(1.0 / d_)
    {
        special = ((((unsigned)__double2hiint(d_) + 0x00100000u) & 0x7fe00000u) == 0u);
    }
    __device__ __forceinline__ double operator()(double a) const
    {
        const double q0 = a * rd;
        const double q = fma(fma(-q0, d, a), rd, q0);
        // biased exponent of a is 0 (zero, subnormal) or 0x7ff (inf, NaN)  <=>  ((e + 1) & 0x7fe) == 0
        const bool sp = special || ((((unsigned)__double2hiint(a) + 0x00100000u) & 0x7fe00000u) == 0u);
        return sp ? q0 : q;
    }
};
// a divisor FDiv may take from the host: a normal number of moderate size whose significand is not all ones
inline bool fdiv_divisor_ok(double d)
{
    unsigned long long b;
    static_assert(sizeof(b) == sizeof(d), "double is 64 bits");
    memcpy(&b, &d, sizeof(b));
    const unsigned e = (unsigned)((b >> 52) & 0x7ffu);
    const unsigned long long m = b & 0xfffffffffffffULL;
    return e >= 1023u - 200u && e <= 1023u + 200u && m != 0xfffffffffffffULL;
}

// the stencil formulas below divide through `by(x, delta)`: a plain double divisor gives the IEEE
// division (fused diagnostic kernels: measured no gain from CDiv there, they are latency bound),
// a CDiv the exact reciprocal-and-correct form (mt_fd: 0.386 -> 0.297 ms on 80 x 1000 x 1000)
__device__ __forceinline__ double by(double a, double d) { return a / d; }
__device__ __forceinline__ double by(double a, const CDiv &d) { return d(a); }
__device__ __forceinline__ double by(double a, const FDiv &d) { return d(a); }

// FD2, differential.py:30-34.  The axis segment [lo, hi) is what the caller may READ (halo
// levels of a level shard included); `first`/`last` tell whether index 0 / n-1 of the segment
// are the GLOBAL ends (one-sided formulas) or interior points of a shard (centred).
template <class V, class D>
__device__ __forceinline__ double fd2(V v, int i, int n, D delta)
{
    if (i == 0) return by((-3 * v(0) + 4 * v(1) - v(2)) / 2, delta);
    if (i == n - 1) return by((3 * v(n - 1) - 4 * v(n - 2) + v(n - 3)) / 2, delta);
    return by((v(i + 1) - v(i - 1)) / 2, delta);
}

// level-shard aware variant: i is a local index in an array of n levels that holds halo levels;
// bottom/top say whether local 0 / n-1 are the global first / last level.
template <class V, class D>
__device__ __forceinline__ double fd2_shard(V v, int i, int n, D delta, bool bottom, bool top)
{
    if (i == 0 && bottom) return by((-3 * v(0) + 4 * v(1) - v(2)) / 2, delta);
    if (i == n - 1 && top) return by((3 * v(n - 1) - 4 * v(n - 2) + v(n - 3)) / 2, delta);
    return by((v(i + 1) - v(i - 1)) / 2, delta);
}

template <class V, class D>
__device__ __forceinline__ double fd4(V v, int i, int n, D delta)  // differential.py:66-72
{
    if (i == 0) return by((-3 * v(0) + 4 * v(1) - v(2)) / 2, delta);
    if (i == 1) return by((v(2) - v(0)) / 2, delta);
    if (i == n - 1) return by((3 * v(n - 1) - 4 * v(n - 2) + v(n - 3)) / 2, delta);
    if (i == n - 2) return by((v(n - 1) - v(n - 3)) / 2, delta);
    return by((-v(i + 2) + 8 * v(i + 1) - 8 * v(i - 1) + v(i - 2)) / 12, delta);
}

template <class V, class D>
__device__ __forceinline__ double fd2_front(V v, int i, int n, D delta)  // :92-94
{
    if (i >= n - 2) return nan_f64();
    return by((-3 * v(i) + 4 * v(i + 1) - v(i + 2)) / 2, delta);
}

template <class V, class D>
__device__ __forceinline__ double fd2_back(V v, int i, int n, D delta)  // :113-115
{
    if (i < 2) return nan_f64();
    return by((3 * v(i) - 4 * v(i - 1) + v(i - 2)) / 2, delta);
}

template <class V, class D>
__device__ __forceinline__ double fd2_2(V v, int i, int n, D delta)  // :134-138
{
    if (i == 0) return by(by((2 * v(0) - 5 * v(1) + 4 * v(2) - v(3)), delta), delta);
    if (i == n - 1) return by(by((2 * v(n - 1) - 5 * v(n - 2) + 4 * v(n - 3) - v(n - 4)), delta), delta);
    return by(by((v(i + 1) - 2 * v(i) + v(i - 1)), delta), delta);
}

template <class V, class D>
__device__ __forceinline__ double fd2_2_front(V v, int i, int n, D delta)  // :157-159
{
    if (i >= n - 3) return nan_f64();
    return by(by((2 * v(i) - 5 * v(i + 1) + 4 * v(i + 2) - v(i + 3)), delta), delta);
}

template <class V, class D>
__device__ __forceinline__ double fd2_2_back(V v, int i, int n, D delta)  // :178-180
{
    if (i < 3) return nan_f64();
    return by(by((2 * v(i) - 5 * v(i - 1) + 4 * v(i - 2) - v(i - 3)), delta), delta);
}

}  // namespace mt
