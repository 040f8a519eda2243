// Point-wise thermodynamics: single-op kernels behind mt_pointwise, the
// condensation-temperature fixed point (mt_calc_tc) and the fused cyl_td set (mt_cyl_td, "F4").
//
// calc_Tc (calc/thermaldynamic.py:392-402) stops when the GLOBAL np.max(|Tc-Tc2|) < 0.1, a
// NaN-sensitive test.  It is reproduced without any host round trip: iteration i reduces its
// max |diff| (as the bit pattern of a non-negative double, for which unsigned integer order ==
// floating order and every NaN sorts above +inf) into scratch[i]; iteration i+1 starts by
// reading scratch[i] and returns immediately when it is < 0.1 (or was never written: 0).
#include <cooperative_groups.h>
#include <cstdlib>

#include "fastmath.cuh"
#include "thermo.cuh"

namespace {

using namespace mt::th;

template <int OP>
__device__ __forceinline__ double apply(double a, double b, double c)
{
    switch (OP) {
        case MT_OP_THETA: return theta(a, b);
        case MT_OP_T: return T_from_theta(a, b);
        case MT_OP_SAT_VAPOR: return sat_vapor(a);
        case MT_OP_VAPOR: return b * sat_vapor(a);                        // :134
        case MT_OP_SAT_QV: return sat_qv_from_es(sat_vapor(a), b);
        case MT_OP_QV_TO_VAPOR: return qv_to_vapor(a, b);
        case MT_OP_THETA_ES: return exp_factor(theta(a, b), sat_qv_from_es(sat_vapor(a), b), a);  // :187-189
        case MT_OP_TD: return Td_from_vapor(a);
        case MT_OP_QV: return sat_qv_from_es(b, a);                       // 0.622*vapor/(P-vapor), :256
        case MT_OP_THETA_V: return theta_v(a, b, c);
        case MT_OP_TV_QV: return Tv_from_qv(a, b);
        case MT_OP_TV_VAPOR: return Tv_from_vapor(a, b, c);
        case MT_OP_RHO: return rho(a, b);
        case MT_OP_THETA_E: return exp_factor(a, b, c);
        case MT_OP_MOIST_DTDZ: return moist_dTdz(a, b);
        case MT_OP_RH: return a / sat_vapor(b);                           // :529-530
        case MT_OP_P_TO_Z: return log(b / a) * (Rd * c / g);              // :33-34: log(P0/P)*H
        case MT_OP_Z_TO_P: return b * exp(-a / (Rd * c / g));             // :51-52
        case MT_OP_ADD: return a + b;
    }
    return mt::nan_f64();
}

template <int OP>
__global__ void __launch_bounds__(256)
k_pointwise(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c,
            double sb, double sc, double *__restrict__ out, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double va = __ldg(a + i);
        const double vb = b ? __ldg(b + i) : sb;
        const double vc = c ? __ldg(c + i) : sc;
        out[i] = apply<OP>(va, vb, vc);
    }
}

// ---- block max of |diff| bit patterns ---------------------------------------------------------
__device__ __forceinline__ void block_max_to(unsigned long long v, unsigned long long *dst)
{
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    __shared__ unsigned long long smax[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smax[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < (int)(blockDim.x >> 5)) ? smax[lane] : 0ull;
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
            v = other > v ? other : v;
        }
        if (lane == 0 && v != 0ull) atomicMax(dst, v);
    }
}

__device__ __forceinline__ unsigned long long absbits(double d)
{
    return (unsigned long long)__double_as_longlong(fabs(d));
}

// g[i] = log(A*0.622/P/qv) + (Cp/Rd)*log(T), the sweep-invariant part (thermo.cuh)
__global__ void __launch_bounds__(256)
k_tc_prepare(const double *__restrict__ T, const double *__restrict__ P, const double *__restrict__ qv,
             double *__restrict__ g, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        g[i] = Tc_g(__ldg(T + i), __ldg(P + i), __ldg(qv + i));
}

// One fixed-point sweep: Tc <- B/(g - a*log(Tc_prev)) where Tc_prev = T for the first sweep.
// Two points per thread and iteration (16-byte loads/stores, two independent log/divide chains).
__global__ void __launch_bounds__(256)
k_tc_iter(const double *__restrict__ T, const double *__restrict__ g, double *__restrict__ Tc, int64_t n,
          int iter, unsigned long long *__restrict__ maxbits, int vec_ok)
{
    if (iter > 0 && __longlong_as_double((long long)maxbits[iter - 1]) < 0.1) return;  // converged
    unsigned long long m = 0ull;
    const int64_t npair = vec_ok ? (n >> 1) : 0;  // vec_ok: all pointers 16-byte aligned
    const double2 *__restrict__ T2 = reinterpret_cast<const double2 *>(T);
    const double2 *__restrict__ g2 = reinterpret_cast<const double2 *>(g);
    double2 *__restrict__ Tc2 = reinterpret_cast<double2 *>(Tc);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npair;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double2 prev = iter == 0 ? __ldg(T2 + i) : Tc2[i];
        const double2 gg = __ldg(g2 + i);
        double2 cur;
        cur.x = Tc_step_g(gg.x, prev.x);
        cur.y = Tc_step_g(gg.y, prev.y);
        Tc2[i] = cur;
        const unsigned long long b0 = absbits(cur.x - prev.x), b1 = absbits(cur.y - prev.y);
        m = b0 > m ? b0 : m;
        m = b1 > m ? b1 : m;
    }
    for (int64_t i = 2 * npair + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {  // odd tail / unaligned fallback
        const double prev = iter == 0 ? T[i] : Tc[i];
        const double cur = Tc_step_g(g[i], prev);
        Tc[i] = cur;
        const unsigned long long b = absbits(cur - prev);
        m = b > m ? b : m;
    }
    block_max_to(m, maxbits + iter);
}

// constants of td_point / tc_finish_point with a non-zero low word, in constant memory (see
// fastmath.cuh: an immediate costs two UMOVs per use)
struct TdConsts {
    double log_1e5, log_A622, log_611_2, kappa, t0, c1767, c6112, eps, LvCp, LvRd, Lv622, g, CpRd, logB, tol_h,
        c061;
};
static __constant__ TdConsts ktd = {11.512925464970229,   // log(100000)
                                    25.781840139430972,   // log(2.53e11*0.622)
                                    6.415424237852311,    // log(611.2)
                                    kappa, 273.15, 17.67, 611.2, 0.622, Lv / Cp, Lv / Rd, Lv * 0.622, g, CpRd,
                                    8.597851094433691,    // log(5420)
                                    1e-4, 0.61};

// The remaining `nrest` CERTAIN sweeps of calc_Tc for one point (t = the iterate after sweep
// `10 - nrest`, gg = the sweep-invariant g), used when the reference is known to run all ten sweeps
// (some point is NaN, see k_tc_finish).
// The sweeps are the fixed-point iteration L <- F(L) = log B - log(g - a L) on L = log Tc
// (a = Cp/Rd, contraction F' = a/D = a Tc/B ~ 0.18).  With L* the fixed point, D* = g - a L*
// and c = a/D*, the iterates obey EXACTLY  delta' = -log1p(-c delta)  for delta = L - L*.
// After the remaining sweeps delta is ~1e-9, so float32 carries it to ~1e-15 of Tc; what needs
// double precision is L* alone: two float32 Newton steps on h(L) = L - F(L) from log Tc_first
// and ONE double step (quadratic: error ~0.02 e^2) -- one double log and two double divisions
// per point instead of nine logs.  Measured against the literal ten sweeps: <= 3e-13 relative
// over T 180..330 K, P 10..1080 hPa, qv 1e-8..0.05 (tests/test_gpu_calc.py); points outside the
// series' range (|c delta|, |Tc_first/Tc* - 1| >= 0.1, no convergence, non-finite) take the
// literal sweeps, NaN points (terrain) stay NaN.
template <bool FLAG>
__device__ __forceinline__ double tc_finish_point_t(double t, double gg, int nrest, int exact,
                                                    const mt::fm::Tables *tb, bool &bad)
{
    constexpr double logB = 8.597851094433691;  // log(5420)
    if (t != t || gg != gg) return mt::nan_f64();
    const float gf = (float)gg, af = (float)CpRd, lbf = (float)logB;
    // MUFU-based float32 log / divide (abs. error ~1e-6 on L): the double step squares it
    float Lf = __logf((float)t);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float Df = gf - af * Lf;
        const float hf = Lf - lbf + __logf(Df);
        Lf = Lf - __fdividef(hf * Df, Df - af);
    }
    double L = (double)Lf;
    double D = gg - ktd.CpRd * L;
    const double h = L - ktd.logB + mt::fm::flog_t<FLAG>(D, tb, bad);
    L = L - mt::fm::fdiv_t<FLAG>(h * D, D - ktd.CpRd, bad);
    D = gg - ktd.CpRd * L;
    const double Tcs = mt::fm::fdiv_t<FLAG>(B, D, bad);
    const float c = __fdividef(af, (float)D);
    const float d = __fdividef((float)(t - Tcs), (float)Tcs);
    // delta_first = log1p(d)
    float dl = d * fmaf(d, fmaf(d, fmaf(d, fmaf(d, 0.2f, -0.25f), 0.33333334f), -0.5f), 1.0f);
    const bool ok = fabs(h) < ktd.tol_h && fabsf(d) < 0.1f && fabsf(c * dl) < 0.1f && D > 0.0;
    if (FLAG) {   // the caller redoes the point with the plain form, which runs the literal sweeps
        bad = bad || !ok || exact;
    } else if (!ok || exact) {   // the literal sweeps
        double Lx = log(t), Dx = gg - CpRd * Lx;
        for (int it = 1; it < nrest; ++it) {
            Lx = logB - log(Dx);
            Dx = gg - CpRd * Lx;
        }
        return B / Dx;
    }
    for (int k = 0; k < nrest; ++k) {
        const float u = c * dl;  // -log1p(-u) = u + u^2/2 + u^3/3 + ...
        dl = u * fmaf(u, fmaf(u, fmaf(u, fmaf(u, 0.2f, 0.25f), 0.33333334f), 0.5f), 1.0f);
    }
    const float e = fmaf(dl * dl, fmaf(dl, 0.16666667f, 0.5f), dl);  // expm1(delta)
    return Tcs + Tcs * (double)e;
}

__device__ __forceinline__ double tc_finish_point(double t, double gg, int nrest, int exact,
                                                  const mt::fm::Tables *tb)
{
    bool b = false;
    return tc_finish_point_t<false>(t, gg, nrest, exact, tb, b);
}

// Everything after the first sweep in ONE cooperative launch (replaces nine per-sweep launches,
// the iteration counter and the theta_e kernel):
//  * NaN fast path.  If the max |diff| of sweep `first-1` is NaN, some point is NaN for good (a NaN
//    iterate stays NaN), so np.max stays NaN and the reference runs all 10 sweeps no matter what
//    (terrain-masked data).  The remaining sweeps are then certain and are done
//    in registers in one pass over HBM.  Only log(Tc) is iterated, L' = log B - log(g - a L)
//    (one log per sweep, no division); Tc = B / (g - a L) is formed once at the end.
//  * otherwise: one grid-stride sweep per iteration, block max -> atomicMax, grid.sync(), stop as
//    soon as the global max |Tc - Tc2| < 0.1 -- the reference's test, evaluated on the device.
//  * finally theta_e = Th * exp(Lv*qv/Cp/Tc) (:456) when requested, from the Tc just produced.
__global__ void __launch_bounds__(256)
k_tc_finish(const double *__restrict__ g, double *__restrict__ Tc, int64_t n, int first,
            unsigned long long *__restrict__ maxbits, int vec_ok, const double *__restrict__ Th,
            const double *__restrict__ qv, double *__restrict__ Th_e, int exact, int speculated,
            const double *__restrict__ T)
{
    namespace cg = cooperative_groups;
    __shared__ mt::fm::Tables tb;
    mt::fm::load_tables(&tb);
    __syncthreads();
    const double prev_max = __longlong_as_double((long long)maxbits[first - 1]);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if (prev_max != prev_max) {  // NaN: all remaining sweeps, in registers
        if (tid == 0) {
            for (int it = first; it < 10; ++it) maxbits[it] = maxbits[first - 1];
            reinterpret_cast<double *>(maxbits)[15] = 10.0;
        }
        if (speculated) return;   // k_td_main already wrote the results of this branch
        const int nrest = 10 - first;  // sweeps still to run (>= 1); t below is the iterate Tc_first
        auto finish = [&](double t, double gg) { return tc_finish_point(t, gg, nrest, exact, &tb); };
        const int64_t npair = vec_ok ? (n >> 1) : 0;
        const double2 *__restrict__ g2 = reinterpret_cast<const double2 *>(g);
        double2 *__restrict__ Tc2 = reinterpret_cast<double2 *>(Tc);
        for (int64_t i = tid; i < npair; i += nthr) {
            double2 t = Tc2[i];
            const double2 gg = __ldg(g2 + i);
            t.x = finish(t.x, gg.x);
            t.y = finish(t.y, gg.y);
            Tc2[i] = t;
            if (Th_e) {
                const double2 th = __ldg(reinterpret_cast<const double2 *>(Th) + i);
                const double2 q = __ldg(reinterpret_cast<const double2 *>(qv) + i);
                __stcs(reinterpret_cast<double2 *>(Th_e) + i,
                       make_double2(th.x * mt::fm::fexp(mt::fm::fdiv((Lv / Cp) * q.x, t.x), &tb),
                                    th.y * mt::fm::fexp(mt::fm::fdiv((Lv / Cp) * q.y, t.y), &tb)));
            }
        }
        for (int64_t i = 2 * npair + tid; i < n; i += nthr) {
            const double t = finish(Tc[i], g[i]);
            Tc[i] = t;
            if (Th_e) __stcs(Th_e + i, __ldg(Th + i) * mt::fm::fexp(mt::fm::fdiv((Lv / Cp) * __ldg(qv + i), t), &tb));
        }
        return;
    }
    cg::grid_group grid = cg::this_grid();
    if (speculated)   // Tc holds k_td_main's speculative result: put the first iterate back (same
                      // expression and operands as td_point -> same bits); every thread rewrites
                      // exactly the elements it sweeps below, so no grid barrier is needed
        for (int64_t i = tid; i < n; i += nthr)
            Tc[i] = mt::fm::fdiv(B, g[i] - CpRd * mt::fm::flog(__ldg(T + i), &tb));
    int count = first;  // sweeps executed so far
    if (!(prev_max < 0.1)) {
        for (int it = first; it < 10; ++it) {
            unsigned long long m = 0ull;
            for (int64_t i = tid; i < n; i += nthr) {
                const double prev = Tc[i];
                const double cur = Tc_step_g(__ldg(g + i), prev);
                Tc[i] = cur;
                const unsigned long long b = absbits(cur - prev);
                m = b > m ? b : m;
            }
            block_max_to(m, maxbits + it);
            grid.sync();
            count = it + 1;
            if (__longlong_as_double((long long)maxbits[it]) < 0.1) break;
        }
    }
    if (tid == 0) reinterpret_cast<double *>(maxbits)[15] = (double)count;
    if (Th_e)
        for (int64_t i = tid; i < n; i += nthr)
            __stcs(Th_e + i, exp_factor(__ldg(Th + i), __ldg(qv + i), Tc[i]));
}

// ---- fused cyl_td main pass: one read of T,p,qv(,qc,qr), every requested output, and the first
//      Tc sweep.  The kernel is FP64-pipe bound (ncu: fp64 pipe 47 %, dram 30 %), so the
//      transcendental chains share their logarithms and avoid divisions where the result is a
//      tolerance-class output (exp/log/pow already differ from NumPy by ulps):
//        lP = log P, lT = log T, lq = log qv, lv = log vapor
//        Th   = T*exp(kappa*(log 1e5 - lP))                      (T*(1e5/P)**kappa)
//        g    = log(A*0.622) - lP - lq + (Cp/Rd)*lT              (log(A*0.622/P/qv) + a*log T)
//        Td   : ln(es/611.2) = lv - log 611.2
//      The pure-arithmetic outputs (vapor, Tv, rho) keep the reference's exact operation order.
struct TdPoint {
    double Th, es, qvs, vap, RH, Td, Th_es, Tv, rho, dTdz, g, Tc1;
};

template <bool FLAG>
__device__ __forceinline__ TdPoint td_point(double T, double P, double qv, const mt::fm::Tables *tb, bool &bad)
{
    auto flog = [&](double x, const mt::fm::Tables *t) { return mt::fm::flog_t<FLAG>(x, t, bad); };
    auto fexp = [&](double x, const mt::fm::Tables *t) { return mt::fm::fexp_t<FLAG>(x, t, bad); };
    auto fdiv = [&](double a, double b) { return mt::fm::fdiv_t<FLAG>(a, b, bad); };
    TdPoint r;
    const double lP = flog(P, tb), lT = flog(T, tb), lq = flog(qv, tb);
    r.Th = T * fexp(ktd.kappa * (ktd.log_1e5 - lP), tb);
    const double Tc_ = T - ktd.t0;
    r.es = ktd.c6112 * fexp(fdiv(ktd.c1767 * Tc_, Tc_ + 243.5), tb);   // :116
    r.qvs = fdiv(ktd.eps * r.es, P - r.es);                    // :151
    r.vap = qv_to_vapor(P, qv);                           // exact class, :168
    r.RH = fdiv(r.vap, r.es);
    const double lnes = flog(r.vap, tb) - ktd.log_611_2;
    r.Td = fdiv(243.5 * lnes, ktd.c1767 - lnes) + ktd.t0;
    const double rT = fdiv(1.0, T);                       // shared by the three tolerance-class quotients
    r.Th_es = r.Th * fexp(ktd.LvCp * r.qvs * rT, tb);    // Lv*qvs/Cp/T
    r.Tv = Tv_from_qv(T, qv);                             // exact class, :335
    r.rho = rho(P, r.Tv);                                 // exact class, :364
    const double a = ktd.LvRd * r.qvs * rT;               // Lv*qvs/Rd/T
    r.dTdz = -ktd.g * fdiv(1 + a, Cp + a * ktd.Lv622 * rT);  // :486
    r.g = ktd.log_A622 - lP - lq + ktd.CpRd * lT;
    r.Tc1 = fdiv(B, r.g - ktd.CpRd * lT);                 // first sweep: Tc2 = T
    return r;
}

// ALL = true: the whole cyl_td set is requested and qc, qr are present (calc_all_thermaldynamic):
// the output tests are compile-time true (the kernel is instruction-issue bound).
#ifndef MT_TD_MINB
#define MT_TD_MINB 4
#endif
template <bool ALL>
__global__ void __launch_bounds__(256, MT_TD_MINB)
k_td_main(mt_td_in_t in, mt_td_out_t out, int64_t n, unsigned long long *__restrict__ maxbits,
          double *__restrict__ gbuf, int exact)
{
    // one point per thread and iteration: 60 registers / 4 CTAs per SM measured 25 % faster than
    // two points per thread (117 registers) -- profiles/experiments_r01.md
    __shared__ mt::fm::Tables tb;
    mt::fm::load_tables(&tb);
    __syncthreads();
    unsigned long long m = 0ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double T = __ldg(in.T + i), P = __ldg(in.p + i), qv = __ldg(in.qv + i);
        // liquid water of theta_v (:66-78: qc+qr, qc, qr or the default 0): loaded with the other
        // inputs so that the latency hides behind td_point (ncu: 15 % of the stall samples sat on
        // these two loads when they were issued at their use)
        double ql = 0.0;
        if (ALL || out.Th_v) {
            if (ALL || (in.qc && in.qr)) ql = __ldg(in.qc + i) + __ldg(in.qr + i);
            else if (in.qc) ql = __ldg(in.qc + i);
            else if (in.qr) ql = __ldg(in.qr + i);
        }
        // fast path without a single branch: every rare-argument test of the ~24 log / exp /
        // divide calls is OR-ed into `bad`; a point that needs a slow path (never, for physical
        // input) is redone with the plain forms, which gives what the branchy calls gave
        bool bad = false;
        TdPoint r = td_point<true>(T, P, qv, &tb, bad);
        double tc = 0.0, the = 0.0;
        if (ALL || out.Tc) {
            tc = tc_finish_point_t<true>(r.Tc1, r.g, 9, exact, &tb, bad);
            if (ALL || out.Th_e) the = r.Th * mt::fm::fexp_t<true>(mt::fm::fdiv_t<true>(ktd.LvCp * qv, tc, bad), &tb, bad);
        }
        if (bad) {   // the same point through the plain (self-guarding) forms
            bool unused = false;
            r = td_point<false>(T, P, qv, &tb, unused);
            if (ALL || out.Tc) {
                tc = tc_finish_point(r.Tc1, r.g, 9, exact, &tb);
                if (ALL || out.Th_e) the = r.Th * mt::fm::fexp(mt::fm::fdiv(ktd.LvCp * qv, tc), &tb);
            }
        }
        if (ALL || out.Th) __stcs(out.Th + i, r.Th);
        if (ALL || out.vapor_s) __stcs(out.vapor_s + i, r.es);
        if (ALL || out.qvs) __stcs(out.qvs + i, r.qvs);
        if (ALL || out.vapor) __stcs(out.vapor + i, r.vap);
        if (ALL || out.RH) __stcs(out.RH + i, r.RH);
        if (ALL || out.Td) __stcs(out.Td + i, r.Td);
        if (ALL || out.Th_es) __stcs(out.Th_es + i, r.Th_es);
        if (ALL || out.Th_v) __stcs(out.Th_v + i, r.Th * (1 + qv * ktd.c061 - ql));   // theta_v, :309
        if (ALL || out.Tv) __stcs(out.Tv + i, r.Tv);
        if (ALL || out.rho) __stcs(out.rho + i, r.rho);
        if (ALL || out.moist_dTdz) __stcs(out.moist_dTdz + i, r.dTdz);
        if (ALL || out.Tc) {
            // SPECULATION: data on a terrain-masked grid hold NaN, and then the reference is
            // certain to run all ten calc_Tc sweeps (k_tc_finish).  The result of that branch --
            // Tc after the nine remaining sweeps and theta_e -- is produced here, while T, p, qv
            // and Th are in registers; k_tc_finish then has nothing left to do.  Without NaN it
            // restores the first iterate from g and T and runs the reference's sweeps.
            gbuf[i] = r.g;
            const unsigned long long b = absbits(r.Tc1 - T);
            m = b > m ? b : m;
            out.Tc[i] = tc;
            if (ALL || out.Th_e) __stcs(out.Th_e + i, the);
        }
    }
    if (ALL || out.Tc) block_max_to(m, maxbits);
}

// sweeps first..9 (+ theta_e): one cooperative launch sized to what is co-resident
bool tc_exact_mode()   // MT_TC_EXACT=1: the literal sweeps for every point (A/B measurements)
{
    static const bool e = getenv("MT_TC_EXACT") && getenv("MT_TC_EXACT")[0] == '1';
    return e;
}

int tc_tail(const double *T, const double *g, double *Tc, int64_t n, unsigned long long *bits, int first_iter,
            const double *Th, const double *qv, double *Th_e, int speculated, cudaStream_t s)
{
    const int vec_ok = (((uintptr_t)T | (uintptr_t)g | (uintptr_t)Tc | (uintptr_t)Th | (uintptr_t)qv |
                         (uintptr_t)Th_e) % 16 == 0) ? 1 : 0;
    if (first_iter == 0) {  // sweep 0 decides whether the NaN fast path applies
        mt::ProfScope prof("k_tc_iter", (mt_stream_t)s);
        k_tc_iter<<<mt::grid_for((n + 1) / 2, 256, 8), 256, 0, s>>>(T, g, Tc, n, 0, bits, vec_ok);
        MT_LAUNCH_CHECK();
        first_iter = 1;
    }
    static thread_local int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        MT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_tc_finish, 256, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    int64_t want = ((n + 1) / 2 + 255) / 256;
    int64_t cap = (int64_t)mt::sm_count() * blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
    int exact = tc_exact_mode() ? 1 : 0;
    void *args[] = {(void *)&g, (void *)&Tc, (void *)&n, (void *)&first_iter, (void *)&bits, (void *)&vec_ok,
                    (void *)&Th, (void *)&qv, (void *)&Th_e, (void *)&exact, (void *)&speculated, (void *)&T};
    mt::ProfScope prof("k_tc_finish", (mt_stream_t)s);
    MT_CUDA(cudaLaunchCooperativeKernel((const void *)k_tc_finish, dim3(grid), dim3(256), args, 0, s));
    MT_LAUNCH_CHECK();
    return MT_OK;
}

}  // namespace

extern "C" {

int mt_pointwise(int op, const double *a, const double *b, const double *c, double sb, double sc,
                 double *out, int64_t n, mt_stream_t stream)
{
    MT_REQUIRE(a && out && n >= 0, "mt_pointwise: null pointer");
    if (n == 0) return MT_OK;
    const unsigned grid = mt::grid_for(n, 256, 8);
    cudaStream_t s = (cudaStream_t)stream;
    mt::ProfScope prof("k_pointwise", stream);
    switch (op) {
#define PW(K) \
    case K: k_pointwise<K><<<grid, 256, 0, s>>>(a, b, c, sb, sc, out, n); break;
        PW(MT_OP_THETA) PW(MT_OP_T) PW(MT_OP_SAT_VAPOR) PW(MT_OP_VAPOR) PW(MT_OP_SAT_QV)
        PW(MT_OP_QV_TO_VAPOR) PW(MT_OP_THETA_ES) PW(MT_OP_TD) PW(MT_OP_QV) PW(MT_OP_THETA_V)
        PW(MT_OP_TV_QV) PW(MT_OP_TV_VAPOR) PW(MT_OP_RHO) PW(MT_OP_THETA_E) PW(MT_OP_MOIST_DTDZ)
        PW(MT_OP_RH) PW(MT_OP_P_TO_Z) PW(MT_OP_Z_TO_P) PW(MT_OP_ADD)
#undef PW
        default: mt::set_error("mt_pointwise: unknown op %d", op); return MT_EINVAL;
    }
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_calc_tc(const double *T, const double *P, const double *qv, double *Tc, int64_t n,
               double *scratch, mt_stream_t stream)
{
    MT_REQUIRE(T && P && qv && Tc && scratch && n > 0, "mt_calc_tc: null pointer or empty array");
    cudaStream_t s = (cudaStream_t)stream;
    MT_CUDA(cudaMemsetAsync(scratch, 0, 16 * sizeof(double), s));
    double *g = scratch + 16;
    {
        mt::ProfScope prof("k_tc_prepare", stream);
        k_tc_prepare<<<mt::grid_for(n, 256, 8), 256, 0, s>>>(T, P, qv, g, n);
        MT_LAUNCH_CHECK();
    }
    return tc_tail(T, g, Tc, n, reinterpret_cast<unsigned long long *>(scratch), 0, nullptr, nullptr, nullptr, 0, s);
}

int mt_cyl_td(const mt_td_in_t *in, const mt_td_out_t *out, int64_t n, double *scratch,
              mt_stream_t stream)
{
    MT_REQUIRE(in && out && in->T && in->p && in->qv && n > 0, "mt_cyl_td: T, p, qv are required");
    MT_REQUIRE(!out->Th_e || (out->Tc && out->Th), "mt_cyl_td: Th_e needs the Tc and Th output buffers");
    MT_REQUIRE(!out->Tc || scratch, "mt_cyl_td: Tc needs the (16 + n)-double scratch");
    cudaStream_t s = (cudaStream_t)stream;
    auto bits = reinterpret_cast<unsigned long long *>(scratch);
    if (out->Tc) MT_CUDA(cudaMemsetAsync(scratch, 0, 16 * sizeof(double), s));
    {
        mt::ProfScope prof("k_td_main", stream);
        const bool all = in->qc && in->qr && out->Th && out->vapor_s && out->qvs && out->vapor && out->RH &&
                         out->Td && out->Th_es && out->Th_v && out->Tv && out->rho && out->moist_dTdz && out->Tc &&
                         out->Th_e;
        // four waves of what is resident (grid-stride loop inside; 1 / 2 / 4 / 16 waves: 0.225 / 0.222 / 0.219 / 0.228 ms)
        static thread_local int occ = 0;
        if (occ == 0) {
            MT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_td_main<true>, 256, 0));
            if (occ < 1) occ = 1;
        }
        const unsigned grid = mt::grid_for(n, 256, 4 * occ);
        if (all)
            k_td_main<true><<<grid, 256, 0, s>>>(*in, *out, n, bits, scratch + 16, tc_exact_mode() ? 1 : 0);
        else
            k_td_main<false><<<grid, 256, 0, s>>>(*in, *out, n, bits, scratch ? scratch + 16 : nullptr,
                                                  tc_exact_mode() ? 1 : 0);
        MT_LAUNCH_CHECK();
    }
    if (out->Tc) {
        // Th_e needs Th: taken from the output buffer when requested, else recomputed is not
        // possible here -> Th is required as an output whenever Th_e is
        int rc = tc_tail(in->T, scratch + 16, out->Tc, n, bits, 1, out->Th, in->qv, out->Th_e, 1, s);
        if (rc != MT_OK) return rc;
    }
    return MT_OK;
}

}  // extern "C"
