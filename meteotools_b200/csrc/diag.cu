// Fused dynamics diagnostics: the cylindrical set of
// grids/cylindrical_grid.py:264-508 ("F2") and the Cartesian set of
// grids/cartesian_grid.py:232-364 ("F3").
//
// The reference evaluates every diagnostic as a chain of full-array NumPy passes (each FD2 alone
// makes ~6 temporaries).  Here a settings file's whole set is two passes over HBM:
//   stage A (point-wise): vr, vt, wind speed, kinetic energy, Tv, rho, density factor, Th,
//                         Th_rho, Coriolis and centrifugal terms;
//   stage B (stencil):    every FD2-based quantity (vorticity, divergence, PV, deformation,
//                         filamentation, pressure-gradient / agradient force) from the stage-A
//                         fields, one thread per point, neighbours through L1/L2.
// Stages are separately launchable so that a level-sharded multi-GPU run can exchange the one
// halo level of the z-differentiated stage-A fields between them.
// Every expression keeps the reference's order of evaluation (bit-exact for pure arithmetic).
#include <type_traits>

#include "diag.cuh"

namespace {

using mt::nan_f64;
using mt::inf_f64;
using namespace mt::dg;

// ------------------------------------------------------------------------------- stage A
// WIND: how the horizontal wind reaches the point-wise outputs (diag.cuh): WIND_NONE -- no wind output is
// requested from this launch (the TMA stencil kernel produces them): u, v are not read; WIND_UV -- (u, v),
// rotated on a cylindrical grid; WIND_GIVEN -- the caller's (vr, vt), used as they are.
template <int LAYOUT, bool CYL, int WIND>
__global__ void __launch_bounds__(256)
k_stage_a(Lay<LAYOUT> L, GridP g, mt_dyn_in_t in, mt::th::Hydro hyd, StageAOut o)
{
    const int64_t total = (int64_t)L.nz * L.nt * L.nr;
    // (iz, it, ir) of the flat index are decoded ONCE per thread (two 64-bit div / mod pairs) and then advanced
    // by the grid stride with adds and carries: per point the decode used to cost more issue slots than the
    // arithmetic (the wind-carrying instantiation ran at 4.2 TB/s, the index-free thermodynamic one at 6.4).
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= total) return;
    // a0: contiguous axis, a1: middle axis, a2: slowest axis of the layout
    const int n_a0 = LAYOUT == MT_LAYOUT_ZTR ? L.nr : L.nz, n_a1 = LAYOUT == MT_LAYOUT_ZTR ? L.nt : L.nr;
    int i0 = (int)(m % n_a0), i1 = (int)((m / n_a0) % n_a1), i2 = (int)(m / ((int64_t)n_a0 * n_a1));
    const int s0 = (int)(stride % n_a0), s1 = (int)((stride / n_a0) % n_a1), s2 = (int)(stride / ((int64_t)n_a0 * n_a1));
    for (; m < total; m += stride) {
        const int iz = LAYOUT == MT_LAYOUT_ZTR ? i2 : i0, it = LAYOUT == MT_LAYOUT_ZTR ? i1 : i2,
                  ir = LAYOUT == MT_LAYOUT_ZTR ? i0 : i1;
        // advance to the next point of this thread (used at the end of the body; computed here so that
        // the body's `continue`-free structure stays as it was)
        int j0 = i0 + s0, j1 = i1 + s1, j2 = i2 + s2;
        if (j0 >= n_a0) j0 -= n_a0, ++j1;
        if (j1 >= n_a1) j1 -= n_a1, ++j2;
        i0 = j0, i1 = j1, i2 = j2;
        (void)iz;
        double vr = 0, vt = 0;
        if (WIND != WIND_NONE) {
            double u = 0, v = 0;
            if (WIND == WIND_UV || o.wind_speed) u = __ldg(in.u + m), v = __ldg(in.v + m);
            if (CYL && WIND == WIND_UV) {
                const double s = __ldg(g.sin_t + it), c = __ldg(g.cos_t + it);
                vr = v * s + u * c;  // cylindrical_grid.py:267
                vt = v * c - u * s;  // :268
                if (o.vr) o.vr[m] = vr;
                if (o.vt) o.vt[m] = vt;
            } else if (CYL) {
                vr = __ldg(in.vr + m), vt = __ldg(in.vt + m);
            }
            if (o.wind_speed) __stcs(o.wind_speed + m, sqrt(u * u + v * v));  // (u**2+v**2)**0.5
            if (o.kinetic_energy && in.w) {
                const double w = __ldg(in.w + m);
                const double ke = CYL ? (vr * vr + vt * vt + w * w) / 2 : (u * u + v * v + w * w) / 2;
                __stcs(o.kinetic_energy + m, ke);
            }
        }
        const bool need_thermo = o.Tv || o.rho || o.density_factor || o.Th || o.Th_rho;
        if (need_thermo) {
            const double qv = __ldg(in.qv + m);
            double T = 0, P = 0;
            if (in.T) T = __ldg(in.T + m);
            if (in.p) P = __ldg(in.p + m);
            if (o.Tv || o.rho) {
                const double Tv = mt::th::Tv_from_qv(T, qv);
                if (o.Tv) __stcs(o.Tv + m, Tv);
                if (o.rho) o.rho[m] = mt::th::rho(P, Tv);
            }
            if (o.density_factor || o.Th || o.Th_rho) {
                double df = 0;
                if (o.density_factor || o.Th_rho) {
                    if (hyd.n < 0) df = (1 + qv / 0.622) / 1.0;  // no hydrometeors: b = 1 (:443-445)
                    else df = mt::th::density_factor(qv, hyd, m);
                    if (o.density_factor) __stcs(o.density_factor + m, df);
                }
                const double Th = in.Th ? __ldg(in.Th + m) : mt::th::theta(T, P);
                if (o.Th) __stcs(o.Th + m, Th);
                if (o.Th_rho) o.Th_rho[m] = Th * df;  // :453
            }
        }
        if (CYL && WIND != WIND_NONE && (o.coriolis_force || o.centrifugal_force || o.aam)) {
            const double rr = __ldg(g.r + ir);
            if (o.coriolis_force) __stcs(o.coriolis_force + m, __ldg(in.f + (int64_t)it * L.nr + ir) * vt);
            if (o.centrifugal_force) __stcs(o.centrifugal_force + m, vt * vt / rr);
            if (o.aam)  // :297  vt*r + 0.5*f*r**2
                __stcs(o.aam + m, vt * rr + 0.5 * __ldg(in.f + (int64_t)it * L.nr + ir) * (rr * rr));
        }
    }
}

// ------------------------------------------------------------------------------- stage B
#ifndef MT_STAGEB_MINB
#define MT_STAGEB_MINB 8   // resident CTAs per SM the interior kernel is compiled for: 64 registers (a few
                           // spilled words) beat 80 registers at 6 CTAs -- the kernel is load-latency bound
#endif
// Stage B marches along the SLOWEST memory axis (z for "ZTR", theta for "TRZ"): one thread per
// position of the (contiguous) plane keeps the previous / current / next plane values of the four
// differentiated fields in registers, so every plane is read from HBM once even when a plane is
// far larger than the L2 (S5: 32 MB per field and level); the two in-plane directions come from
// neighbouring threads' lines through L1.  The march axis is cut into segments (grid.y) only to
// create enough threads for small planes (S3: 18 000 positions x 360 theta).
struct Roll {
    double p, c, n;
};

// FD2 along the march axis from the rolling registers; `far(k)` loads element k of the axis and
// is needed only by the one-sided formulas at the global ends (differential.py:32,34).
template <class Far>
__device__ __forceinline__ double dmarch(const Roll &r, int i, int n, bool first, bool last, double delta,
                                         Far far)
{
    if (i == 0 && first) return (-3 * r.c + 4 * r.n - far(2)) / 2 / delta;
    if (i == n - 1 && last) return (3 * r.c - 4 * r.p + far(n - 3)) / 2 / delta;
    return (r.n - r.p) / 2 / delta;
}

// SHELL = false: the marching kernel proper, INTERIOR points only (centred differences, no
//                 branches, modest register count);
// SHELL = true:  one thread per point of the outermost shell (global first / last index of any
//                 axis), general body with the one-sided formulas; planes are loaded directly.
// FULL = true: every stencil output of the set is requested (calc_diagnostics): the output tests are
//                compile-time true, the body is straight-line code and the compiler can issue all
//                neighbour loads of a point together (the generic body issues them in small groups
//                between run-time branches and pays the L2 latency once per group).
template <int LAYOUT, bool CYL, bool SHELL, bool FULL>
// the shell launch (two end planes: one thread per point, ~35 dependent-free loads) is latency bound: twelve CTAs of
// 40 registers per SM beat six of 80 although a few values spill (S5: 0.56 -> 0.38 ms; 4 / 6 / 8 / 10 / 12 / 16 CTAs on
// 80 x 1000 x 1000: 0.158 / 0.140 / 0.129 / 0.111 / 0.101 / 0.107 ms)
#ifndef MT_SHELL_MINB
#define MT_SHELL_MINB 12
#endif
__global__ void __launch_bounds__(128, SHELL ? MT_SHELL_MINB : MT_STAGEB_MINB)
k_stage_b(Lay<LAYOUT> L, GridP g, StageBIn in, StageBOut o, int seg_len)
{
    constexpr bool ZTR = LAYOUT == MT_LAYOUT_ZTR;
    const int nz = L.nz, nt = L.nt, nr = L.nr;
    const int n0 = ZTR ? nz : nt;   // march axis (slowest in memory)
    const int n2 = ZTR ? nr : nz;   // contiguous axis
    const int64_t plane = (int64_t)(ZTR ? nt : nr) * n2;
    int64_t p;
    int i1, i2, s_lo, s_hi;
    bool first = true, last = true;
    if (ZTR) first = g.bottom, last = g.top;
    if (!SHELL) {
        p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        if (p >= plane) return;
        i1 = (int)(p / n2), i2 = (int)(p - (int64_t)i1 * n2);
        int m_lo = 0, m_hi = n0;
        if (ZTR) {  // a level shard marches over its owned levels and reads the halo levels
            m_lo = g.own_lo, m_hi = g.own_hi;
        } else if (i2 < g.own_lo || i2 >= g.own_hi) {
            return;  // TRZ: z is the contiguous axis; halo levels produce no output
        }
        s_lo = m_lo + (int)blockIdx.y * seg_len;
        s_hi = min(m_hi, s_lo + seg_len);
        if (s_lo >= s_hi) return;
    } else {
        int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        int iz, it, ir;
        if (seg_len == 2) {
            // only the two END PLANES of the march axis (the TMA kernel of csrc/diag_tma.cu produces every
            // other point, the edges of the planes included): whole contiguous planes, coalesced
            const int nf = (first ? 1 : 0) + (last ? 1 : 0);
            if (q >= (int64_t)nf * plane) return;
            const int f = (int)(q / plane);
            const int64_t rem = q - (int64_t)f * plane;
            const int m_end = (f == 0 && first) ? (ZTR ? g.own_lo : 0) : (ZTR ? g.own_hi - 1 : n0 - 1);
            const int j1 = (int)(rem / n2), j2 = (int)(rem - (int64_t)j1 * n2);
            iz = ZTR ? m_end : j2, it = ZTR ? j1 : m_end, ir = ZTR ? j2 : j1;
        } else {
        // compact enumeration of the outermost shell of the owned block: the global z faces, the two
        // theta faces of the remaining levels, the two r faces of what is left -- one thread per
        // shell point (the first version launched a thread per GRID point and let 95 % of them
        // leave at once: 0.26 ms of the 0.63 ms of stage B on S3)
        const int zin_lo = g.own_lo + (g.bottom ? 1 : 0), zin_hi = g.own_hi - (g.top ? 1 : 0);
        const int nzin = max(zin_hi - zin_lo, 0);
        const int64_t nA = (int64_t)((g.bottom ? 1 : 0) + (g.top ? 1 : 0)) * nt * nr;
        const int64_t nB = 2LL * nzin * nr, nC = 2LL * nzin * (nt - 2);
        if (q < nA) {
            const int f = (int)(q / ((int64_t)nt * nr));
            const int64_t rem = q - (int64_t)f * nt * nr;
            it = (int)(rem / nr), ir = (int)(rem - (int64_t)it * nr);
            iz = (f == 0 && g.bottom) ? g.own_lo : g.own_hi - 1;
        } else if ((q -= nA) < nB) {
            const int f = (int)(q / ((int64_t)nzin * nr));
            const int64_t rem = q - (int64_t)f * nzin * nr;
            if (ZTR) {
                iz = zin_lo + (int)(rem / nr), ir = (int)(rem % nr);
            } else {
                ir = (int)(rem / nzin), iz = zin_lo + (int)(rem % nzin);
            }
            it = f ? nt - 1 : 0;
        } else if ((q -= nB) < nC) {
            const int f = (int)(q / ((int64_t)nzin * (nt - 2)));
            const int64_t rem = q - (int64_t)f * nzin * (nt - 2);
            iz = zin_lo + (int)(rem / (nt - 2)), it = 1 + (int)(rem % (nt - 2));
            ir = f ? nr - 1 : 0;
        } else {
            return;
        }
        }
        s_lo = ZTR ? iz : it, s_hi = s_lo + 1;
        i1 = ZTR ? it : ir, i2 = ZTR ? ir : iz;
        p = (int64_t)i1 * n2 + i2;
    }

    const bool want_zr = FULL || o.zeta_r || o.PV;
    const bool want_zt = FULL || o.zeta_t || o.PV;
    const bool want_zz = FULL || o.zeta_z || o.abs_zeta_z || o.PV || o.filamentation_time || o.strain_region;
    const bool want_def = FULL || o.shear_deform || o.stretch_deform || o.filamentation_time || o.strain_region;
    const bool want_div = FULL || o.div_hori || o.div;
    const bool use_ab = FULL || want_zr || want_zt || want_zz || want_def || want_div || o.agradient_force;
    const bool use_w = FULL || (in.w && (want_zr || want_zt || o.div));
    const bool use_t = FULL || o.PV != nullptr;

    int64_t m = (int64_t)s_lo * plane + p;
    Roll ra = {0, 0, 0}, rb = {0, 0, 0}, rw = {0, 0, 0}, rt = {0, 0, 0};
    auto prime = [&](const double *F, Roll &r) {
        r.n = __ldg(F + m);
        r.c = s_lo > 0 ? __ldg(F + m - plane) : 0.0;
    };
    if (use_ab) prime(in.a, ra), prime(in.b, rb);
    if (use_w) prime(in.w, rw);
    if (use_t) prime(in.thr, rt);

    for (int i0 = s_lo; i0 < s_hi; ++i0, m += plane) {
        auto roll = [&](const double *F, Roll &r) {
            r.p = r.c, r.c = r.n;
            r.n = (i0 + 1 < n0) ? __ldg(F + m + plane) : 0.0;
        };
        if (use_ab) roll(in.a, ra), roll(in.b, rb);
        if (use_w) roll(in.w, rw);
        if (use_t) roll(in.thr, rt);
        const int iz = ZTR ? i0 : i2, it = ZTR ? i1 : i0, ir = ZTR ? i2 : i1;
        const int64_t sr = ZTR ? 1 : nz;
        const double rr = CYL ? __ldg(g.r + ir) : 1.0;
        // INTERIOR points (all but the outermost shell) only ever need centred differences: they
        // take a branch-free instantiation of the body; the shell takes the general one with the
        // one-sided formulas.  Same expressions, same order of evaluation in both.
        auto body = [&](auto interior_tag) {
            constexpr bool IN = decltype(interior_tag)::value;
            auto dz = [&](const double *F, const Roll &r) -> double {
                if (ZTR) {
                    if (IN) return (r.n - r.p) / 2 / g.dz;
                    return dmarch(r, i0, n0, first, last, g.dz, [&](int k) { return __ldg(F + p + (int64_t)k * plane); });
                }
                if (IN) return (__ldg(F + m + 1) - __ldg(F + m - 1)) / 2 / g.dz;
                return mt::fd2_shard([&](int k) { return __ldg(F + m + (k - iz)); }, iz, nz, g.dz, g.bottom, g.top);
            };
            auto dt = [&](const double *F, const Roll &r) -> double {
                if (ZTR) {
                    if (IN) return (__ldg(F + m + nr) - __ldg(F + m - nr)) / 2 / g.dt;
                    return mt::fd2([&](int k) { return __ldg(F + m + (int64_t)(k - it) * nr); }, it, nt, g.dt);
                }
                if (IN) return (r.n - r.p) / 2 / g.dt;
                return dmarch(r, i0, n0, true, true, g.dt, [&](int k) { return __ldg(F + p + (int64_t)k * plane); });
            };
            auto dr = [&](const double *F) -> double {
                if (IN) return (__ldg(F + m + sr) - __ldg(F + m - sr)) / 2 / g.dr;
                return mt::fd2([&](int k) { return __ldg(F + m + (int64_t)(k - ir) * sr); }, ir, nr, g.dr);
            };
            // FD2 along r of  F * r  (vt*r, :391) and of  r * F  (r*vr, :408)
            auto dr_times_r = [&](const double *F, bool r_first) -> double {
                auto v = [&](int k) {
                    const double fv = __ldg(F + m + (int64_t)(k - ir) * sr), rk = __ldg(g.r + k);
                    return r_first ? rk * fv : fv * rk;
                };
                if (IN) return (v(ir + 1) - v(ir - 1)) / 2 / g.dr;
                return mt::fd2(v, ir, nr, g.dr);
            };

            double zeta_r = 0, zeta_t = 0, zeta_z = 0, abs_zz = 0;
            double da_dt = 0, db_dt = 0;  // d(vr|u)/dtheta|dy , d(vt|v)/dtheta|dy
            if (want_zz || want_def || want_div) {
                da_dt = dt(in.a, ra);
                db_dt = dt(in.b, rb);
            }
            double db_dz = 0;
            if (want_zr || (!CYL && want_zt)) db_dz = dz(in.b, rb);
            if (want_zr) {
                const double dw_dt = dt(in.w, rw);
                // cyl :387-388  FD2(w,dtheta,1)/r - FD2(vt,dz,0) ; car :250  FD2(w,dy,1) - FD2(v,dz,0)
                zeta_r = CYL ? dw_dt / rr - db_dz : dw_dt - db_dz;
                if (FULL || o.zeta_r) __stcs(o.zeta_r + m, zeta_r);
            }
            if (want_zt) {
                const double dw_dr = dr(in.w);
                if (CYL) zeta_t = dz(in.a, ra) - dw_dr;  // :389-390
                else zeta_t = db_dz - dw_dr;             // cartesian_grid.py:251 (sic: v, not u)
                if (FULL || o.zeta_t) __stcs(o.zeta_t + m, zeta_t);
            }
            double da_dr = 0, db_dr = 0;
            if (!CYL && (want_zz || want_def || want_div)) {
                da_dr = dr(in.a);
                db_dr = dr(in.b);
            }
            if (want_zz) {
                if (CYL) zeta_z = dr_times_r(in.b, false) / rr - da_dt / rr;  // :391-392
                else zeta_z = db_dr - da_dt;                                   // :252 FD2(v,dx,2) - FD2(u,dy,1)
                if (FULL || o.zeta_z) __stcs(o.zeta_z + m, zeta_z);
                if (FULL || o.abs_zeta_z || o.PV) {
                    abs_zz = zeta_z + __ldg(in.f + (int64_t)it * nr + ir);  // :400
                    if (FULL || o.abs_zeta_z) __stcs(o.abs_zeta_z + m, abs_zz);
                }
            }
            if (want_div) {
                double dh;
                if (CYL) dh = dr_times_r(in.a, true) / rr + db_dt / rr;  // :408-409
                else dh = da_dr + db_dt;                                  // :266-267
                if (FULL || o.div_hori) __stcs(o.div_hori + m, dh);
                if (FULL || o.div) __stcs(o.div + m, dh + dz(in.w, rw));
            }
            if (FULL || o.PV) {
                const double dthr_dr = dr(in.thr);
                const double dthr_dt = dt(in.thr, rt);
                const double dthr_dz = dz(in.thr, rt);
                const double r_part = zeta_r * dthr_dr;                                  // :463 / :319
                const double t_part = CYL ? zeta_t * dthr_dt / rr : zeta_t * dthr_dt;    // :464-465 / :320
                const double z_part = abs_zz * dthr_dz;                                  // :466 / :321
                __stcs(o.PV + m, (r_part + t_part + z_part) / __ldg(in.rho + m));
            }
            if (want_def) {
                double sh, stc;
                if (CYL) {
                    sh = dr(in.a) - ra.c / rr - db_dt / rr;   // :490-491
                    stc = dr(in.b) - rb.c / rr + da_dt / rr;  // :492-493
                } else {
                    sh = da_dr - db_dt;   // :344-345 FD2(u,dx) - FD2(v,dy)
                    stc = db_dr + da_dt;  // :346-347 FD2(v,dx) + FD2(u,dy)
                }
                if (FULL || o.shear_deform) __stcs(o.shear_deform + m, sh);
                if (FULL || o.stretch_deform) __stcs(o.stretch_deform + m, stc);
                if (FULL || o.filamentation_time || o.strain_region) {
                    const double s2 = sh * sh + stc * stc, z2 = zeta_z * zeta_z;
                    const bool strain = s2 > z2;  // :504-505
                    if ((FULL && CYL) || o.strain_region) o.strain_region[m] = strain ? 1 : 0;
                    if (FULL || o.filamentation_time)  // :501-508, x**-0.5 as 1/sqrt(x): <= 1 ulp
                        __stcs(o.filamentation_time + m, strain ? 2 * (1.0 / sqrt(s2 - z2)) : inf_f64());
                }
            }
            if (CYL && (FULL || o.radial_PG_force || o.agradient_force)) {
                const double pg = -dr(in.p) / __ldg(in.rho + m);  // :343
                if (FULL || o.radial_PG_force) __stcs(o.radial_PG_force + m, pg);
                if (FULL || o.agradient_force) {
                    const double vt = rb.c;
                    const double cor = __ldg(in.f + (int64_t)it * nr + ir) * vt;  // :341
                    const double cen = vt * vt / rr;                             // :342
                    __stcs(o.agradient_force + m, cor + cen + pg);               // :344-346
                }
            }
        };
        if (SHELL) {
            body(std::false_type{});
        } else {
            const bool z_edge = (iz == 0 && g.bottom) || (iz == nz - 1 && g.top);
            if (!z_edge && it > 0 && it < nt - 1 && ir > 0 && ir < nr - 1) body(std::true_type{});
        }
    }
}

// nanmean over theta + anomaly (cylindrical_grid.py:225-229).  One thread per (z, r); the sum
// runs over theta in index order.  theta is a strided axis of the reference's (z, theta, r) arrays,
// where NumPy's reduction adds sequentially (pairwise blocks are used along the contiguous axis
// only): the result is bit-identical to np.nanmean(axis=1) (tests assert equality).
template <int LAYOUT>
__global__ void __launch_bounds__(128)
k_azimuthal(Lay<LAYOUT> L, const double *__restrict__ var, double *__restrict__ sym,
            double *__restrict__ asy)
{
    const int64_t total = (int64_t)L.nz * L.nr;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
         q += (int64_t)gridDim.x * blockDim.x) {
        int iz, ir;
        if (LAYOUT == MT_LAYOUT_ZTR) {
            ir = (int)(q % L.nr);
            iz = (int)(q / L.nr);
        } else {
            iz = (int)(q % L.nz);
            ir = (int)(q / L.nz);
        }
        const int64_t base = iz * L.sz + ir * L.sr;
        double s = 0.0;
        int cnt = 0;
        for (int t = 0; t < L.nt; ++t) {
            const double v = __ldg(var + base + t * L.st);
            if (!isnan(v)) {
                s += v;
                ++cnt;
            }
        }
        const double mean = cnt ? s / (double)cnt : nan_f64();
        sym[(int64_t)iz * L.nr + ir] = mean;  // sym is (nz, nr) C-order
        if (asy)
            for (int t = 0; t < L.nt; ++t) asy[base + t * L.st] = __ldg(var + base + t * L.st) - mean;
    }
}

bool check_grid(const mt_grid_t *g, bool cyl)
{
    if (!g) {
        mt::set_error("grid is null");
        return false;
    }
    if (g->nz < 1 || g->nt < 3 || g->nr < 3 || g->nz > 2147483647LL || g->nt > 2147483647LL ||
        g->nr > 2147483647LL) {
        mt::set_error("grid extents must satisfy nt,nr >= 3 (FD2 needs 3 points)");
        return false;
    }
    if (g->layout != MT_LAYOUT_ZTR && g->layout != MT_LAYOUT_TRZ) {
        mt::set_error("unknown layout %d", g->layout);
        return false;
    }
    if (cyl && (!g->r || !g->sin_t || !g->cos_t)) {
        mt::set_error("cylindrical grid needs r, sin_t, cos_t tables");
        return false;
    }
    if (g->halo_lo < 0 || g->halo_hi < 0 || g->halo_lo + g->halo_hi >= g->nz) {
        mt::set_error("bad halo description");
        return false;
    }
    const int64_t owned = g->nz - g->halo_lo - g->halo_hi;
    if (g->sub_lo < 0 || g->sub_hi < 0 || g->sub_hi > owned || (g->sub_hi != 0 && g->sub_lo >= g->sub_hi) ||
        (g->sub_hi == 0 && g->sub_lo != 0)) {
        mt::set_error("bad level sub-range [%lld, %lld) of %lld owned levels", (long long)g->sub_lo,
                      (long long)g->sub_hi, (long long)owned);
        return false;
    }
    return true;
}

GridP grid_params(const mt_grid_t *g)
{
    GridP p;
    p.dz = g->dz, p.dt = g->dt, p.dr = g->dr;
    p.r = g->r, p.sin_t = g->sin_t, p.cos_t = g->cos_t;
    p.halo_lo = (int)g->halo_lo, p.halo_hi = (int)g->halo_hi;
    const int64_t owned = g->nz - g->halo_lo - g->halo_hi;
    const int64_t lo = g->sub_hi ? g->sub_lo : 0, hi = g->sub_hi ? g->sub_hi : owned;
    p.own_lo = (int)(g->halo_lo + lo);
    p.own_hi = (int)(g->halo_lo + hi);
    // a sub-range that does not start / end at the shard's own end has ordinary (centred) levels there
    p.bottom = g->is_bottom != 0 && lo == 0;
    p.top = g->is_top != 0 && hi == owned;
    return p;
}

mt::th::Hydro hydro_of(const mt_dyn_in_t *in)
{
    mt::th::Hydro h;
    h.n = 0;
    for (int k = 0; k < 6; ++k) h.q[k] = nullptr;
    for (int k = 0; k < 6 && in->q[k]; ++k) h.q[h.n++] = in->q[k];
    if (h.n == 0) h.n = -1;
    return h;
}

template <bool CYL>
int launch_stage_a(const mt_grid_t *g, const GridP &p, const mt_dyn_in_t *in, const StageAOut &a, int wind,
                   cudaStream_t s)
{
    const bool any = a.vr || a.vt || a.wind_speed || a.kinetic_energy || a.Tv || a.rho || a.density_factor || a.Th ||
                     a.Th_rho || a.coriolis_force || a.centrifugal_force || a.aam;
    if (!any) return MT_OK;
    const int64_t total = g->nz * g->nt * g->nr;
    const unsigned grid = mt::grid_for(total, 256, 8);
    mt::th::Hydro hyd = hydro_of(in);
    mt::ProfScope prof(CYL ? "k_stage_a_cyl" : "k_stage_a_car", (mt_stream_t)s);
#define SA_LAUNCH(LAYOUT_, WIND_)                                                                           \
    k_stage_a<LAYOUT_, CYL, WIND_><<<grid, 256, 0, s>>>(Lay<LAYOUT_>((int)g->nz, (int)g->nt, (int)g->nr), p, *in, hyd, a)
#define SA_DISPATCH(LAYOUT_)                                       \
    do {                                                           \
        if (wind == WIND_NONE) SA_LAUNCH(LAYOUT_, WIND_NONE);      \
        else if (wind == WIND_UV) SA_LAUNCH(LAYOUT_, WIND_UV);     \
        else SA_LAUNCH(LAYOUT_, WIND_GIVEN);                       \
    } while (0)
    if (g->layout == MT_LAYOUT_ZTR) SA_DISPATCH(MT_LAYOUT_ZTR);
    else SA_DISPATCH(MT_LAYOUT_TRZ);
#undef SA_DISPATCH
#undef SA_LAUNCH
    MT_LAUNCH_CHECK();
    return MT_OK;
}

// `a`: every point-wise output requested (wind-derived and thermodynamic); `wind`: where the horizontal
// wind comes from (WIND_UV: in->u, in->v; WIND_GIVEN: bi.a, bi.b are the caller's vr, vt).
// Work split (include/meteotools_b200.h, MT_STAGE_*): the wind-derived point-wise outputs belong to
// stage A unless MT_STAGE_WIND_IN_B is set -- or both stages run in this call and the TMA stencil kernel
// can take them: it has u, v, w staged in shared memory anyway.
template <bool CYL>
int run_sets(const mt_grid_t *g, const mt_dyn_in_t *in, const StageAOut &a, const StageBIn &bi,
             const StageBOut &bo, int stages, int wind, bool need_b, cudaStream_t s)
{
    GridP p = grid_params(g);
    StageAOut thermo = a, windo = a;   // the two halves of the point-wise outputs
    thermo.vr = thermo.vt = thermo.wind_speed = thermo.kinetic_energy = nullptr;
    thermo.coriolis_force = thermo.centrifugal_force = thermo.aam = nullptr;
    windo.Tv = windo.rho = windo.density_factor = windo.Th = windo.Th_rho = nullptr;
    const bool any_wind = windo.vr || windo.vt || windo.wind_speed || windo.kinetic_energy ||
                          windo.coriolis_force || windo.centrifugal_force || windo.aam;
    const bool tma_ok = need_b && stencil_tma_supported(g);
    // the TMA kernel rotates (u, v) itself; a caller-given (vr, vt) pair needs no wind pass in it
    const bool tma_wind = tma_ok && wind == WIND_UV;
    const bool wind_in_b = (stages & MT_STAGE_WIND_IN_B) || ((stages & MT_STAGE_ALL) == MT_STAGE_ALL && tma_wind);
    int rc = MT_OK;
    if (stages & MT_STAGE_A) {
        rc = launch_stage_a<CYL>(g, p, in, wind_in_b ? thermo : a, (wind_in_b || !any_wind) ? WIND_NONE : wind, s);
        if (rc != MT_OK) return rc;
    }
    if (!(stages & MT_STAGE_B)) return MT_OK;
    bool wind_done = !(wind_in_b && any_wind);
    if (need_b) {
        const bool ztr = g->layout == MT_LAYOUT_ZTR;
        // the whole set requested -> the straight-line instantiations (MT_STAGEB_GENERIC=1: A/B runs)
        static const bool force_generic = getenv("MT_STAGEB_GENERIC") && getenv("MT_STAGEB_GENERIC")[0] == '1';
        const bool full = !force_generic && bi.a && bi.b && bi.w && bi.thr && bi.rho && bi.f && bo.zeta_r &&
                          bo.zeta_t && bo.zeta_z && bo.abs_zeta_z && bo.div_hori && bo.div && bo.PV &&
                          bo.shear_deform && bo.stretch_deform && bo.filamentation_time &&
                          (!CYL || (bi.p && bo.radial_PG_force && bo.agradient_force && bo.strain_region));
        bool interior_done = false;
        if (tma_ok) {
            WindOut wo{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            const bool fold = wind_in_b && any_wind && tma_wind;
            if (fold)
                wo = WindOut{windo.vr, windo.vt, windo.wind_speed, in->w ? windo.kinetic_energy : nullptr,
                             windo.coriolis_force, windo.centrifugal_force, windo.aam};
            const bool from_uv = tma_wind && (wind_in_b || !CYL);
            rc = stencil_tma<CYL>(g, p, bi, (CYL && from_uv) ? in->u : nullptr, (CYL && from_uv) ? in->v : nullptr, wo,
                                  bo, full, s);
            if (rc == MT_OK) {
                interior_done = true;
                if (fold) wind_done = true;
            } else if (rc != MT_ENOSUP) {
                return rc;
            }
        }
        if (!wind_done) {   // the marching kernels read vr / vt from global memory
            rc = launch_stage_a<CYL>(g, p, in, windo, wind, s);
            if (rc != MT_OK) return rc;
            wind_done = true;
        }
        const int64_t plane = ztr ? g->nt * g->nr : g->nr * g->nz;
        const int64_t n_march = ztr ? (p.own_hi - p.own_lo) : g->nt;
        // enough segments of the march axis for ~2 full waves of 128-thread CTAs
        int64_t want_threads = (int64_t)mt::sm_count() * 2048 * 2;
        int64_t nseg = (want_threads + plane - 1) / plane;
        if (nseg > n_march / 4) nseg = n_march / 4;
        if (nseg < 1) nseg = 1;
        const int seg_len = (int)((n_march + nseg - 1) / nseg);
        dim3 gridb((unsigned)((plane + 127) / 128), (unsigned)((n_march + seg_len - 1) / seg_len));
        // shell launch: one thread per point of the outermost shell of the owned block
        const int64_t own = p.own_hi - p.own_lo;
        const int64_t nzin = own - (p.bottom ? 1 : 0) - (p.top ? 1 : 0);
        const int64_t n_shell = ((p.bottom ? 1 : 0) + (p.top ? 1 : 0)) * g->nt * g->nr +
                                2 * (nzin > 0 ? nzin : 0) * (g->nr + g->nt - 2);
        // after the TMA kernel only the two end planes of the march axis are left (mode 2 of the shell kernel)
        const int64_t n_ends = ((ztr ? (p.bottom ? 1 : 0) + (p.top ? 1 : 0) : 2)) * plane;
        const int64_t n_sh = interior_done ? n_ends : n_shell;
        const int shell_mode = interior_done ? 2 : 1;
        dim3 grids((unsigned)((n_sh + 127) / 128 > 0 ? (n_sh + 127) / 128 : 1));
#define SB_LAUNCH(LAYOUT_)                                                                             \
    do {                                                                                               \
        Lay<LAYOUT_> lay((int)g->nz, (int)g->nt, (int)g->nr);                                          \
        if (!interior_done) {                                                                          \
            mt::ProfScope prof(CYL ? "k_stage_b_cyl" : "k_stage_b_car", (mt_stream_t)s);               \
            if (full) k_stage_b<LAYOUT_, CYL, false, true><<<gridb, 128, 0, s>>>(lay, p, bi, bo, seg_len); \
            else k_stage_b<LAYOUT_, CYL, false, false><<<gridb, 128, 0, s>>>(lay, p, bi, bo, seg_len);  \
            MT_LAUNCH_CHECK();                                                                         \
        }                                                                                              \
        mt::ProfScope prof(CYL ? "k_stage_b_shell_cyl" : "k_stage_b_shell_car", (mt_stream_t)s);       \
        if (full) k_stage_b<LAYOUT_, CYL, true, true><<<grids, 128, 0, s>>>(lay, p, bi, bo, shell_mode);   \
        else k_stage_b<LAYOUT_, CYL, true, false><<<grids, 128, 0, s>>>(lay, p, bi, bo, shell_mode);     \
        MT_LAUNCH_CHECK();                                                                             \
    } while (0)
        if (ztr) SB_LAUNCH(MT_LAYOUT_ZTR);
        else SB_LAUNCH(MT_LAYOUT_TRZ);
#undef SB_LAUNCH
    }
    if (!wind_done) {   // MT_STAGE_WIND_IN_B without any stencil output
        rc = launch_stage_a<CYL>(g, p, in, windo, wind, s);
        if (rc != MT_OK) return rc;
    }
    return MT_OK;
}

}  // namespace

extern "C" {

int mt_cyl_d(const mt_grid_t *g, const mt_dyn_in_t *in, const mt_cyl_d_out_t *out, int stages,
             mt_stream_t stream)
{
    if (!check_grid(g, true)) return MT_EINVAL;
    MT_REQUIRE(in && out, "mt_cyl_d: null argument");
    MT_REQUIRE((stages & MT_STAGE_ALL) != 0 && (stages & ~(MT_STAGE_ALL | MT_STAGE_WIND_IN_B)) == 0,
               "mt_cyl_d: bad stages %d", stages);
    const mt_cyl_d_out_t &o = *out;
    // the caller's vr / vt are used as they are; one of them alone is enough when no requested output reads
    // the other (calc_agradient_force needs vt only, cylindrical_grid.py:335-348)
    const bool given = in->vr || in->vt;
    const bool reads_vr = o.zeta_t || o.zeta_z || o.abs_zeta_z || o.div_hori || o.div || o.PV || o.shear_deform ||
                          o.stretch_deform || o.filamentation_time || o.strain_region || o.kinetic_energy;
    const bool reads_vt = o.zeta_r || o.zeta_z || o.abs_zeta_z || o.div_hori || o.div || o.PV || o.shear_deform ||
                          o.stretch_deform || o.filamentation_time || o.strain_region || o.kinetic_energy ||
                          o.coriolis_force || o.centrifugal_force || o.agradient_force || o.aam;
    MT_REQUIRE(!given || ((!reads_vr || in->vr) && (!reads_vt || in->vt)),
               "mt_cyl_d: the requested outputs read both vr and vt, only one was given");
    MT_REQUIRE(!given || (!o.vr && !o.vt), "mt_cyl_d: vr / vt are inputs of this call, they cannot be outputs too");
    const bool have_uv = in->u && in->v;
    const bool need_ab = o.zeta_r || o.zeta_t || o.zeta_z || o.abs_zeta_z || o.div_hori || o.div || o.PV ||
                         o.shear_deform || o.stretch_deform || o.filamentation_time || o.strain_region;
    const bool need_b = need_ab || o.radial_PG_force || o.agradient_force;
    const bool wind_pw = o.vr || o.vt || o.kinetic_energy || o.coriolis_force || o.centrifugal_force || o.aam;
    const bool b_only = (stages & MT_STAGE_ALL) == MT_STAGE_B && !(stages & MT_STAGE_WIND_IN_B);
    // stage B alone reads the horizontal wind from the vr / vt buffers of an earlier stage-A call
    // (a component no output reads is still loaded by the kernels: it may alias the other one)
    const double *vr_src = given ? (in->vr ? in->vr : in->vt) : o.vr, *vt_src = given ? (in->vt ? in->vt : in->vr) : o.vt;
    MT_REQUIRE(!o.wind_speed || have_uv, "mt_cyl_d: wind_speed needs u and v");
    MT_REQUIRE(!wind_pw || given || have_uv, "mt_cyl_d: u and v (or vr and vt) are required");
    MT_REQUIRE(!((stages & MT_STAGE_B) && need_ab) || (vr_src && vt_src),
               "mt_cyl_d: stencil outputs need vr and vt (inputs, or output buffers filled from u, v)");
    MT_REQUIRE(!((stages & MT_STAGE_B) && need_ab) || given || b_only || have_uv,
               "mt_cyl_d: u and v (or vr and vt) are required");
    MT_REQUIRE(!((stages & MT_STAGE_B) && o.agradient_force) || vt_src, "mt_cyl_d: agradient_force needs vt");
    const bool z_needed = o.zeta_r || o.zeta_t || o.div || o.PV;
    MT_REQUIRE(!z_needed || (g->nz - g->halo_lo - g->halo_hi >= 1 && g->nz >= 3),
               "mt_cyl_d: z-derivatives need at least 3 local levels");
    MT_REQUIRE(!z_needed || in->w, "mt_cyl_d: w is required for the requested outputs");
    MT_REQUIRE(!(o.PV) || (o.Th_rho && o.rho && in->f), "mt_cyl_d: PV needs Th_rho, rho buffers and f");
    MT_REQUIRE(!(o.radial_PG_force || o.agradient_force) || (o.rho && in->p),
               "mt_cyl_d: the pressure-gradient force needs p and the rho buffer");
    MT_REQUIRE(!(o.abs_zeta_z || o.agradient_force || o.coriolis_force || o.aam) || in->f, "mt_cyl_d: f is required");
    const bool thermo = o.Tv || o.rho || o.density_factor || o.Th || o.Th_rho;
    const bool thermo_now = thermo && (stages & MT_STAGE_A);
    MT_REQUIRE(!thermo_now || in->qv, "mt_cyl_d: qv is required for Tv/rho/density_factor/Th_rho");
    MT_REQUIRE(!(thermo_now && (o.Tv || o.rho)) || (in->T && in->p), "mt_cyl_d: Tv/rho need T and p");
    MT_REQUIRE(!(thermo_now && (o.Th || o.Th_rho)) || in->Th || (in->T && in->p), "mt_cyl_d: Th needs Th or (T, p)");
    StageAOut a{o.vr, o.vt, o.wind_speed, o.kinetic_energy, o.Tv, o.rho, o.density_factor, o.Th,
                o.Th_rho, o.coriolis_force, o.centrifugal_force, o.aam};
    StageBIn bi{vr_src, vt_src, in->w, o.Th_rho, in->p, o.rho, in->f};
    StageBOut bo{o.radial_PG_force, o.agradient_force, o.zeta_r, o.zeta_t, o.zeta_z, o.abs_zeta_z,
                 o.div_hori, o.div, o.PV, o.shear_deform, o.stretch_deform, o.filamentation_time,
                 o.strain_region};
    // WIND_GIVEN also describes stage B alone on the vr / vt buffers of an earlier call: nothing is rotated
    const int wind = (given || b_only) ? WIND_GIVEN : WIND_UV;
    mt_dyn_in_t in2 = *in;
    if (wind == WIND_GIVEN) in2.vr = vr_src, in2.vt = vt_src;
    if (b_only)   // the point-wise outputs of an earlier stage-A call are inputs here
        a = StageAOut{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                      nullptr, nullptr};
    return run_sets<true>(g, &in2, a, bi, bo, stages, wind, need_b, (cudaStream_t)stream);
}

int mt_car_d(const mt_grid_t *g, const mt_dyn_in_t *in, const mt_car_d_out_t *out, int stages,
             mt_stream_t stream)
{
    if (!check_grid(g, false)) return MT_EINVAL;
    MT_REQUIRE(in && out && in->u && in->v, "mt_car_d: u and v are required");
    MT_REQUIRE((stages & MT_STAGE_ALL) != 0 && (stages & ~(MT_STAGE_ALL | MT_STAGE_WIND_IN_B)) == 0,
               "mt_car_d: bad stages %d", stages);
    const mt_car_d_out_t &o = *out;
    const bool b_uses_w = o.zeta_r || o.zeta_t || o.div || o.PV;
    MT_REQUIRE(!b_uses_w || (in->w && g->nz >= 3), "mt_car_d: w and >= 3 local levels are required");
    MT_REQUIRE(!(o.PV) || (o.Th_rho && o.rho && in->f), "mt_car_d: PV needs Th_rho, rho buffers and f");
    MT_REQUIRE(!(o.abs_zeta_z) || in->f, "mt_car_d: f is required");
    const bool thermo = o.Tv || o.rho || o.density_factor || o.Th || o.Th_rho;
    const bool thermo_now = thermo && (stages & MT_STAGE_A);
    MT_REQUIRE(!thermo_now || in->qv, "mt_car_d: qv is required for Tv/rho/density_factor/Th_rho");
    MT_REQUIRE(!(thermo_now && (o.Tv || o.rho)) || (in->T && in->p), "mt_car_d: Tv/rho need T and p");
    MT_REQUIRE(!(thermo_now && (o.Th || o.Th_rho)) || in->Th || (in->T && in->p), "mt_car_d: Th needs Th or (T, p)");
    const bool need_b = o.zeta_r || o.zeta_t || o.zeta_z || o.abs_zeta_z || o.div_hori || o.div || o.PV ||
                        o.shear_deform || o.stretch_deform || o.filamentation_time;
    const bool b_only = (stages & MT_STAGE_ALL) == MT_STAGE_B && !(stages & MT_STAGE_WIND_IN_B);
    StageAOut a{nullptr, nullptr, o.wind_speed, o.kinetic_energy, o.Tv, o.rho, o.density_factor, o.Th,
                o.Th_rho, nullptr, nullptr, nullptr};
    if (b_only) a.wind_speed = a.kinetic_energy = nullptr;
    StageBIn bi{in->u, in->v, in->w, o.Th_rho, in->p, o.rho, in->f};
    StageBOut bo{nullptr, nullptr, o.zeta_r, o.zeta_t, o.zeta_z, o.abs_zeta_z, o.div_hori, o.div, o.PV,
                 o.shear_deform, o.stretch_deform, o.filamentation_time, nullptr};
    return run_sets<false>(g, in, a, bi, bo, stages, WIND_UV, need_b, (cudaStream_t)stream);
}

int mt_azimuthal_mean(const double *var, const mt_grid_t *g, double *sym, double *asy,
                      mt_stream_t stream)
{
    MT_REQUIRE(var && sym && g && g->nz >= 1 && g->nt >= 1 && g->nr >= 1, "mt_azimuthal_mean: bad arguments");
    const unsigned grid = mt::grid_for(g->nz * g->nr, 128, 8);
    mt::ProfScope prof("k_azimuthal", stream);
    if (g->layout == MT_LAYOUT_ZTR)
        k_azimuthal<MT_LAYOUT_ZTR><<<grid, 128, 0, (cudaStream_t)stream>>>(
            Lay<MT_LAYOUT_ZTR>((int)g->nz, (int)g->nt, (int)g->nr), var, sym, asy);
    else if (g->layout == MT_LAYOUT_TRZ)
        k_azimuthal<MT_LAYOUT_TRZ><<<grid, 128, 0, (cudaStream_t)stream>>>(
            Lay<MT_LAYOUT_TRZ>((int)g->nz, (int)g->nt, (int)g->nr), var, sym, asy);
    else {
        mt::set_error("mt_azimuthal_mean: unknown layout");
        return MT_EINVAL;
    }
    MT_LAUNCH_CHECK();
    return MT_OK;
}

}  // extern "C"
