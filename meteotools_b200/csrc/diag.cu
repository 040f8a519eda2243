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

#include "fd.cuh"
#include "thermo.cuh"

namespace {

using mt::nan_f64;
using mt::inf_f64;

template <int LAYOUT>
struct Lay {
    int nz, nt, nr;
    int64_t sz, st, sr;
    __host__ __device__ Lay(int nz_, int nt_, int nr_) : nz(nz_), nt(nt_), nr(nr_)
    {
        if (LAYOUT == MT_LAYOUT_ZTR) {
            sr = 1, st = nr, sz = (int64_t)nt * nr;
        } else {
            sz = 1, sr = nz, st = (int64_t)nr * nz;
        }
    }
    __device__ __forceinline__ void decode(int64_t m, int &iz, int &it, int &ir) const
    {
        if (LAYOUT == MT_LAYOUT_ZTR) {
            ir = (int)(m % nr);
            const int64_t q = m / nr;
            it = (int)(q % nt);
            iz = (int)(q / nt);
        } else {
            iz = (int)(m % nz);
            const int64_t q = m / nz;
            ir = (int)(q % nr);
            it = (int)(q / nr);
        }
    }
};

struct GridP {
    double dz, dt, dr;
    const double *r, *sin_t, *cos_t;
    int own_lo, own_hi;  // owned local level range [own_lo, own_hi)
    bool bottom, top;
};

struct StageAOut {
    double *vr, *vt, *wind_speed, *kinetic_energy, *Tv, *rho, *density_factor, *Th, *Th_rho,
        *coriolis_force, *centrifugal_force, *aam;
};

// ------------------------------------------------------------------------------- stage A
template <int LAYOUT, bool CYL>
__global__ void __launch_bounds__(256)
k_stage_a(Lay<LAYOUT> L, GridP g, mt_dyn_in_t in, mt::th::Hydro hyd, StageAOut o)
{
    const int64_t total = (int64_t)L.nz * L.nt * L.nr;
    for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < total;
         m += (int64_t)gridDim.x * blockDim.x) {
        int iz, it, ir;
        L.decode(m, iz, it, ir);
        const double u = __ldg(in.u + m), v = __ldg(in.v + m);
        double vr = 0, vt = 0;
        if (CYL) {
            const double s = __ldg(g.sin_t + it), c = __ldg(g.cos_t + it);
            vr = v * s + u * c;  // cylindrical_grid.py:267
            vt = v * c - u * s;  // :268
            if (o.vr) o.vr[m] = vr;
            if (o.vt) o.vt[m] = vt;
        }
        if (o.wind_speed) __stcs(o.wind_speed + m, sqrt(u * u + v * v));  // (u**2+v**2)**0.5
        if (o.kinetic_energy && in.w) {
            const double w = __ldg(in.w + m);
            const double ke = CYL ? (vr * vr + vt * vt + w * w) / 2 : (u * u + v * v + w * w) / 2;
            __stcs(o.kinetic_energy + m, ke);
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
        if (CYL && (o.coriolis_force || o.centrifugal_force || o.aam)) {
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
struct StageBIn {
    const double *a, *b, *w, *thr, *p, *rho, *f;  // cyl: a=vr, b=vt ; car: a=u, b=v
};
struct StageBOut {
    double *radial_PG_force, *agradient_force;
    double *zeta_r, *zeta_t, *zeta_z, *abs_zeta_z, *div_hori, *div, *PV;
    double *shear_deform, *stretch_deform, *filamentation_time;
    uint8_t *strain_region;
};

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
__global__ void __launch_bounds__(128, SHELL ? 4 : MT_STAGEB_MINB)
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
        // compact enumeration of the outermost shell of the owned block: the global z faces, the two
        // theta faces of the remaining levels, the two r faces of what is left -- one thread per
        // shell point (the first version launched a thread per GRID point and let 95 % of them
        // leave at once: 0.26 ms of the 0.63 ms of stage B on S3)
        const int zin_lo = g.own_lo + (g.bottom ? 1 : 0), zin_hi = g.own_hi - (g.top ? 1 : 0);
        const int nzin = max(zin_hi - zin_lo, 0);
        const int64_t nA = (int64_t)((g.bottom ? 1 : 0) + (g.top ? 1 : 0)) * nt * nr;
        const int64_t nB = 2LL * nzin * nr, nC = 2LL * nzin * (nt - 2);
        int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        int iz, it, ir;
        if (q < nA) {
            const int f = (int)(q / ((int64_t)nt * nr));
            const int64_t rem = q - (int64_t)f * nt * nr;
            it = (int)(rem / nr), ir = (int)(rem - (int64_t)it * nr);
            iz = (f == 0 && g.bottom) ? 0 : nz - 1;
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
    return true;
}

GridP grid_params(const mt_grid_t *g)
{
    GridP p;
    p.dz = g->dz, p.dt = g->dt, p.dr = g->dr;
    p.r = g->r, p.sin_t = g->sin_t, p.cos_t = g->cos_t;
    p.own_lo = (int)g->halo_lo;
    p.own_hi = (int)(g->nz - g->halo_hi);
    p.bottom = g->is_bottom != 0;
    p.top = g->is_top != 0;
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
int run_sets(const mt_grid_t *g, const mt_dyn_in_t *in, const StageAOut &a, const StageBIn &bi,
             const StageBOut &bo, int stages, bool need_a, bool need_b, cudaStream_t s)
{
    const int64_t total = g->nz * g->nt * g->nr;
    const unsigned grid = mt::grid_for(total, 256, 8);
    GridP p = grid_params(g);
    mt::th::Hydro hyd = hydro_of(in);
    if ((stages & 1) && need_a) {
        mt::ProfScope prof(CYL ? "k_stage_a_cyl" : "k_stage_a_car", (mt_stream_t)s);
        if (g->layout == MT_LAYOUT_ZTR)
            k_stage_a<MT_LAYOUT_ZTR, CYL><<<grid, 256, 0, s>>>(Lay<MT_LAYOUT_ZTR>((int)g->nz, (int)g->nt, (int)g->nr), p, *in, hyd, a);
        else
            k_stage_a<MT_LAYOUT_TRZ, CYL><<<grid, 256, 0, s>>>(Lay<MT_LAYOUT_TRZ>((int)g->nz, (int)g->nt, (int)g->nr), p, *in, hyd, a);
        MT_LAUNCH_CHECK();
    }
    if ((stages & 2) && need_b) {
        mt::ProfScope prof(CYL ? "k_stage_b_cyl" : "k_stage_b_car", (mt_stream_t)s);
        const bool ztr = g->layout == MT_LAYOUT_ZTR;
        const int64_t plane = ztr ? g->nt * g->nr : g->nr * g->nz;
        const int64_t n_march = ztr ? (g->nz - g->halo_lo - g->halo_hi) : g->nt;
        // enough segments of the march axis for ~2 full waves of 128-thread CTAs
        int64_t want_threads = (int64_t)mt::sm_count() * 2048 * 2;
        int64_t nseg = (want_threads + plane - 1) / plane;
        if (nseg > n_march / 4) nseg = n_march / 4;
        if (nseg < 1) nseg = 1;
        const int seg_len = (int)((n_march + nseg - 1) / nseg);
        dim3 gridb((unsigned)((plane + 127) / 128), (unsigned)((n_march + seg_len - 1) / seg_len));
        // shell launch: one thread per point of the outermost shell of the owned block
        const int64_t own = g->nz - g->halo_lo - g->halo_hi;
        const int64_t nzin = own - (g->is_bottom ? 1 : 0) - (g->is_top ? 1 : 0);
        const int64_t n_shell = ((g->is_bottom ? 1 : 0) + (g->is_top ? 1 : 0)) * g->nt * g->nr +
                                2 * (nzin > 0 ? nzin : 0) * (g->nr + g->nt - 2);
        dim3 grids((unsigned)((n_shell + 127) / 128 > 0 ? (n_shell + 127) / 128 : 1));
        // the whole set requested -> the straight-line instantiation (MT_STAGEB_GENERIC=1: A/B runs)
        static const bool force_generic = getenv("MT_STAGEB_GENERIC") && getenv("MT_STAGEB_GENERIC")[0] == '1';
        const bool full = !force_generic && bi.a && bi.b && bi.w && bi.thr && bi.rho && bi.f && bo.zeta_r &&
                          bo.zeta_t && bo.zeta_z && bo.abs_zeta_z && bo.div_hori && bo.div && bo.PV &&
                          bo.shear_deform && bo.stretch_deform && bo.filamentation_time &&
                          (!CYL || (bi.p && bo.radial_PG_force && bo.agradient_force && bo.strain_region));
#define SB_LAUNCH(LAYOUT_)                                                                         \
    do {                                                                                           \
        Lay<LAYOUT_> lay((int)g->nz, (int)g->nt, (int)g->nr);                                      \
        if (full) k_stage_b<LAYOUT_, CYL, false, true><<<gridb, 128, 0, s>>>(lay, p, bi, bo, seg_len); \
        else k_stage_b<LAYOUT_, CYL, false, false><<<gridb, 128, 0, s>>>(lay, p, bi, bo, seg_len);  \
        MT_LAUNCH_CHECK();                                                                         \
        k_stage_b<LAYOUT_, CYL, true, false><<<grids, 128, 0, s>>>(lay, p, bi, bo, 1);             \
    } while (0)
        if (ztr) SB_LAUNCH(MT_LAYOUT_ZTR);
        else SB_LAUNCH(MT_LAYOUT_TRZ);
#undef SB_LAUNCH
        MT_LAUNCH_CHECK();
    }
    return MT_OK;
}

}  // namespace

extern "C" {

int mt_cyl_d(const mt_grid_t *g, const mt_dyn_in_t *in, const mt_cyl_d_out_t *out, int stages,
             mt_stream_t stream)
{
    if (!check_grid(g, true)) return MT_EINVAL;
    MT_REQUIRE(in && out && in->u && in->v, "mt_cyl_d: u and v are required");
    const mt_cyl_d_out_t &o = *out;
    const bool z_needed = o.zeta_r || o.zeta_t || o.div || o.PV;
    MT_REQUIRE(!z_needed || (g->nz - g->halo_lo - g->halo_hi >= 1 &&
                             g->nz >= 3),
               "mt_cyl_d: z-derivatives need at least 3 local levels");
    const bool b_uses_w = o.zeta_r || o.zeta_t || o.div || o.PV;
    MT_REQUIRE(!b_uses_w || in->w, "mt_cyl_d: w is required for the requested outputs");
    const bool need_b = o.radial_PG_force || o.agradient_force || o.zeta_r || o.zeta_t || o.zeta_z ||
                        o.abs_zeta_z || o.div_hori || o.div || o.PV || o.shear_deform ||
                        o.stretch_deform || o.filamentation_time || o.strain_region;
    const bool need_ab = o.zeta_r || o.zeta_t || o.zeta_z || o.abs_zeta_z || o.div_hori || o.div || o.PV ||
                         o.shear_deform || o.stretch_deform || o.filamentation_time || o.strain_region;
    MT_REQUIRE(!((stages & MT_STAGE_B) && need_ab) || (o.vr && o.vt),
               "mt_cyl_d: stencil outputs need the vr and vt buffers");
    MT_REQUIRE(!((stages & MT_STAGE_B) && o.agradient_force) || o.vt, "mt_cyl_d: agradient_force needs vt");
    MT_REQUIRE(!(o.PV) || (o.Th_rho && o.rho && in->f), "mt_cyl_d: PV needs Th_rho, rho buffers and f");
    MT_REQUIRE(!(o.radial_PG_force || o.agradient_force) || (o.rho && in->p),
               "mt_cyl_d: the pressure-gradient force needs p and the rho buffer");
    MT_REQUIRE(!(o.abs_zeta_z || o.agradient_force || o.coriolis_force || o.aam) || in->f, "mt_cyl_d: f is required");
    const bool thermo = o.Tv || o.rho || o.density_factor || o.Th || o.Th_rho;
    MT_REQUIRE(!thermo || in->qv, "mt_cyl_d: qv is required for Tv/rho/density_factor/Th_rho");
    MT_REQUIRE(!(o.Tv || o.rho) || (in->T && in->p), "mt_cyl_d: Tv/rho need T and p");
    MT_REQUIRE(!(o.Th || o.Th_rho) || in->Th || (in->T && in->p), "mt_cyl_d: Th needs Th or (T, p)");
    StageAOut a{o.vr, o.vt, o.wind_speed, o.kinetic_energy, o.Tv, o.rho, o.density_factor, o.Th,
                o.Th_rho, o.coriolis_force, o.centrifugal_force, o.aam};
    StageBIn bi{o.vr, o.vt, in->w, o.Th_rho, in->p, o.rho, in->f};
    StageBOut bo{o.radial_PG_force, o.agradient_force, o.zeta_r, o.zeta_t, o.zeta_z, o.abs_zeta_z,
                 o.div_hori, o.div, o.PV, o.shear_deform, o.stretch_deform, o.filamentation_time,
                 o.strain_region};
    return run_sets<true>(g, in, a, bi, bo, stages, true, need_b, (cudaStream_t)stream);
}

int mt_car_d(const mt_grid_t *g, const mt_dyn_in_t *in, const mt_car_d_out_t *out, int stages,
             mt_stream_t stream)
{
    if (!check_grid(g, false)) return MT_EINVAL;
    MT_REQUIRE(in && out && in->u && in->v, "mt_car_d: u and v are required");
    const mt_car_d_out_t &o = *out;
    const bool b_uses_w = o.zeta_r || o.zeta_t || o.div || o.PV;
    MT_REQUIRE(!b_uses_w || (in->w && g->nz >= 3), "mt_car_d: w and >= 3 local levels are required");
    MT_REQUIRE(!(o.PV) || (o.Th_rho && o.rho && in->f), "mt_car_d: PV needs Th_rho, rho buffers and f");
    MT_REQUIRE(!(o.abs_zeta_z) || in->f, "mt_car_d: f is required");
    const bool thermo = o.Tv || o.rho || o.density_factor || o.Th || o.Th_rho;
    MT_REQUIRE(!thermo || in->qv, "mt_car_d: qv is required for Tv/rho/density_factor/Th_rho");
    MT_REQUIRE(!(o.Tv || o.rho) || (in->T && in->p), "mt_car_d: Tv/rho need T and p");
    MT_REQUIRE(!(o.Th || o.Th_rho) || in->Th || (in->T && in->p), "mt_car_d: Th needs Th or (T, p)");
    const bool need_a = o.wind_speed || o.kinetic_energy || thermo;
    const bool need_b = o.zeta_r || o.zeta_t || o.zeta_z || o.abs_zeta_z || o.div_hori || o.div || o.PV ||
                        o.shear_deform || o.stretch_deform || o.filamentation_time;
    StageAOut a{nullptr, nullptr, o.wind_speed, o.kinetic_energy, o.Tv, o.rho, o.density_factor, o.Th,
                o.Th_rho, nullptr, nullptr, nullptr};
    StageBIn bi{in->u, in->v, in->w, o.Th_rho, in->p, o.rho, in->f};
    StageBOut bo{nullptr, nullptr, o.zeta_r, o.zeta_t, o.zeta_z, o.abs_zeta_z, o.div_hori, o.div, o.PV,
                 o.shear_deform, o.stretch_deform, o.filamentation_time, nullptr};
    return run_sets<false>(g, in, a, bi, bo, stages, need_a, need_b, (cudaStream_t)stream);
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
