// Stage B of the fused dynamics diagnostics (cyl_d: grids/cylindrical_grid.py:264-508, car_d:
// grids/cartesian_grid.py:232-364) as a warp-specialised TMA PLANE PIPELINE.
//
// The register-marching kernel of csrc/diag.cu fetches every stencil neighbour through L1/L2 and is
// bound by the latency of those loads (profiles/experiments_r01.md).  Here no consumer thread ever
// waits on global memory:
//   * a CTA owns a 2-D tile (TR rows x 32 contiguous points) of the plane spanned by the two fastest
//     memory axes and marches along the slowest one ("ZTR" arrays: z; "TRZ" arrays: theta);
//   * one producer thread streams the planes of the tile -- one 3-D TMA box {36, TR + 2, 1} per
//     differentiated field (u|vr, v|vt, w, theta_rho, p), i.e. the tile with its one-point halo (two
//     columns on the contiguous axis: a TMA box must start 16-byte aligned), and the {32, TR, 1} tile of
//     rho -- through a ring of NS = 5 plane stages guarded by full / empty mbarriers.  While plane k is
//     worked on, plane k + 1 has landed and planes k + 2, k + 3 are in flight; parts of a box outside the array
//     are zero-filled by the TMA unit and never used;
//   * TR consumer warps (one point per thread and plane) take all 3 x 3 x 3 stencil operands from shared
//     memory: the march-axis difference from the stages of planes k - 1 / k + 1, the two in-plane
//     differences from the halo of plane k.
// Folded into the same pass, because u, v, w are staged anyway: the cylindrical wind rotation
// (vr, vt; done once per landed plane IN shared memory, halo included, so that every stencil sees
// vr / vt), wind speed, kinetic energy, Coriolis / centrifugal terms and the absolute angular momentum.
// With the point-wise thermodynamic pass (k_stage_a: Tv, rho, density factor, theta_rho) the whole
// set reads every input once (+ theta_rho and rho a second time) and writes every output once.
//
// Only INTERIOR points (centred differences) are produced here; the outermost shell of the grid, with
// the one-sided formulas of calc/differential.py:30-34, stays with k_stage_b<SHELL> (csrc/diag.cu).
// Every expression is the one of k_stage_b, in the same order: results are bit-identical.
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "diag.cuh"
#include "tma.cuh"

namespace mt {
namespace dg {
namespace {

using namespace mt::tma;

constexpr int TC = 32;             // tile points along the contiguous axis = lanes of a warp
constexpr int HX = 2, HY = 1;      // halo: 2 columns (16-byte aligned box start), 1 row

// NS: plane stages of the ring.  Three are in use by the stencil (k - 1, k, k + 1); the others are the
// prefetch distance.  Tile height, ring depth and resident CTAs were swept together at the end of round 2
// (profiles/experiments_r02.md section 10; Cartesian stencil on 80 x 1000 x 1000 / S5, ms):
//   8 rows, 2 CTAs / SM, NS = 4 (until then) 2.13 / 8.73     8 rows, 2 CTAs, NS = 5  1.93 / 8.00     NS = 6  2.29 / 9.31
//   4 rows, 4 CTAs / SM, NS = 4              1.94 / 7.98     4 rows, 4 CTAs, NS = 5  1.86 / 7.5-7.6  NS = 6  2.15 / 8.11
//   4 rows, 5 CTAs (72 registers) 2.02 / 7.90   2 rows, 8 CTAs 2.28 / 8.27   6 rows, 3 CTAs 1.91 / 7.72   NS = 3  2.92 / 11.3
// Four small CTAs per SM are four independent pipelines (their waits for a landing plane fall at different
// times) and a fifth stage requests a plane two steps ahead; deeper rings cost a resident CTA.  The cylindrical
// kernel (FP64 / issue bound) does not care: 0.375-0.380 ms in every shape.
// MSH (0, 16 or 24): the tile's columns are SHIFTED per row so that every warp's 32 doubles start on a
// 256-BYTE boundary of the output arrays -- on B200 a warp-wide store that straddles a 256-byte block costs
// ~20 % of the write bandwidth (scratch/stream_probe.cu: 12 write streams run at 6.0 TB/s aligned, 4.7 TB/s
// when offset by 64 or 128 bytes; profiles/experiments_r02.md).  Row `gr` of a ZTR array starts at element
// gr*n2 (planes are multiples of 32 elements): its lanes take the columns c0 - sh .. c0 - sh + 31 with
// sh = (gr*n2) mod 32 <= MSH, and the box is MSH columns wider.  The geometry is a compile-time constant:
// every shared-memory offset of the stencil folds into the immediate field of its LDS (a run-time box pitch
// cost the cylindrical kernel 15 %).
template <int TR, bool CYL, int NS, int MSH>
struct Geo {
    static constexpr int BX = TC + 2 * HX + MSH;                 // columns per box row
    static constexpr int BY = TR + 2 * HY;
    static constexpr int BOX = BX * BY;                          // elements of a halo box
    static constexpr int BOXD = ((BOX * 8 + 127) & ~127) / 8;    // doubles, 128-byte aligned (TMA destination)
    static constexpr int NHALO = CYL ? 5 : 4;                    // a, b, w, theta_rho (, p)
    static constexpr int RW = TC + MSH;                          // columns of the rho tile
    static constexpr int RHOD = ((RW * TR * 8 + 127) & ~127) / 8;
    static constexpr int STAGED = NHALO * BOXD + RHOD;           // doubles per plane stage
    static constexpr int NCONS = 32 * TR, NTHR = NCONS + 32;
    static constexpr int FA = 0, FB = BOXD, FW = 2 * BOXD, FT = 3 * BOXD, FP = 4 * BOXD, FRHO = NHALO * BOXD;
    static constexpr size_t SMEM = (size_t)NS * STAGED * 8 + 2 * NS * 8 + NS * 4 + 128;
};

// Operands of FD2 along one in-plane axis for one thread, as element offsets from its point (step = 1 along
// the contiguous axis, the box pitch along the row axis): interior  (v[oa] - v[ob])/2  with (oa, ob) = (+1, -1);
// first point of the axis: oa, ob, oc = 0, +1, +2 and c3 = -3; last point: 0, -1, -2 and c3 = +3.
struct Edge {
    int oa, ob, oc;
    double c3;
    bool edge, first;
    __device__ __forceinline__ Edge(int i, int n, int step)
    {
        first = i == 0, edge = i == 0 || i >= n - 1;
        const int dir = first ? step : -step;
        oa = edge ? 0 : step, ob = edge ? dir : -step, oc = edge ? 2 * dir : 0;
        c3 = first ? -3.0 : 3.0;
    }
};

struct __align__(64) Maps {
    CUtensorMap a, b, w, thr, p, rho;
};

enum : unsigned { K_W = 1, K_THR = 2, K_RHO = 4, K_P = 8, K_ROT = 16, K_WIND = 32 };

struct KP {
    int n0, n1, n2;        // extents: march (slowest), row, contiguous
    int m_lo, m_hi;        // march range that receives outputs
    int halo_lo, halo_hi;  // ZTR level shards: halo planes at the ends of the march axis
    int seg_len, nseg, tiles_r, tiles_c;
    int *counter;          // task counter of this launch (zero at launch)
    unsigned flags;
    double dz, dt, dr;
    const double *r, *sin_t, *cos_t, *f;
    StageBOut o;
    WindOut wo;
};

// Divisions by the grid spacings and by the radius go through mt::FDiv (fd.cuh): reciprocal once per thread,
// three FP64 instructions per quotient, a branch-free select for the special operands -- bit-identical to the
// IEEE division for every normal operand and for 0, NaN, inf, r = 0.  The Cartesian kernel gains 10 %, the
// cylindrical one (23 such quotients per point) 32 %.  Variants that BRANCH to an IEEE division per quotient
// (mt::CDiv) or redo flagged points in a second body measured 4-9 % slower than plain IEEE divisions: the
// ~2000-instruction straight-line body lives on overlapping its dependent FP64 chains, and every convergence
// region or second body costs more than the shorter sequence saves (profiles/experiments_r02.md).
// MB: resident CTAs per SM the kernel is compiled for.
template <int LAYOUT, bool CYL, bool FULL, int TR, int NS, int MB, int MSH>
__global__ void __launch_bounds__(Geo<TR, CYL, NS, MSH>::NTHR, MB)
k_stencil_tma(const __grid_constant__ Maps M, const KP p)
{
    using G = Geo<TR, CYL, NS, MSH>;
    constexpr int BX = G::BX, STAGED = G::STAGED;
    constexpr int FA = G::FA, FB = G::FB, FW = G::FW, FT = G::FT, FP = G::FP, FRHO = G::FRHO;
    constexpr bool ZTR = LAYOUT == MT_LAYOUT_ZTR;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)NS * STAGED * 8);
    uint64_t *empty = full + NS;
    // task of the plane in each stage (written by the producer before the stage's barrier is armed): tasks
    // are drawn from a global counter, so a CTA that starts late -- its SM was busy with the NCCL kernels of
    // the halo exchange -- simply takes fewer of them (a static round-robin made every other CTA wait for it)
    volatile int *slot_task = reinterpret_cast<volatile int *>(empty + NS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, TR);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int per_seg = p.tiles_r * p.tiles_c, ntask = p.nseg * per_seg;
    // first row of row-tile `tr`: the last tile is moved up so that it is full (it then overlaps its
    // neighbour, which computes the same values): the one-sided formula at the last row reads two rows
    // below it, and the box has one halo row
    auto row0 = [&](int tr) { return (tr == p.tiles_r - 1 && p.n1 > TR) ? p.n1 - TR : tr * TR; };
    const bool has_w = p.flags & K_W, has_thr = p.flags & K_THR, has_rho = p.flags & K_RHO;
    const bool has_p = CYL && (p.flags & K_P);
    const bool rotate = CYL && (p.flags & K_ROT);

    if (warp == TR) {
        // ======================================= producer =======================================
        if (lane != 0) return;
        const uint32_t bytes = (uint32_t)((2 + (has_w ? 1 : 0) + (has_thr ? 1 : 0) + (has_p ? 1 : 0)) * BX * G::BY * 8 +
                                          (has_rho ? G::RW * TR * 8 : 0));
        int slot = 0;
        uint32_t par = 1;   // a fresh mbarrier passes a wait on parity 1: the ring starts empty
        while (true) {
            const int task = atomicAdd(p.counter, 1);
            if (task >= ntask) {   // end marker: complete the phase of the next stage without data
                mbar_wait(empty + slot, par);
                slot_task[slot] = -1;
                mbar_arrive(full + slot);
                break;
            }
            const int seg = task / per_seg, rem = task - seg * per_seg;
            const int tr = rem / p.tiles_c, tc = rem - tr * p.tiles_c;
            const int s_lo = p.m_lo + seg * p.seg_len, s_hi = min(p.m_hi, s_lo + p.seg_len);
            const int j_lo = max(s_lo - 1, 0), j_hi = min(s_hi, p.n0 - 1);
            const int x = tc * TC - MSH - HX, y = row0(tr) - HY;
            for (int j = j_lo; j <= j_hi; ++j) {
                mbar_wait(empty + slot, par);
                unsigned char *st = smem + (size_t)slot * STAGED * 8;
                slot_task[slot] = task;
                mbar_expect_tx(full + slot, bytes);
                load_3d(st + FA * 8, &M.a, full + slot, x, y, j);
                load_3d(st + FB * 8, &M.b, full + slot, x, y, j);
                if (has_w) load_3d(st + FW * 8, &M.w, full + slot, x, y, j);
                if (has_thr) load_3d(st + FT * 8, &M.thr, full + slot, x, y, j);
                if (has_p) load_3d(st + FP * 8, &M.p, full + slot, x, y, j);
                if (has_rho) load_3d(st + FRHO * 8, &M.rho, full + slot, x + HX, y + HY, j);
                if (++slot == NS) slot = 0, par ^= 1u;
            }
        }
        return;
    }

    // ========================================= consumers =========================================
    const StageBOut &o = p.o;
    const int n0 = p.n0, n1 = p.n1, n2 = p.n2;
    const int64_t plane = (int64_t)n1 * n2;
    const double *sm = reinterpret_cast<const double *>(smem);
    double *smw = reinterpret_cast<double *>(smem);

    const bool want_zr = FULL || o.zeta_r || o.PV;
    const bool want_zt = FULL || o.zeta_t || o.PV;
    const bool want_zz = FULL || o.zeta_z || o.abs_zeta_z || o.PV || o.filamentation_time || o.strain_region;
    const bool want_def = FULL || o.shear_deform || o.stretch_deform || o.filamentation_time || o.strain_region;
    const bool want_div = FULL || o.div_hori || o.div;
    const bool wind_out = (p.flags & K_WIND) != 0;
    const FDiv ddz(p.dz), ddt(p.dt), ddr(p.dr);   // the host has checked the three spacings (fdiv_divisor_ok)

    int slot = 0;          // stage of the plane that lands next
    uint32_t par = 0;
    int rel_slot = 0;      // oldest stage not yet handed back to the producer
    int held = 0;          // number of stages held by this warp
    auto release_oldest = [&]() {
        if (rotate) fence_proxy_async();   // the rotation wrote this stage through the generic proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + rel_slot);
        if (++rel_slot == NS) rel_slot = 0;
        --held;
    };

    while (true) {
        mbar_wait(full + slot, par);   // the first plane of the next task, or the end marker
        const int task = slot_task[slot];
        if (task < 0) break;
        const int seg = task / per_seg, rem = task - seg * per_seg;
        const int tr = rem / p.tiles_c, tc = rem - tr * p.tiles_c;
        const int s_lo = p.m_lo + seg * p.seg_len, s_hi = min(p.m_hi, s_lo + p.seg_len);
        const int j_lo = max(s_lo - 1, 0), j_hi = min(s_hi, n0 - 1);
        const int r0 = row0(tr), c0 = tc * TC;
        const int gr = r0 + warp;
        // 256-byte aligned stores: this row's lanes start `sh` columns to the left (see Geo)
        const int sh = MSH ? (int)(((int64_t)gr * n2) & 31) : 0;
        const int gc = c0 - sh + lane;
        const bool in_dom = gr < n1 && gc >= 0 && gc < n2;
        const int ctr = (warp + HY) * BX + lane - sh + MSH + HX;   // this thread's point inside a halo box
        const int ctr_rho = warp * G::RW + lane - sh + MSH;
        // a tile that touches no edge of the plane takes the branch-free centred instantiation
        const bool tile_inside = r0 > 0 && r0 + TR < n1 && c0 - MSH > 0 && c0 + TC < n2;
        const Edge erow(gr, n1, BX), ecol(gc, n2, 1);
        const int64_t pofs = (int64_t)gr * n2 + gc;
        // radius of the point and of its two radial neighbours (r is the contiguous axis of a ZTR
        // array, the row axis of a TRZ array)
        const int ir = ZTR ? gc : gr, nr = ZTR ? n2 : n1;
        // ... and the radii that go with the three FD2 operands along r (r_a, r_b, r_c)
        double rr = 1.0, r_a = 1.0, r_b = 1.0, r_c = 1.0;
        if (CYL) {
            const Edge &er = ZTR ? ecol : erow;
            const int st = ZTR ? 1 : BX;
            auto rad = [&](int k) { return __ldg(p.r + min(max(k, 0), nr - 1)); };
            rr = rad(ir);
            r_a = rad(ir + er.oa / st), r_b = rad(ir + er.ob / st), r_c = rad(ir + er.oc / st);
        }
        const FDiv drr(rr);   // r[0] = 0: rd = inf, and a * inf is the reference's +-inf / NaN (FDiv::special)
        double f_fix = 0.0;   // ZTR: f(t, r) does not change along the march
        if (ZTR && p.f && in_dom) f_fix = __ldg(p.f + pofs);

        int slot_m = 0, slot_p = 0;   // stages of planes j - 1 and j - 2
        for (int j = j_lo; j <= j_hi; ++j) {
            if (j > j_lo) mbar_wait(full + slot, par);
            ++held;
            const int sc = slot * STAGED;
            const bool owned = j >= s_lo && j < s_hi;
            // ---------------- arrival of plane j: wind point-wise outputs, rotation ----------------
            if (wind_out || rotate) {
                const double a0 = sm[sc + FA + ctr], b0 = sm[sc + FB + ctr];
                double va = a0, vb = b0;   // (vr, vt) on a cylindrical grid, (u, v) on a Cartesian one
                if (rotate) {
                    const int it = ZTR ? min(gr, n1 - 1) : j;
                    const double s = __ldg(p.sin_t + it), c = __ldg(p.cos_t + it);
                    va = b0 * s + a0 * c;   // cylindrical_grid.py:267
                    vb = b0 * c - a0 * s;   // :268
                }
                const bool halo_plane = ZTR && (j < p.halo_lo || j >= n0 - p.halo_hi);
                if (wind_out && in_dom && (owned || halo_plane)) {
                    const int64_t m = (int64_t)j * plane + pofs;
                    const WindOut &wo = p.wo;
                    if (CYL) {
                        if (wo.vr) wo.vr[m] = va;
                        if (wo.vt) wo.vt[m] = vb;
                    }
                    if (owned) {
                        if (wo.wind_speed) __stcs(wo.wind_speed + m, sqrt(a0 * a0 + b0 * b0));  // (u**2+v**2)**0.5
                        if (wo.kinetic_energy) {
                            const double w0 = sm[sc + FW + ctr];
                            __stcs(wo.kinetic_energy + m, (va * va + vb * vb + w0 * w0) / 2);
                        }
                        if (CYL && (wo.coriolis_force || wo.centrifugal_force || wo.aam)) {
                            const double fv = ZTR ? f_fix : __ldg(p.f + (int64_t)j * n1 + gr);
                            if (wo.coriolis_force) __stcs(wo.coriolis_force + m, fv * vb);
                            if (wo.centrifugal_force) __stcs(wo.centrifugal_force + m, by(vb * vb, drr));
                            if (wo.aam) __stcs(wo.aam + m, vb * rr + 0.5 * fv * (rr * rr));  // :297
                        }
                    }
                }
                if (rotate) {
                    // (u, v) -> (vr, vt) for the whole box of plane j, halo included, in place
                    __syncwarp();
                    asm volatile("bar.sync 1, %0;" ::"n"(G::NCONS) : "memory");   // every own-point read is done
                    for (int e = tid; e < BX * G::BY; e += G::NCONS) {
                        int it = j;
                        if (ZTR) it = min(max(r0 - HY + e / BX, 0), n1 - 1);
                        const double s = __ldg(p.sin_t + it), c = __ldg(p.cos_t + it);
                        const double u_ = sm[sc + FA + e], v_ = sm[sc + FB + e];
                        smw[sc + FA + e] = v_ * s + u_ * c;
                        smw[sc + FB + e] = v_ * c - u_ * s;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(G::NCONS) : "memory");
                }
            }
            // ---------------- stencil outputs of plane m = j - 1 ----------------
            const int mm = j - 1;
            // mm < s_hi and mm <= n0 - 2 hold by construction: the march-axis difference is always centred
            // here.  IN = true: the point is interior in the plane too (branch-free centred differences);
            // IN = false: it lies on an edge of the plane -> calc/differential.py's one-sided formulas along
            // that axis, operands still from the box (its two-column / one-row halo reaches the third point).
            auto body = [&](auto in_tag) {
                constexpr bool IN = decltype(in_tag)::value;
                const int sn = sc, s0 = slot_m * STAGED, sp = slot_p * STAGED;
                const int64_t m = (int64_t)mm * plane + pofs;
                auto d_march = [&](int F, const FDiv &delta) { return by((sm[sn + F + ctr] - sm[sp + F + ctr]) / 2, delta); };
                // FD2 along an in-plane axis from the three operands of `e` (see Edge); ka / kb / kc scale them
                // (1 except for the r-weighted differences)
                auto fd_sel = [&](int F, const Edge &e, double ka, double kb, double kc, const FDiv &delta) {
                    const double A = sm[s0 + F + ctr + e.oa] * ka, B = sm[s0 + F + ctr + e.ob] * kb;
                    if (IN) return by((A - B) / 2, delta);
                    const double C = sm[s0 + F + ctr + e.oc] * kc;
                    // first: (-3*v0 + 4*v1 - v2)/2 ; last: (3*v0 - 4*v1 + v2)/2 with v1, v2 the points BEHIND
                    // (differential.py:32,34): sign flips are exact, so one branch-free sequence serves both
                    const double t1 = e.c3 * A, t2 = 4 * B;
                    const double s12 = t1 + (e.first ? t2 : -t2);
                    const double one_sided = (s12 + (e.first ? -C : C)) / 2;
                    return by(e.edge ? one_sided : (A - B) / 2, delta);
                };
                auto d_row = [&](int F, const FDiv &delta) { return fd_sel(F, erow, 1.0, 1.0, 1.0, delta); };
                auto d_col = [&](int F, const FDiv &delta) { return fd_sel(F, ecol, 1.0, 1.0, 1.0, delta); };
                auto dz = [&](int F) { return ZTR ? d_march(F, ddz) : d_col(F, ddz); };
                auto dt = [&](int F) { return ZTR ? d_row(F, ddt) : d_march(F, ddt); };
                auto dr = [&](int F) { return ZTR ? d_col(F, ddr) : d_row(F, ddr); };
                // FD2 along r of  F * r  (vt*r, :391) and of  r * F  (r*vr, :408)
                auto dr_times_r = [&](int F) { return fd_sel(F, ZTR ? ecol : erow, r_a, r_b, r_c, ddr); };
                const double fv = (FULL || p.f) ? (ZTR ? f_fix : __ldg(p.f + (int64_t)mm * n1 + gr)) : 0.0;
                const double a_c = sm[s0 + FA + ctr], b_c = sm[s0 + FB + ctr];

                double zeta_r = 0, zeta_t = 0, zeta_z = 0, abs_zz = 0;
                double da_dt = 0, db_dt = 0;
                if (want_zz || want_def || want_div) {
                    da_dt = dt(FA);
                    db_dt = dt(FB);
                }
                double db_dz = 0;
                if (want_zr || (!CYL && want_zt)) db_dz = dz(FB);
                if (want_zr) {
                    const double dw_dt = dt(FW);
                    // cyl :387-388  FD2(w,dtheta,1)/r - FD2(vt,dz,0) ; car :250  FD2(w,dy,1) - FD2(v,dz,0)
                    zeta_r = CYL ? by(dw_dt, drr) - db_dz : dw_dt - db_dz;
                    if (FULL || o.zeta_r) __stcs(o.zeta_r + m, zeta_r);
                }
                if (want_zt) {
                    const double dw_dr = dr(FW);
                    if (CYL) zeta_t = dz(FA) - dw_dr;   // :389-390
                    else zeta_t = db_dz - dw_dr;           // cartesian_grid.py:251 (sic: v, not u)
                    if (FULL || o.zeta_t) __stcs(o.zeta_t + m, zeta_t);
                }
                double da_dr = 0, db_dr = 0;
                if (!CYL && (want_zz || want_def || want_div)) {
                    da_dr = dr(FA);
                    db_dr = dr(FB);
                }
                if (want_zz) {
                    if (CYL) zeta_z = by(dr_times_r(FB), drr) - by(da_dt, drr);   // :391-392
                    else zeta_z = db_dr - da_dt;                             // :252 FD2(v,dx,2) - FD2(u,dy,1)
                    if (FULL || o.zeta_z) __stcs(o.zeta_z + m, zeta_z);
                    if (FULL || o.abs_zeta_z || o.PV) {
                        abs_zz = zeta_z + fv;   // :400
                        if (FULL || o.abs_zeta_z) __stcs(o.abs_zeta_z + m, abs_zz);
                    }
                }
                if (want_div) {
                    double dh;
                    if (CYL) dh = by(dr_times_r(FA), drr) + by(db_dt, drr);   // :408-409
                    else dh = da_dr + db_dt;                             // :266-267
                    if (FULL || o.div_hori) __stcs(o.div_hori + m, dh);
                    if (FULL || o.div) __stcs(o.div + m, dh + dz(FW));
                }
                const double rho_c = (FULL || has_rho) ? sm[s0 + FRHO + ctr_rho] : 1.0;
                if (FULL || o.PV) {
                    const double dthr_dr = dr(FT);
                    const double dthr_dt = dt(FT);
                    const double dthr_dz = dz(FT);
                    const double r_part = zeta_r * dthr_dr;                                  // :463 / :319
                    const double t_part = CYL ? by(zeta_t * dthr_dt, drr) : zeta_t * dthr_dt;    // :464-465 / :320
                    const double z_part = abs_zz * dthr_dz;                                  // :466 / :321
                    __stcs(o.PV + m, (r_part + t_part + z_part) / rho_c);
                }
                if (want_def) {
                    double sh, stc;
                    if (CYL) {
                        sh = dr(FA) - by(a_c, drr) - by(db_dt, drr);    // :490-491
                        stc = dr(FB) - by(b_c, drr) + by(da_dt, drr);   // :492-493
                    } else {
                        sh = da_dr - db_dt;    // :344-345 FD2(u,dx) - FD2(v,dy)
                        stc = db_dr + da_dt;   // :346-347 FD2(v,dx) + FD2(u,dy)
                    }
                    if (FULL || o.shear_deform) __stcs(o.shear_deform + m, sh);
                    if (FULL || o.stretch_deform) __stcs(o.stretch_deform + m, stc);
                    if (FULL || o.filamentation_time || o.strain_region) {
                        const double s2 = sh * sh + stc * stc, z2 = zeta_z * zeta_z;
                        const bool strain = s2 > z2;   // :504-505
                        if ((FULL && CYL) || o.strain_region) o.strain_region[m] = strain ? 1 : 0;
                        if (FULL || o.filamentation_time)   // :501-508, x**-0.5 as 1/sqrt(x): <= 1 ulp
                            __stcs(o.filamentation_time + m, strain ? 2 * (1.0 / sqrt(s2 - z2)) : inf_f64());
                    }
                }
                if (CYL && (FULL || o.radial_PG_force || o.agradient_force)) {
                    const double pg = -dr(FP) / rho_c;   // :343
                    if (FULL || o.radial_PG_force) __stcs(o.radial_PG_force + m, pg);
                    if (FULL || o.agradient_force) {
                        const double cor = fv * b_c;          // :341
                        const double cen = by(b_c * b_c, drr);    // :342
                        __stcs(o.agradient_force + m, cor + cen + pg);   // :344-346
                    }
                }
            };
            if (mm >= s_lo && mm >= 1 && in_dom) {
                if (tile_inside) body(std::true_type{});   // CTA-uniform: no divergence either way
                else body(std::false_type{});
            }
            // plane j - 2 is no longer needed
            if (held > 2) release_oldest();
            slot_p = slot_m, slot_m = slot;
            if (++slot == NS) slot = 0, par ^= 1u;
        }
        while (held > 0) release_oldest();
    }
}

// ------------------------------------------------------------------------------------- host side
struct MapKey {
    const void *ptr;
    int64_t n0, n1, n2;
    int bx, by;
    bool operator==(const MapKey &o) const
    {
        return ptr == o.ptr && n0 == o.n0 && n1 == o.n1 && n2 == o.n2 && bx == o.bx && by == o.by;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const
    {
        size_t h = std::hash<const void *>()(k.ptr);
        for (int64_t v : {k.n0, k.n1, k.n2, (int64_t)k.bx, (int64_t)k.by}) h = h * 1000003u ^ std::hash<int64_t>()(v);
        return h;
    }
};

int plane_map(const double *ptr, int64_t n0, int64_t n1, int64_t n2, int bx, int by, CUtensorMap *out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, n0, n1, n2, bx, by};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return MT_OK;
    }
    auto enc = encode_fn();
    if (!enc) {
        set_error("stencil_tma: cuTensorMapEncodeTiled is not available in this driver");
        return MT_ECUDA;
    }
    cuuint64_t gdim[3] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0};
    cuuint64_t gstr[2] = {(cuuint64_t)n2 * 8, (cuuint64_t)n2 * (cuuint64_t)n1 * 8};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(ptr), gdim, gstr, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("stencil_tma: cuTensorMapEncodeTiled failed (%d) for a (%lld,%lld,%lld) array", (int)r,
                  (long long)n0, (long long)n1, (long long)n2);
        return MT_ECUDA;
    }
    if (cache.size() > 1024) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return MT_OK;
}

// One zero-initialised task counter per launch: a pool of 1024 per device, used round-robin (launches on
// different streams may overlap, so they must not share one); the memset travels on the launch's stream.
int *task_counter(cudaStream_t s)
{
    constexpr int POOL = 1024, MAXDEV = 64;
    static std::mutex mu;
    static int *pool[MAXDEV] = {};
    static unsigned next[MAXDEV] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
    int *c = nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!pool[dev] && cudaMalloc(&pool[dev], POOL * sizeof(int)) != cudaSuccess) return nullptr;
        c = pool[dev] + (next[dev]++ % POOL);
    }
    if (cudaMemsetAsync(c, 0, sizeof(int), s) != cudaSuccess) return nullptr;
    return c;
}

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <int LAYOUT, bool CYL, bool FULL, int TR, int NS, int MB, int MSH>
int launch(const Maps &M, const KP &kp, unsigned grid, cudaStream_t s)
{
    using G = Geo<TR, CYL, NS, MSH>;
    auto kern = k_stencil_tma<LAYOUT, CYL, FULL, TR, NS, MB, MSH>;
    constexpr size_t smem = G::SMEM;
    MT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // several CTAs share an SM: ask for the full shared-memory carve-out
    MT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<grid, G::NTHR, smem, s>>>(M, kp);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

}  // namespace

bool stencil_tma_supported(const mt_grid_t *g)
{
    const bool off = env_int("MT_DIAG_TMA", 1) == 0;   // read per call: A/B runs flip it inside one process
    if (off || !encode_fn()) return false;
    const bool ztr = g->layout == MT_LAYOUT_ZTR;
    const int64_t n0 = ztr ? g->nz : g->nt, n1 = ztr ? g->nt : g->nr, n2 = ztr ? g->nr : g->nz;
    if (!ztr && (g->halo_lo || g->halo_hi)) return false;   // level shards are ZTR arrays
    if (n2 % 2 != 0 || n2 < 4 || n1 < 3 || n0 < 3) return false;   // 16-byte aligned rows
    if (n0 > 2147483000LL || n1 > 2147483000LL || n2 > 2147483000LL) return false;
    return true;
}

template <bool CYL>
int stencil_tma(const mt_grid_t *g, const GridP &gp, const StageBIn &bi, const double *u, const double *v,
                const WindOut &wo, const StageBOut &bo, bool full, cudaStream_t s)
{
    if (!stencil_tma_supported(g)) return MT_ENOSUP;
    // spacings the branch-free quotient cannot take exactly -> the register-marching kernels (IEEE divisions)
    if (!fdiv_divisor_ok(gp.dz) || !fdiv_divisor_ok(gp.dt) || !fdiv_divisor_ok(gp.dr)) return MT_ENOSUP;
    const bool ztr = g->layout == MT_LAYOUT_ZTR;
    const int64_t n0 = ztr ? g->nz : g->nt, n1 = ztr ? g->nt : g->nr, n2 = ztr ? g->nr : g->nz;
    const double *A = u ? u : bi.a, *B = v ? v : bi.b;
    for (const void *ptr : {(const void *)A, (const void *)B, (const void *)bi.w, (const void *)bi.thr,
                            (const void *)bi.p, (const void *)bi.rho})
        if ((uintptr_t)ptr % 16 != 0) return MT_ENOSUP;
    if (!A || !B) return MT_ENOSUP;

#ifndef MT_DIAG_TR
#define MT_DIAG_TR 4
#endif
#ifndef MT_DIAG_MB
#define MT_DIAG_MB 4
#endif
    constexpr int TRv = MT_DIAG_TR;   // rows per tile = consumer warps per CTA (see the sweep at the top of the file)
    KP kp;
    kp.n0 = (int)n0, kp.n1 = (int)n1, kp.n2 = (int)n2;
    // march range: the owned levels of a ZTR level shard; every theta plane of a TRZ array
    kp.m_lo = ztr ? gp.own_lo : 0;
    kp.m_hi = ztr ? gp.own_hi : (int)n0;
    kp.halo_lo = ztr ? gp.halo_lo : 0;
    kp.halo_hi = ztr ? gp.halo_hi : 0;
    // row shift for 256-byte aligned stores (Geo): rows of a ZTR array whose planes are multiples of 32
    // elements start at (gr * n2) mod 32 = a multiple of g = gcd(n2 mod 32, 32); every output must itself
    // start on a 256-byte boundary
    int msh = 0;
    // (narrow rows gain nothing: the wider boxes and the extra column tile cost more than the stores save)
    if (ztr && n2 % 32 != 0 && n2 >= 512 && (n1 * n2) % 32 == 0 && env_int("MT_DIAG_SHIFT", 1) != 0) {
        int g = 32;
        while ((n2 % 32) % g != 0) g >>= 1;
        bool aligned = true;
        for (const void *ptr : {(const void *)bo.radial_PG_force, (const void *)bo.agradient_force, (const void *)bo.zeta_r,
                                (const void *)bo.zeta_t, (const void *)bo.zeta_z, (const void *)bo.abs_zeta_z,
                                (const void *)bo.div_hori, (const void *)bo.div, (const void *)bo.PV,
                                (const void *)bo.shear_deform, (const void *)bo.stretch_deform,
                                (const void *)bo.filamentation_time, (const void *)wo.vr, (const void *)wo.vt,
                                (const void *)wo.wind_speed, (const void *)wo.kinetic_energy,
                                (const void *)wo.coriolis_force, (const void *)wo.centrifugal_force, (const void *)wo.aam})
            aligned = aligned && (uintptr_t)ptr % 256 == 0;
        if (aligned && g >= 8) msh = 32 - g;   // 16 or 24; finer row phases (28, 30) are not instantiated
    }
    kp.tiles_r = (int)((n1 + TRv - 1) / TRv);
    kp.tiles_c = (int)((n2 + msh + TC - 1) / TC);
    kp.dz = gp.dz, kp.dt = gp.dt, kp.dr = gp.dr;
    kp.r = gp.r, kp.sin_t = gp.sin_t, kp.cos_t = gp.cos_t, kp.f = bi.f;
    kp.o = bo, kp.wo = wo;
    const bool any_wind = wo.vr || wo.vt || wo.wind_speed || wo.kinetic_energy || wo.coriolis_force ||
                          wo.centrifugal_force || wo.aam;
    const bool rotate = CYL && u && v;
    kp.flags = (bi.w ? K_W : 0) | (bi.thr ? K_THR : 0) | (bi.rho ? K_RHO : 0) | ((CYL && bi.p) ? K_P : 0) |
               (rotate ? K_ROT : 0) | (any_wind ? K_WIND : 0);
    const int n_march = kp.m_hi - kp.m_lo;
    if (n_march <= 0) return MT_OK;

    // persistent CTAs; the march axis is cut into segments only as far as that balances the SMs: pick the
    // segment count that minimises  rounds x (planes per segment + 2 ramp planes)
    // MT_DIAG_MB CTAs of MT_DIAG_TR consumer warps (96 registers) per SM; where the shifted boxes are wide,
    // shared memory admits one fewer
    const int ctas_per_sm = MT_DIAG_MB;
    const int64_t slots = (int64_t)sm_count() * ctas_per_sm;
    const int64_t ntiles = (int64_t)kp.tiles_r * kp.tiles_c;
    int best_nseg = 1;
    double best_cost = 1e300;
    for (int nseg = 1; nseg <= (n_march + 3) / 4 && nseg <= 512; ++nseg) {
        const int len = (n_march + nseg - 1) / nseg;
        const int64_t tasks = ntiles * ((n_march + len - 1) / len);
        const double rounds = (double)((tasks + slots - 1) / slots);
        const double cost = rounds * (len + 2);
        if (cost < best_cost * 0.999) best_cost = cost, best_nseg = nseg;
    }
    const int nseg_env = env_int("MT_DIAG_NSEG", 0);
    if (nseg_env > 0) best_nseg = nseg_env < n_march ? nseg_env : n_march;
    kp.seg_len = (n_march + best_nseg - 1) / best_nseg;
    kp.nseg = (n_march + kp.seg_len - 1) / kp.seg_len;
    const int64_t ntask = ntiles * kp.nseg;
    const unsigned grid = (unsigned)(ntask < slots ? ntask : slots);

    Maps M;
    const int by = TRv + 2 * HY;
    const int BX = TC + 2 * HX + msh;
    int rc = plane_map(A, n0, n1, n2, BX, by, &M.a);
    if (rc == MT_OK) rc = plane_map(B, n0, n1, n2, BX, by, &M.b);
    M.w = M.a, M.thr = M.a, M.p = M.a, M.rho = M.a;
    if (rc == MT_OK && bi.w) rc = plane_map(bi.w, n0, n1, n2, BX, by, &M.w);
    if (rc == MT_OK && bi.thr) rc = plane_map(bi.thr, n0, n1, n2, BX, by, &M.thr);
    if (rc == MT_OK && CYL && bi.p) rc = plane_map(bi.p, n0, n1, n2, BX, by, &M.p);
    if (rc == MT_OK && bi.rho) rc = plane_map(bi.rho, n0, n1, n2, TC + msh, TRv, &M.rho);
    if (rc != MT_OK) return rc;

    kp.counter = task_counter(s);
    if (!kp.counter) {
        set_error("stencil_tma: cannot allocate the task counter");
        return MT_ECUDA;
    }
    ProfScope prof(CYL ? "k_stencil_tma_cyl" : "k_stencil_tma_car", (mt_stream_t)s);
#ifndef MT_DIAG_NS
#define MT_DIAG_NS 5
#endif
#define DT_L(LAYOUT_, FULL_, MSH_) launch<LAYOUT_, CYL, FULL_, MT_DIAG_TR, MT_DIAG_NS, MT_DIAG_MB, MSH_>(M, kp, grid, s)
#define DT_ZTR(FULL_)                                                                           \
    (msh == 0 ? DT_L(MT_LAYOUT_ZTR, FULL_, 0) : msh == 16 ? DT_L(MT_LAYOUT_ZTR, FULL_, 16) : DT_L(MT_LAYOUT_ZTR, FULL_, 24))
    if (ztr) return full ? DT_ZTR(true) : DT_ZTR(false);
    return full ? DT_L(MT_LAYOUT_TRZ, true, 0) : DT_L(MT_LAYOUT_TRZ, false, 0);
#undef DT_ZTR
#undef DT_L
}

template int stencil_tma<true>(const mt_grid_t *, const GridP &, const StageBIn &, const double *, const double *,
                               const WindOut &, const StageBOut &, bool, cudaStream_t);
template int stencil_tma<false>(const mt_grid_t *, const GridP &, const StageBIn &, const double *, const double *,
                                const WindOut &, const StageBOut &, bool, cudaStream_t);

}  // namespace dg
}  // namespace mt
