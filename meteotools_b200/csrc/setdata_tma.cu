// set_data as ONE warp-specialised TMA kernel (F1, SURVEY 8a rows a1 + a5 + a7 fused):
// vertical interpolation (wrf.interplevel, grids/cylindrical_grid.py:113) -> bilinear gather
// (interp2D_fast_layers, fastcompute.f90:60-109) -> terrain mask (cylindrical_grid.py:171-181).
// The level-interpolated columns never leave shared memory: the [by][bx][nlev] intermediate of the
// two-kernel path (522 MB written + 429 MB re-read on S2) does not exist.
//
// Work decomposition (by SOURCE tile): the targets are bucketed once per plan (host side) by the
// 14 x 4 block of source cells that holds their (x0, y0) corner; a work item = (tile, <= 128 of its
// targets).  A tile needs 15 x 5 source columns -> one 3-D TMA box {16, 5, nz} per field: 128-byte
// rows, dense [nz][5][16] in shared memory.
//
// One persistent CTA per SM = 16 consumer warps + 1 producer warp:
//   producer  takes items from an atomic counter, stages the item's targets (stencil corners as
//             tile-local column ids, weights, terrain index, output row) into a 3-deep ring of
//             shared-memory tables -- software-pipelined ONE ITEM AHEAD, one dependent global load
//             per issued tile, so its latency never meets the consumers -- and streams the tiles
//             coordinate, var 0, var 1, ... of consecutive items through a ring of TMA stages
//             (full/empty mbarriers); the prefetch runs across item boundaries.
//   consumers per item: monotonicity check + bracket search (kp, w2) on the coordinate tile, once
//             for all variables; then per variable  B: level-interpolated columns lev[80][nlev] in
//             shared memory (lanes over columns: conflict-free on the dense TMA layout),
//             C: one warp per target, lanes over the contiguous level index, three lerps in the
//             order of fastcompute.f90:100-102, one 16-byte streaming store per lane.
// Consumers never read global memory; they only wait on mbarriers and store results.
// Arithmetic and its order are identical to k_interplevel_tile + k_gather2d_lf (bit-exact tests).
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "tma.cuh"

namespace {

using mt::nan_f64;

// The TMA unit requires the global address of a box to be 16-byte aligned: tiles start at even
// (float64) / multiple-of-4 (float32) columns and are 14 / 12 cells wide, so that the 16-column box
// holds the tile's width + 1 columns (measured: an odd start column -> cudaErrorIllegalInstruction).
constexpr int BX = 16;                                   // columns per box row (128-byte rows of float64)
static_assert(MT_TMA_TILE_W_F64 + 1 <= BX && MT_TMA_TILE_W_F64 % 2 == 0, "float64 tile width");
static_assert(MT_TMA_TILE_W_F32 + 1 <= BX && MT_TMA_TILE_W_F32 % 4 == 0, "float32 tile width");
constexpr int BY = MT_TMA_TILE_H + 1, NC = BX * BY;     // columns per box: 16 x 5 = 80
constexpr int NG = 6;                                   // level groups of the column phases
// 15 consumer warps (every consumer thread owns one (column, level group)) + one producer warp =
// 16 warps, four per SM sub-partition: 128 registers per thread
constexpr int NCONS = NG * NC, NTHREADS = NCONS + 32;
static_assert(NCONS % 32 == 0, "consumer threads must be whole warps");
constexpr int MAXT = MT_TMA_MAX_TARGETS;                // targets per item
constexpr int NMETA = 3;                                // target-table ring
static_assert(2 * NC <= NTHREADS && MAXT % 32 == 0, "shape");

struct __align__(64) TmaParams {
    CUtensorMap maps[MT_MAX_VARS + 1];   // [0]: vertical coordinate, [v+1]: variable v
    double *out[MT_MAX_VARS];
};

struct __align__(16) Meta {   // one staged target (two 16-byte words; a third word measured slower:
                              // phase C is bound by shared-memory wavefronts as much as by issue slots)
    double fx, fy;
    int tgt;                    // target index (output row = tgt * o_n)
    int terrain;                // terrain level index; bit 30: same four corners as the predecessor
    // the four corner columns as BYTE offsets into `lev` (column id * column pitch, < 65536: 80 columns of
    // <= 98 doubles): corner (y0, x0) in the low half of `b01`, the step to x1 in its high half, the step to
    // y1 in `b23` -- the consumer forms the four addresses with five integer instructions
    uint32_t b01, b23;
};
static_assert(sizeof(Meta) == 32, "Meta is read as two 16-byte words");

using namespace mt::tma;   // smem_u32, mbar_*, load_3d (csrc/tma.cuh)
// store of the 16-byte result pairs (A/B builds with -DMT_TMA_STORE_KIND=n: scratch/build_variant.sh)
#if !defined(MT_TMA_STORE_KIND) || MT_TMA_STORE_KIND == 0
#define MT_TMA_ST(ptr, val) __stcs(ptr, val)
#elif MT_TMA_STORE_KIND == 1
#define MT_TMA_ST(ptr, val) (*(ptr) = (val))
#elif MT_TMA_STORE_KIND == 2
#define MT_TMA_ST(ptr, val) __stcg(ptr, val)
#else
#define MT_TMA_ST(ptr, val) __stwt(ptr, val)
#endif
#ifdef MT_TMA_TIMING
// phase clocks of consumer warp 0 (debug builds only: scratch/tma_timing.py)
__device__ unsigned long long g_tma_timing[8];
#define TMARK(var) const long long var = clock64()
#define TACC(i, a, b) tacc[i] += (unsigned long long)((b) - (a))
#else
#define TMARK(var)
#define TACC(i, a, b)
#endif
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

// The producer warp (shared by both kernels): takes items from the atomic counter, stages the item's targets
// into the ring of shared-memory tables ONE ITEM AHEAD, and streams the tiles coordinate, var 0, var 1, ... of
// consecutive items through the ring of TMA stages (full / empty mbarriers).
template <typename T>
__device__ __forceinline__ void producer_warp(const TmaParams &P, int NT, int nz, int LP, int stages, int stage_bytes,
                                              int org_x, int org_y, const int4 *__restrict__ pix,
                                              const double2 *__restrict__ pw, const int32_t *__restrict__ order,
                                              const int4 *__restrict__ items, int nitems, int n_order,
                                              int *__restrict__ counter, const int32_t *__restrict__ hgt_idx,
                                              unsigned char *smem, Meta *meta, uint64_t *full, uint64_t *empty,
                                              int4 *s_item4, int *s_itemid, int lane)
{
    // =============================== producer warp ===============================
    int q_slot = 0;
    uint32_t q_par = 1;  // a fresh mbarrier passes a wait on parity 1: the ring starts empty
    int n = 0;
    // staging pipeline of the NEXT item (registers)
    int nid = 0;
    int4 nit = make_int4(0, 0, 0, 0);
    constexpr int TS = MAXT / 32;
    const uint32_t cpitch_p = (uint32_t)LP * 8u;   // bytes per level-interpolated column
    int tgt[TS];
    int4 six[TS];
    double2 sw[TS];
    int ster[TS];
    auto stage_phase = [&](int ph, int slot_meta) {
        switch (ph) {
            case 0:
                if (lane == 0) nid = atomicAdd(counter, 1);
                break;
            case 1:
                nid = __shfl_sync(0xffffffffu, nid, 0);
                if (nid < nitems) nit = items[nid];
                else nit = make_int4(0, 0, 0, 0);
                if (nid < nitems) {
                    // `items` comes from the caller.  A tile origin that is not 16-byte aligned makes the
                    // TMA unit raise cudaErrorIllegalInstruction (sticky: the context is lost), a bad
                    // start / count reads beyond `order`: such an item is made harmless here -- origin
                    // rounded down, no targets -- and reported through the status word counter[1]
                    // (mt_setdata_tma_status -> MT_EINVAL).
                    constexpr int ALIGN = 16 / (int)sizeof(T);
                    const int rel = nit.x - org_x;
                    const int mis = ((rel % ALIGN) + ALIGN) % ALIGN;
                    const bool bad = mis != 0 || nit.w < 0 || nit.w > MAXT || nit.z < 0 || nit.z > n_order - nit.w;
                    if (bad) {
                        if (lane == 0) atomicOr(counter + 1, 1);
                        nit.x -= mis, nit.w = 0, nit.z = 0;
                    }
                }
                break;
            case 2:
#pragma unroll
                for (int k = 0; k < TS; ++k) {
                    const int j = lane + 32 * k;
                    tgt[k] = j < nit.w ? order[nit.z + j] : -1;
                }
                break;
            case 3:
#pragma unroll
                for (int k = 0; k < TS; ++k)
                    if (tgt[k] >= 0) {
                        six[k] = pix[tgt[k]];
                        sw[k] = pw[tgt[k]];
                        ster[k] = hgt_idx ? hgt_idx[tgt[k]] : 0;
                    }
                break;
            case 4: {
                Meta *mt_ = meta + slot_meta * MAXT;
#pragma unroll
                for (int k = 0; k < TS; ++k) {
                    const unsigned act = __ballot_sync(0xffffffffu, tgt[k] >= 0);
                    if (tgt[k] >= 0) {
                        Meta m;
                        m.fx = sw[k].x, m.fy = sw[k].y, m.tgt = tgt[k];
                        const uint32_t r0 = (uint32_t)(six[k].z - nit.y) * BX, r1 = (uint32_t)(six[k].w - nit.y) * BX;
                        const uint32_t c0 = (uint32_t)(six[k].x - nit.x), c1 = (uint32_t)(six[k].y - nit.x);
                        // x1 >= x0 and y1 >= y0 (equal on the clamped edge of the source grid)
                        m.b01 = (r0 + c0) * cpitch_p | (((c1 - c0) * cpitch_p) << 16);
                        m.b23 = (r1 - r0) * cpitch_p;
                        // the predecessor of an odd position is the previous lane's target
                        const uint32_t p01 = __shfl_up_sync(act, m.b01, 1), p23 = __shfl_up_sync(act, m.b23, 1);
                        const bool same = (lane & 1) && p01 == m.b01 && p23 == m.b23;
                        m.terrain = ster[k] | (same ? (1 << 30) : 0);
                        mt_[lane + 32 * k] = m;
                    }
                }
                __syncwarp();
                break;
            }
            default: break;
        }
    };
    for (int ph = 0; ph < 5; ++ph) stage_phase(ph, 0);  // first item: synchronously
    while (true) {
        const int id = nid;
        const int4 it = nit;
        if (id >= nitems) {  // end marker: complete the phase of the next stage without data
            if (lane == 0) {
                mbar_wait(empty + q_slot, q_par);
                s_itemid[n & 3] = -1;
                mbar_arrive(full + q_slot);
            }
            break;
        }
        int ph = 0;
        for (int t = 0; t < NT; ++t) {
            if (lane == 0) {
                mbar_wait(empty + q_slot, q_par);
                if (t == 0) {
                    s_itemid[n & 3] = id;
                    s_item4[n & 3] = it;
                }
                mbar_expect_tx(full + q_slot, (uint32_t)(NC * nz * (int)sizeof(T)));
                load_3d(smem + (size_t)q_slot * stage_bytes, &P.maps[t], full + q_slot, it.x - org_x,
                            it.y - org_y, 0);
            }
            if (++q_slot == stages) q_slot = 0, q_par ^= 1u;
            __syncwarp();
            if (ph < 5) stage_phase(ph++, (n + 1) % NMETA);
        }
        while (ph < 5) stage_phase(ph++, (n + 1) % NMETA);
        ++n;
    }
}

// PP = level pairs per column thread (>= ceil(nlev / 2 / NG)): the brackets of a thread's own
// (column, level pair) set live in registers for all variables of the item.
// EXACT: nlev == 2 * NG * PP, every thread owns exactly PP level pairs (no guards in the column
// phases; with PP = 5 the 60 levels are one step of the 30-lane level loop of phase C).
template <typename T, bool INCLUSIVE, bool ROUND, int PP, bool EXACT>
__global__ void __launch_bounds__(NTHREADS, 1)
k_setdata_tma(const __grid_constant__ TmaParams P, int nvar, int nz, int nlev, int LP, int stages,
              int stage_bytes, int org_x, int org_y, const double *__restrict__ levels, int decreasing,
              const int4 *__restrict__ pix, const double2 *__restrict__ pw,
              const int32_t *__restrict__ order, const int4 *__restrict__ items, int nitems, int n_order,
              int *__restrict__ counter, const int32_t *__restrict__ hgt_idx, int64_t o_n, int off_lev,
              int lev_bytes, int off_meta, int off_misc)
{
    extern __shared__ unsigned char smem_dyn[];
    // TMA destinations must be 128-byte aligned
    unsigned char *smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    Meta *meta = reinterpret_cast<Meta *>(smem + off_meta);                   // [NMETA][MAXT]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + off_misc);           // [8]
    uint64_t *empty = full + 8;                                               // [8]
    int4 *s_item4 = reinterpret_cast<int4 *>(empty + 8);                      // [4]
    int *s_itemid = reinterpret_cast<int *>(s_item4 + 4);                     // [4]
    double *lev_s = reinterpret_cast<double *>(s_itemid + 4);                 // [nlev + 1]
    unsigned char *mono_s = reinterpret_cast<unsigned char *>(lev_s + nlev + 1);  // [2][NC]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = nvar + 1;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_fence_init();
    }
    // an odd level count is handled as pairs too: the last level is its own partner (the second
    // half of that pair is computed and never stored)
    for (int l = tid; l <= nlev; l += NTHREADS) lev_s[l] = levels[min(l, nlev - 1)];
    if (tid < 2 * NC) mono_s[tid] = 1;
    __syncthreads();

    if (warp == NCONS / 32) {
        producer_warp<T>(P, NT, nz, LP, stages, stage_bytes, org_x, org_y, pix, pw, order, items, nitems, n_order,
                         counter, hgt_idx, smem, meta, full, empty, s_item4, s_itemid, lane);
        return;
    }

    // ================================= consumers =================================
    const int im = decreasing ? 0 : 1, ip = decreasing ? 1 : 0;
    const double sgn = decreasing ? -1.0 : 1.0;
    const int dhi = decreasing ? -NC : NC;      // element offset from the "lo" sample to the "hi" sample
    const int klo_missing = decreasing ? 1 : 0; // any pair of valid rows: a missing bracket has w2 = NaN
    const int stage_elems = stage_bytes / (int)sizeof(T);
    const T *raw = reinterpret_cast<const T *>(smem);
    // (column, level-group) ownership of the column phases: thread -> column tid % 80 and the ppn
    // consecutive level pairs from (tid / 80) * ppn
    const int c_own = tid % NC, g_own = tid / NC;
    const bool col_thread = tid < NG * NC;
    const int npair = (nlev + 1) >> 1, ppn = EXACT ? PP : (npair + NG - 1) / NG, lp_own = g_own * ppn;
    const bool wide = ((nlev | (int)o_n) & 1) == 0;   // even rows: 16-byte stores; else two 8-byte stores
    const uint32_t o_pitch_b = (uint32_t)o_n * 8u;   // output row pitch in bytes (the host checks that it fits)
    // this thread's brackets (row offsets of the "lo" samples of both levels of a pair, weights)
    int kk[PP];
    double wx[PP], wy[PP];
    int c_slot = 0;
    uint32_t c_par = 0;
    int lev_sel = 0;   // `lev` is double-buffered: B of variable v + 1 may overlap C of variable v
#ifdef MT_TMA_TIMING
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int n = 0;; ++n) {
        TMARK(ta);
        mbar_wait(full + c_slot, c_par);
        TMARK(tb0);
        TACC(0, ta, tb0);   // wait for the coordinate tile
        if (s_itemid[n & 3] < 0) break;
        const int cnt = s_item4[n & 3].w;
        const T *vt = raw + (size_t)c_slot * stage_elems + c_own;
        unsigned char *mono = mono_s + (n & 1) * NC;
        // monotone (in the chosen direction) per column: six threads per column, 1/6 of the pairs each
        if (col_thread) {
            const int per = (nz - 1 + NG - 1) / NG;
            const int k0 = g_own * per, k1 = min(k0 + per, nz - 1);
            bool ok = true;
            for (int k = k0; k < k1; ++k) ok = ok && (sgn * (double)vt[(k + 1) * NC] >= sgn * (double)vt[k * NC]);
            if (!ok) mono[c_own] = 0;
        }
        cons_sync();
        // brackets of this thread's (level pair, column) set, shared by all variables: klo = row of
        // the "lo" sample (kp - im of DINTERP3DZ), w2 = its weight; no bracket -> w2 = NaN, so that
        // w1*f + w2*f is NaN without a branch in the per-variable loop.
        if (col_thread) {
            const bool is_mono = mono[c_own] != 0;
            // K = first row with sgn * v[K] > sgn * want (upper bound).  The thread's levels are
            // consecutive: when they ascend (in the direction of the coordinate) K only moves
            // forward, so after the first binary search the rows are walked, not searched.
            int K = 0;
            double ww_prev = 0.0;
            bool have_prev = false;
            // Two passes: (1) the rows of all the thread's levels (a dependent chain of shared-memory
            // loads: search, then walk), (2) their weights -- 2 * PP independent divisions that
            // overlap instead of one division at the end of every walk.
            auto find_row = [&](double want, int &klo) {
                int kp = -1;
                if (is_mono) {
                    const double ww = sgn * want;
                    if (have_prev && ww >= ww_prev) {
                        while (K < nz && sgn * (double)vt[K * NC] <= ww) ++K;
                    } else {
                        int lo = 0, hi = nz;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (sgn * (double)vt[mid * NC] > ww) hi = mid;
                            else lo = mid + 1;
                        }
                        K = lo;
                    }
                    ww_prev = ww, have_prev = true;
                    if (K >= 1 && K <= nz - 1) {
                        if (INCLUSIVE) kp = K;
                        else if (sgn * (double)vt[(K - 1) * NC] < ww) kp = K;
                    } else if (INCLUSIVE && K == nz && sgn * (double)vt[(nz - 1) * NC] == ww) {
                        kp = nz - 1;
                    }
                } else {
                    for (int q = nz - 1; q >= 1; --q) {  // literal DINTERP3DZ scan from the top
                        const double vlo = (double)vt[(q - im) * NC], vhi = (double)vt[(q - ip) * NC];
                        const bool hit = INCLUSIVE ? (vlo <= want && vhi >= want) : (vlo < want && vhi > want);
                        if (hit) {
                            kp = q;
                            break;
                        }
                    }
                }
                klo = kp >= 0 ? kp - im : klo_missing;
                return kp >= 0;
            };
            unsigned found = 0;
#pragma unroll
            for (int i = 0; i < PP; ++i) {
                const int lp = lp_own + i;
                int k0 = klo_missing, k1 = klo_missing;
                if (EXACT || (i < ppn && lp < npair)) {
                    if (find_row(lev_s[2 * lp], k0)) found |= 1u << (2 * i);
                    if (find_row(lev_s[2 * lp + 1], k1)) found |= 2u << (2 * i);
                }
                kk[i] = (k0 * NC) | ((k1 * NC) << 16);   // nz <= 254: k * 80 < 65536
            }
            // missing bracket: the weight is NaN (the quotient of two valid rows is computed and dropped)
#pragma unroll
            for (int i = 0; i < PP; ++i) {
                const int lp = lp_own + i;
                const bool act = EXACT || (i < ppn && lp < npair);
                const double want0 = act ? lev_s[2 * lp] : 0.0, want1 = act ? lev_s[2 * lp + 1] : 0.0;
                const int ka = kk[i] & 0xffff, kb = kk[i] >> 16;
                const double a0 = (double)vt[ka], a1 = (double)vt[ka + dhi];
                const double b0 = (double)vt[kb], b1 = (double)vt[kb + dhi];
                const double q0 = (want0 - a0) / (a1 - a0), q1 = (want1 - b0) / (b1 - b0);
                wx[i] = (found >> (2 * i)) & 1u ? q0 : nan_f64();
                wy[i] = (found >> (2 * i + 1)) & 1u ? q1 : nan_f64();
            }
        }
        if (tid < NC) mono_s[((n + 1) & 1) * NC + tid] = 1;
        cons_sync();
        if (tid == 0) mbar_arrive(empty + c_slot);
        if (++c_slot == stages) c_slot = 0, c_par ^= 1u;
        TMARK(tb1);
        TACC(1, tb0, tb1);   // monotonicity + brackets (incl. their two barriers)

        const Meta *mt_ = meta + (n % NMETA) * MAXT;
        for (int v = 0; v < nvar; ++v) {
            unsigned char *levb_w = smem + off_lev + lev_sel * lev_bytes;   // [NC][LP] doubles
            TMARK(t0);
            mbar_wait(full + c_slot, c_par);
            TMARK(t1);
            TACC(2, t0, t1);   // wait for the variable's tile
            // B: level-interpolated columns, one level PAIR per step (branch-free, brackets from
            // registers, conflict-free 16-byte stores)
            if (col_thread) {
                const T *fc = raw + (size_t)c_slot * stage_elems + c_own;
                double *lv = reinterpret_cast<double *>(levb_w) + c_own * LP;
                // all the samples first, then the lerps and stores: a store to `lev` may alias the tile as far as
                // the compiler knows, so per-pair code is a chain load -> lerp -> store -> next load (5 round trips)
                constexpr int BB = PP <= 5 ? PP : 1;   // pairs per batch (4 doubles each in registers; PP = 8 has none to spare)
#pragma unroll
                for (int i0 = 0; i0 < PP; i0 += BB) {
                    double f[BB][4];
#pragma unroll
                    for (int i = i0; i < i0 + BB && i < PP; ++i) {
                        const int ka = kk[i] & 0xffff, kb = kk[i] >> 16;   // valid rows also for unused pairs
                        f[i - i0][0] = (double)fc[ka], f[i - i0][1] = (double)fc[ka + dhi];
                        f[i - i0][2] = (double)fc[kb], f[i - i0][3] = (double)fc[kb + dhi];
                    }
#pragma unroll
                    for (int i = i0; i < i0 + BB && i < PP; ++i) {
                        const int lp = lp_own + i;
                        if (EXACT || (i < ppn && lp < npair)) {
                            double2 r;
                            r.x = (1.0 - wx[i]) * f[i - i0][0] + wx[i] * f[i - i0][1];
                            r.y = (1.0 - wy[i]) * f[i - i0][2] + wy[i] * f[i - i0][3];
                            if (ROUND) r.x = (double)(float)r.x, r.y = (double)(float)r.y;
                            *reinterpret_cast<double2 *>(lv + 2 * lp) = r;
                        }
                    }
                }
            }
            TMARK(t2);
            TACC(3, t1, t2);   // B
            cons_sync();   // the only block barrier per variable: `lev` complete, tile consumed
            TMARK(t3);
            TACC(4, t2, t3);   // barrier
            if (tid == 0) mbar_arrive(empty + c_slot);  // the ring slot may be refilled
            if (++c_slot == stages) c_slot = 0, c_par ^= 1u;
            // C: bilinear, one warp per target, two consecutive targets per step; the item's targets
            // are sorted by source cell, so a pair usually shares its four corner columns: one set
            // of shared-memory loads then serves both
            // this lane's level pair of output row 0
            unsigned char *out_b = reinterpret_cast<unsigned char *>(P.out[v]) + lane * 16;
            const uint32_t lane_a = smem_u32(levb_w) + lane * 16;   // this lane's level pair, as a shared address
            auto lds2 = [](uint32_t addr) {
                double2 r;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
                return r;
            };
            for (int j = 2 * warp; j < cnt; j += 2 * (NCONS / 32)) {
                const bool two = j + 1 < cnt;
                const Meta m0 = mt_[j];
                const Meta m1 = mt_[two ? j + 1 : j];
                // row address = base + tgt * pitch in bytes: ONE 32 x 32 + 64 multiply-add per target
                double *o0l = reinterpret_cast<double *>(out_b + (uint64_t)(uint32_t)m0.tgt * o_pitch_b);
                double *o1l = reinterpret_cast<double *>(out_b + (uint64_t)(uint32_t)m1.tgt * o_pitch_b);
                const bool same = !two || (m1.terrain >> 30) != 0;
                const int ter0 = m0.terrain & 0x3fffffff, ter1 = m1.terrain & 0x3fffffff;
                const bool masked = (ter0 | ter1) > 0;                                   // warp-uniform
                const uint32_t pa0 = lane_a + (m0.b01 & 0xffffu), pb0 = pa0 + (m0.b01 >> 16);
                const uint32_t pc0 = pa0 + m0.b23, pd0 = pc0 + (m0.b01 >> 16);
                const uint32_t pa1 = lane_a + (m1.b01 & 0xffffu), pb1 = pa1 + (m1.b01 >> 16);
                const uint32_t pc1 = pa1 + m1.b23, pd1 = pc1 + (m1.b01 >> 16);
                constexpr bool ONE_STEP = EXACT && 2 * NG * PP <= 64;
                for (int l = lane * 2, lo8 = 0; l < nlev; l += 64, lo8 += 512) {
                    const double2 a0 = lds2(pa0 + lo8), b0 = lds2(pb0 + lo8);
                    const double2 c0 = lds2(pc0 + lo8), d0 = lds2(pd0 + lo8);
                    // both targets of the step from their corner values (the SAME registers when the pair
                    // shares its source cell: two instantiations of the body instead of eight register copies)
                    auto finish = [&](const double2 &a1, const double2 &b1, const double2 &c1, const double2 &d1) {
                        double2 r0, r1;
                        {   // fastcompute.f90:100-102
                            const double t0x = (b0.x - a0.x) * m0.fx + a0.x, t1x = (d0.x - c0.x) * m0.fx + c0.x;
                            const double t0y = (b0.y - a0.y) * m0.fx + a0.y, t1y = (d0.y - c0.y) * m0.fx + c0.y;
                            r0.x = (t1x - t0x) * m0.fy + t0x;
                            r0.y = (t1y - t0y) * m0.fy + t0y;
                        }
                        {
                            const double t0x = (b1.x - a1.x) * m1.fx + a1.x, t1x = (d1.x - c1.x) * m1.fx + c1.x;
                            const double t0y = (b1.y - a1.y) * m1.fx + a1.y, t1y = (d1.y - c1.y) * m1.fx + c1.y;
                            r1.x = (t1x - t0x) * m1.fy + t0x;
                            r1.y = (t1y - t0y) * m1.fy + t0y;
                        }
                        if (masked) {   // terrain: level index below hgt_idx -> NaN (cylindrical_grid.py:171-181)
                            if (ter0 > l) r0.x = nan_f64();
                            if (ter0 > l + 1) r0.y = nan_f64();
                            if (ter1 > l) r1.x = nan_f64();
                            if (ter1 > l + 1) r1.y = nan_f64();
                        }
                        double *o0 = o0l + (lo8 >> 3), *o1 = o1l + (lo8 >> 3);   // = row + l
                        if (EXACT || wide) {
                            MT_TMA_ST(reinterpret_cast<double2 *>(o0), r0);
                            if (two) MT_TMA_ST(reinterpret_cast<double2 *>(o1), r1);
                        } else {   // odd level count / odd row pitch: rows are only 8-byte aligned
                            __stcs(o0, r0.x);
                            if (l + 1 < nlev) __stcs(o0 + 1, r0.y);
                            if (two) {
                                __stcs(o1, r1.x);
                                if (l + 1 < nlev) __stcs(o1 + 1, r1.y);
                            }
                        }
                    };
                    if (same) {
                        finish(a0, b0, c0, d0);
                    } else {
                        const double2 a1 = lds2(pa1 + lo8), b1 = lds2(pb1 + lo8);
                        const double2 c1 = lds2(pc1 + lo8), d1 = lds2(pd1 + lo8);
                        finish(a1, b1, c1, d1);
                    }
                    if (ONE_STEP) break;
                }
            }
            TMARK(t4);
            TACC(5, t3, t4);   // C
            lev_sel ^= 1;   // no barrier here: the next variable's B writes the other `lev` buffer
        }
    }
#ifdef MT_TMA_TIMING
    if (lane == 0 && (warp == 0 || warp == 7 || warp == 14))
        for (int i = 0; i < 6; ++i) atomicAdd(&g_tma_timing[i], tacc[i]);
#endif
}

// targets whose coordinates are NaN / the legacy sentinel: no stencil, just the fill value
struct OutPtrs {
    double *out[MT_MAX_VARS];
};
__global__ void k_fill_flagged2(OutPtrs ptrs, int nvar, const int4 *__restrict__ pix, int64_t n, int nlev,
                                int64_t o_n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int flag = pix[i].x;
        if (flag >= 0) continue;
        const double fill = flag == -1 ? MT_SENTINEL : nan_f64();
        for (int v = 0; v < nvar; ++v)
            for (int l = 0; l < nlev; ++l) ptrs.out[v][i * o_n + l] = fill;
    }
}

// ---- host side: tensor maps (cached per source buffer) ---------------------------------------
struct MapKey {
    const void *ptr;
    int64_t nz, ny, nx, pitch;
    int dtype, by;
    bool operator==(const MapKey &o) const
    {
        return ptr == o.ptr && nz == o.nz && ny == o.ny && nx == o.nx && pitch == o.pitch && dtype == o.dtype &&
               by == o.by;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const
    {
        size_t h = std::hash<const void *>()(k.ptr);
        for (int64_t v : {k.nz, k.ny, k.nx, k.pitch, (int64_t)k.dtype, (int64_t)k.by}) h = h * 1000003u ^ std::hash<int64_t>()(v);
        return h;
    }
};

int tensor_map_for(const void *ptr, int dtype, int64_t nz, int64_t ny, int64_t nx, int64_t pitch, int by,
                   CUtensorMap *out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, nz, ny, nx, pitch, dtype, by};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return MT_OK;
    }
    auto enc = encode_fn();
    if (!enc) {
        mt::set_error("mt_setdata_tma: cuTensorMapEncodeTiled is not available in this driver");
        return MT_ECUDA;
    }
    const cuuint64_t eb = dtype == MT_F64 ? 8 : 4;
    cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch * eb, (cuuint64_t)pitch * (cuuint64_t)ny * eb};
    cuuint32_t box[3] = {(cuuint32_t)BX, (cuuint32_t)by, (cuuint32_t)nz};
    cuuint32_t es[3] = {1, 1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, dtype == MT_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           const_cast<void *>(ptr), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        mt::set_error("mt_setdata_tma: cuTensorMapEncodeTiled failed (%d) for a (%lld,%lld,%lld) array, pitch %lld",
                      (int)r, (long long)nz, (long long)ny, (long long)nx, (long long)pitch);
        return MT_ECUDA;
    }
    if (cache.size() > 512) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return MT_OK;
}

struct Layout {
    int stages, stage_bytes, LP, off_lev, lev_bytes, off_meta, off_misc;
    size_t total;
};

bool make_layout(int dtype, int64_t nz, int64_t nlev, Layout *L)
{
    static const int want = [] {   // MT_TMA_STAGES=2..4: ring depth (experiments); default: what fits
        const char *e = getenv("MT_TMA_STAGES");
        const int v = e ? atoi(e) : 4;
        return v < 2 ? 2 : (v > 4 ? 4 : v);
    }();
    const size_t eb = dtype == MT_F64 ? 8 : 4;
    const size_t stage = ((size_t)NC * nz * eb + 127) & ~size_t(127);
    const int64_t ne = (nlev + 1) & ~int64_t(1);               // levels rounded up to whole pairs
    const int LP = (int)((ne % 4 == 2) ? ne : ne + 2);         // even; LP = 2 (mod 4): 2-way conflicts at worst
    const size_t lev_bytes = ((size_t)NC * LP * 8 + 127) & ~size_t(127);
    for (int stages = want; stages >= 2; --stages) {
        size_t off = stage * stages;
        const size_t off_lev = off;
        off += 2 * lev_bytes;   // double-buffered level-interpolated columns
        const size_t off_meta = off;
        off += (size_t)NMETA * MAXT * sizeof(Meta);
        const size_t off_misc = off;
        off += 16 * 8 + 4 * 16 + 4 * 4 + (size_t)(nlev + 1) * 8 + 2 * NC;
        const size_t total = off + 128;  // slack for the 128-byte alignment of the base
        if (total <= 227 * 1024) {
            L->stages = stages, L->stage_bytes = (int)stage, L->LP = LP;
            L->off_lev = (int)off_lev, L->lev_bytes = (int)lev_bytes;
            L->off_meta = (int)off_meta, L->off_misc = (int)off_misc;
            L->total = total;
            return true;
        }
    }
    return false;
}

constexpr int PP_SMALL = 5, PP_LARGE = 8;   // level pairs per column thread: nlev <= 60 / <= 96

int supported_impl(int dtype, int64_t nz, int64_t nlev, int64_t row_pitch)
{
    if (dtype != MT_F64 && dtype != MT_F32) return 0;
    const int64_t eb = dtype == MT_F64 ? 8 : 4;
    if (nz < 2 || nz > 254 || nlev < 1 || nlev > 2 * NG * PP_LARGE || (row_pitch * eb) % 16 != 0)
        return 0;
    Layout L;
    if (!make_layout(dtype, nz, nlev, &L)) return 0;
    return encode_fn() != nullptr ? 1 : 0;
}

int setdata_tma_impl(const void *const *fields, int nvar, const void *vert, int dtype, int64_t nz, int64_t ny,
                     int64_t nx, int64_t row_pitch, int64_t org_y, int64_t org_x, const double *levels,
                     int64_t nlev, int direction, const int32_t *plan_ix, const double *plan_w, int64_t n,
                     const int32_t *order, const int32_t *items, int64_t nitems, int64_t nflagged,
                     const int32_t *hgt_idx, double *const *out, int64_t o_n, int flags, int32_t *counter,
                     mt_stream_t stream)
{
    MT_REQUIRE(fields && vert && levels && plan_ix && plan_w && order && items && out && counter,
               "mt_setdata_tma: null pointer");
    MT_REQUIRE(nvar >= 1 && nvar <= MT_MAX_VARS, "mt_setdata_tma: nvar must be 1..%d", MT_MAX_VARS);
    MT_REQUIRE(dtype == MT_F64 || dtype == MT_F32, "mt_setdata_tma: bad dtype");
    MT_REQUIRE(direction == 0 || direction == 1, "mt_setdata_tma: direction must be 0 or 1 (decided by the caller "
                                                 "from column (0,0) of the full coordinate array)");
    const int64_t eb = dtype == MT_F64 ? 8 : 4;
    MT_REQUIRE(nz >= 2 && nz <= 254 && nlev >= 1 && nlev <= 2 * NG * PP_LARGE && o_n >= nlev &&
                   o_n < (1LL << 28) && n < 2147483647LL,
               "mt_setdata_tma: need 2 <= nz <= 254 and 1..%d target levels", 2 * NG * PP_LARGE);
    const bool wide_rows = nlev % 2 == 0 && o_n % 2 == 0;   // 16-byte stores need 16-byte aligned rows
    MT_REQUIRE(row_pitch >= nx && (row_pitch * eb) % 16 == 0 && ((uintptr_t)vert % 16) == 0,
               "mt_setdata_tma: source rows must be 16-byte aligned (row pitch %lld elements)", (long long)row_pitch);
    Layout L;
    MT_REQUIRE(make_layout(dtype, nz, nlev, &L),
               "mt_setdata_tma: tile does not fit in shared memory (nz=%lld, nlev=%lld)", (long long)nz,
               (long long)nlev);
    TmaParams P;
    int rc = tensor_map_for(vert, dtype, nz, ny, nx, row_pitch, BY, &P.maps[0]);
    if (rc != MT_OK) return rc;
    OutPtrs op;
    for (int v = 0; v < nvar; ++v) {
        MT_REQUIRE(fields[v] && out[v] && ((uintptr_t)out[v] % (wide_rows ? 16 : 8) == 0) &&
                       ((uintptr_t)fields[v] % 16 == 0),
                   "mt_setdata_tma: field %d is null or not aligned (sources 16 bytes, outputs %d)", v,
                   wide_rows ? 16 : 8);
        rc = tensor_map_for(fields[v], dtype, nz, ny, nx, row_pitch, BY, &P.maps[v + 1]);
        if (rc != MT_OK) return rc;
        P.out[v] = out[v];
        op.out[v] = out[v];
    }
    cudaStream_t s = (cudaStream_t)stream;
    const auto pix = reinterpret_cast<const int4 *>(plan_ix);
    const auto pw = reinterpret_cast<const double2 *>(plan_w);
    if (nflagged > 0) {
        mt::ProfScope prof("k_fill_flagged", stream);
        k_fill_flagged2<<<mt::grid_for(n, 256, 4), 256, 0, s>>>(op, nvar, pix, n, (int)nlev, o_n);
        MT_LAUNCH_CHECK();
    }
    if (nitems == 0) return MT_OK;
    MT_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), s));
    int64_t grid64 = mt::sm_count();   // persistent: one CTA per SM
    if (grid64 > nitems) grid64 = nitems;
    const unsigned grid = (unsigned)grid64;
    const bool rf = (flags & MT_IL_ROUND_F32) != 0, inc = (flags & MT_IL_INCLUSIVE) != 0;
    mt::ProfScope prof("k_setdata_tma", stream);
    const bool small = nlev <= 2 * NG * PP_SMALL, exact_small = nlev == 2 * NG * PP_SMALL && wide_rows;
#define ST_LAUNCH(T, INC, RND, PPV, EX)                                                                              \
    do {                                                                                                           \
        MT_CUDA(cudaFuncSetAttribute(k_setdata_tma<T, INC, RND, PPV, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)L.total));                                                               \
        k_setdata_tma<T, INC, RND, PPV, EX><<<grid, NTHREADS, L.total, s>>>(                                           \
            P, nvar, (int)nz, (int)nlev, L.LP, L.stages, L.stage_bytes, (int)org_x, (int)org_y, levels, direction, \
            pix, pw, order, reinterpret_cast<const int4 *>(items), (int)nitems, (int)(n - nflagged), counter,  \
            hgt_idx, o_n, L.off_lev,                                                                               \
            L.lev_bytes, L.off_meta, L.off_misc);                                                                  \
    } while (0)
#define ST_DISPATCH2(T, INC, RND)                      \
    do {                                               \
        if (exact_small) ST_LAUNCH(T, INC, RND, PP_SMALL, true);   \
        else if (small) ST_LAUNCH(T, INC, RND, PP_SMALL, false);    \
        else ST_LAUNCH(T, INC, RND, PP_LARGE, false);               \
    } while (0)
#define ST_DISPATCH(T)                              \
    do {                                            \
        if (inc && rf) ST_DISPATCH2(T, true, true); \
        else if (inc) ST_DISPATCH2(T, true, false); \
        else if (rf) ST_DISPATCH2(T, false, true);  \
        else ST_DISPATCH2(T, false, false);         \
    } while (0)
    if (dtype == MT_F64) ST_DISPATCH(double);
    else ST_DISPATCH(float);
#undef ST_DISPATCH
#undef ST_DISPATCH2
#undef ST_LAUNCH
    MT_LAUNCH_CHECK();
    return MT_OK;
}

}  // namespace

extern "C" {

int mt_setdata_tma_supported(int dtype, int64_t nz, int64_t nlev, int64_t row_pitch)
{
    return supported_impl(dtype, nz, nlev, row_pitch);
}

int mt_setdata_tma(const void *const *fields, int nvar, const void *vert, int dtype, int64_t nz, int64_t ny,
                   int64_t nx, int64_t row_pitch, int64_t org_y, int64_t org_x, const double *levels,
                   int64_t nlev, int direction, const int32_t *plan_ix, const double *plan_w, int64_t n,
                   const int32_t *order, const int32_t *items, int64_t nitems, int64_t nflagged,
                   const int32_t *hgt_idx, double *const *out, int64_t o_n, int flags, int32_t *counter,
                   mt_stream_t stream)
{
    return setdata_tma_impl(fields, nvar, vert, dtype, nz, ny, nx, row_pitch, org_y, org_x, levels, nlev, direction,
                            plan_ix, plan_w, n, order, items, nitems, nflagged, hgt_idx, out, o_n, flags, counter,
                            stream);
}

int mt_setdata_tma_status(int32_t *counter, mt_stream_t stream)
{
    MT_REQUIRE(counter, "mt_setdata_tma_status: null counter");
    int32_t st = 0;
    MT_CUDA(cudaMemcpyAsync(&st, counter + 1, sizeof(st), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (st == 0) return MT_OK;
    MT_CUDA(cudaMemsetAsync(counter + 1, 0, sizeof(st), (cudaStream_t)stream));
    mt::set_error("mt_setdata_tma: `items` held entries that break the contract (tile_x - org_x not a multiple of "
                  "16 bytes, count outside 0..%d, or start + count beyond the `order` array); their targets were "
                  "NOT written", MAXT);
    return MT_EINVAL;
}

#ifdef MT_TMA_TIMING
int mt_setdata_tma_timing(unsigned long long *out8, int reset)
{
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaDeviceSynchronize() != cudaSuccess) return MT_ECUDA;
    if (out8 && cudaMemcpyFromSymbol(out8, g_tma_timing, sizeof(z)) != cudaSuccess) return MT_ECUDA;
    if (reset && cudaMemcpyToSymbol(g_tma_timing, z, sizeof(z)) != cudaSuccess) return MT_ECUDA;
    return MT_OK;
}
#endif

}  // extern "C"
