// Reductions and smoothers next to the hot path:
//   * range averages of calc/averages.py:13-114 (calc_Zaverage, calc_Raverage, calc_RZaverage) and
//     the grid methods built on them (cylindrical_grid.py:513-546, cartesian_grid.py:370-391);
//   * azimuthal mean / anomaly / axisymmetry (cylindrical_grid.py:187-235);
//   * integrated kinetic energy (cylindrical_grid.py:554-600);
//   * the 1-c-1 smoothers of cartesian_grid.py:142-229 (Jacobi passes over a padded copy whose
//     ghost cells stay frozen at the INITIAL edge values).
// Sums are NaN-skipping exactly where the reference uses nansum / nanmean, and the sums along one
// axis are added in NumPy's order -- sequentially along a strided axis, in pairwise blocks along
// the contiguous one -- so they are bit-identical to the reference's.  The full-grid sums (IKE,
// the last step of calc_RZaverage) use a block tree and stay tolerance-class (1e-10 relative,
// north star); the smoothers are pure +,*,/ in the reference's order: bit-exact.
#include <cstdlib>

#include "common.cuh"
#include "fd.cuh"

namespace {

using mt::nan_f64;

// ---- weighted NaN-skipping sum along one axis --------------------------------------------------
// out[f*o_f + s*o_s] = ( sum_k var[f*s_f + s*s_s + k*s_k] * w[k], NaN products skipped ) / D
// MODE 0: D = denom (calc_Zaverage: number of selected points; calc_Raverage: nansum(select*r))
// MODE 1: D = sum of w[k] over the terms that were not skipped (nanmean over the selected rows,
//         calc_RZaverage :47)
// Thread-per-output variant: consecutive threads <-> consecutive `f` (the unit-stride index of the
// kept axes), marching over k: every step is one coalesced row.
template <int MODE>
__global__ void __launch_bounds__(256)
k_axis_wsum(const double *__restrict__ var, int64_t n_f, int64_t n_s, int64_t n_k, int64_t s_f, int64_t s_s,
            int64_t s_k, const double *__restrict__ w, double denom, double *__restrict__ out, int64_t o_f,
            int64_t o_s)
{
    const int64_t total = n_f * n_s;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = q / n_f, f = q - s * n_f;
        const double *p = var + f * s_f + s * s_s;
        double acc = 0.0, wsum = 0.0;
        // 32 terms are loaded before they are added IN ORDER (8 / 16 / 32: 0.039 / 0.039 / 0.034 ms on the S3 grid): the sum is sequential by contract (NumPy's order
        // along a strided axis), but its loads need not wait for each other
        constexpr int U = 32;
        int64_t k = 0;
        for (; k + U <= n_k; k += U) {
            double wk[U], t[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                wk[j] = __ldg(w + k + j);
                t[j] = __ldg(p + (k + j) * s_k);
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const double x = t[j] * wk[j];
                if (!isnan(x)) {
                    acc += x;
                    wsum += wk[j];
                }
            }
        }
        for (; k < n_k; ++k) {
            const double wk = __ldg(w + k);
            const double t = __ldg(p + k * s_k) * wk;
            if (!isnan(t)) {
                acc += t;
                wsum += wk;
            }
        }
        out[f * o_f + s * o_s] = acc / (MODE == 0 ? denom : wsum);
    }
}

// MT_WSUM_PAIRWISE: the axis that is CONTIGUOUS IN THE REFERENCE'S ARRAY (the last axis of its C-order
// array, whatever the strides of the device layout are).  There np.nansum(..., axis=-1) adds PAIRWISE (DOUBLE_pairwise_sum of
// numpy/core/src/umath/loops_utils.h.src): fewer than 8 terms sequentially from 0; up to 128 terms
// in eight interleaved accumulators r[j] += a[i + j] combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
// plus the remainder in order; longer runs split at n/2 rounded down to a multiple of 8, left half
// first.  NaN terms take part as 0.0 in their position.  One thread per output follows exactly that
// order (checked against NumPy for 1 .. 6.5 M terms with a Python mirror of this routine), so the
// result is bit-identical to the reference's; rows are short (a grid axis), the loads go through L1.
template <class Term>
__device__ __forceinline__ double numpy_pairwise_leaf(Term term, int64_t lo, int64_t n)   // n <= 128
{
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += term(lo + i);
        return res;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = term(lo + j);
        const int64_t m = n - (n % 8);
        for (int64_t i = 8; i < m; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += term(lo + i + j);
        }
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (int64_t i = m; i < n; ++i) res += term(lo + i);
        return res;
    }
    return 0.0;   // not reached
}

template <class Term>
__device__ double numpy_pairwise(Term term, int64_t lo, int64_t n)
{
    if (n <= 128) return numpy_pairwise_leaf(term, lo, n);
    // depth-first, left half first, with an explicit stack (2^40 terms would need 34 frames)
    int64_t st_lo[40], st_n[40];
    double st_acc[40];
    int st_state[40], sp = 0;
    st_lo[0] = lo, st_n[0] = n, st_state[0] = 0;
    double ret = 0.0;
    while (sp >= 0) {
        const int64_t cn = st_n[sp], clo = st_lo[sp];
        if (cn <= 128) {
            ret = numpy_pairwise_leaf(term, clo, cn);
            --sp;
            continue;
        }
        int64_t n2 = cn / 2;
        n2 -= n2 % 8;
        if (st_state[sp] == 0) {          // descend into the left half
            st_state[sp] = 1;
            ++sp;
            st_lo[sp] = clo, st_n[sp] = n2, st_state[sp] = 0;
        } else if (st_state[sp] == 1) {   // left done: keep it, descend into the right half
            st_acc[sp] = ret;
            st_state[sp] = 2;
            ++sp;
            st_lo[sp] = clo + n2, st_n[sp] = cn - n2, st_state[sp] = 0;
        } else {                          // both done
            ret = st_acc[sp] + ret;
            --sp;
        }
    }
    return ret;
}

template <int MODE>
__global__ void __launch_bounds__(128)
k_axis_wsum_pairwise(const double *__restrict__ var, int64_t n_f, int64_t n_s, int64_t n_k, int64_t s_f,
                     int64_t s_s, int64_t s_k, const double *__restrict__ w, double denom,
                     double *__restrict__ out, int64_t o_f, int64_t o_s)
{
    const int64_t total = n_f * n_s;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = q / n_f, f = q - s * n_f;
        const double *p = var + f * s_f + s * s_s;
        const double acc = numpy_pairwise(
            [&](int64_t k) {
                const double t = __ldg(p + k * s_k) * __ldg(w + k);
                return isnan(t) ? 0.0 : t;
            },
            0, n_k);
        double wsum = 0.0;   // MODE 1: the weights are 0 / 1 selectors, their sum is exact in any order
        if (MODE == 1)
            for (int64_t k = 0; k < n_k; ++k) {
                const double wk = __ldg(w + k);
                if (!isnan(__ldg(p + k * s_k) * wk)) wsum += wk;
            }
        out[f * o_f + s * o_s] = acc / (MODE == 0 ? denom : wsum);
    }
}

// ---- azimuthal statistics ------------------------------------------------------------------------
// sym = nanmean over theta (:225), asy = var - sym (:226-227),
// axisymmetry = sym**2 / (sym**2 + nansum(asy**2, theta) * dtheta / 2 / pi)   (:199-203)
// One thread per (z, r); consecutive threads <-> the unit-stride index of the layout.
__global__ void __launch_bounds__(128)
k_azimuthal_stats(const double *__restrict__ var, int nz, int nt, int nr, int64_t sz, int64_t st, int64_t sr,
                  int z_fast, double dtheta, double *__restrict__ sym, double *__restrict__ asy,
                  double *__restrict__ axisym)
{
    const int64_t total = (int64_t)nz * nr;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
         q += (int64_t)gridDim.x * blockDim.x) {
        int iz, ir;
        if (z_fast) {
            iz = (int)(q % nz);
            ir = (int)(q / nz);
        } else {
            ir = (int)(q % nr);
            iz = (int)(q / nr);
        }
        const int64_t base = iz * sz + ir * sr;
        double s = 0.0;
        int cnt = 0;
        constexpr int U = 16;   // loads in flight; the additions stay in order (8 / 16 / 32: 0.064 / 0.049 / 0.050 ms, S3 grid)
        int t0 = 0;
        for (; t0 + U <= nt; t0 += U) {
            double v[U];
#pragma unroll
            for (int j = 0; j < U; ++j) v[j] = __ldg(var + base + (int64_t)(t0 + j) * st);
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (!isnan(v[j])) {
                    s += v[j];
                    ++cnt;
                }
        }
        for (int t = t0; t < nt; ++t) {
            const double v = __ldg(var + base + t * st);
            if (!isnan(v)) {
                s += v;
                ++cnt;
            }
        }
        const double mean = cnt ? s / (double)cnt : nan_f64();
        if (sym) sym[(int64_t)iz * nr + ir] = mean;  // (nz, nr) C-order
        double ssq = 0.0;
        if (asy || axisym) {
            int t1 = 0;
            for (; t1 + U <= nt; t1 += U) {
                double v[U];
#pragma unroll
                for (int j = 0; j < U; ++j) v[j] = __ldg(var + base + (int64_t)(t1 + j) * st);
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const double a = v[j] - mean;
                    if (asy) asy[base + (int64_t)(t1 + j) * st] = a;
                    const double a2 = a * a;
                    if (!isnan(a2)) ssq += a2;
                }
            }
            for (int t = t1; t < nt; ++t) {
                const double a = __ldg(var + base + t * st) - mean;
                if (asy) asy[base + t * st] = a;
                const double a2 = a * a;
                if (!isnan(a2)) ssq += a2;
            }
        }
        if (axisym) {
            const double m2 = mean * mean;
            axisym[(int64_t)iz * nr + ir] = m2 / (m2 + ssq * dtheta / 2 / 3.141592653589793);
        }
    }
}

// ---- nansum of a*b*c over a 3-D grid with a separable weight ---------------------------------------
// partial[block] = sum over the block's points of (a*b) * (((wr[r]) * wz[z]) * fr[r]), NaN skipped:
// rho * kinetic_energy * volume of cylindrical_grid.py:566-575 with volume = r*dr*dtheta*dz (wr,
// evaluated by NumPy on the host) times the z and r range filters.  b may be NULL (a alone).
__global__ void __launch_bounds__(256)
k_nansum_sep(const double *__restrict__ a, const double *__restrict__ b, int nz, int nt, int nr, int64_t sz,
             int64_t st, int64_t sr, int z_fast, const double *__restrict__ wr, const double *__restrict__ wz,
             const double *__restrict__ fr, const double *__restrict__ w2d, double *__restrict__ partial)
{
    const int64_t total = (int64_t)nz * nt * nr;
    double acc = 0.0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
         q += (int64_t)gridDim.x * blockDim.x) {
        int iz, it, ir;
        if (z_fast) {  // memory order (t, r, z)
            iz = (int)(q % nz);
            const int64_t q2 = q / nz;
            ir = (int)(q2 % nr);
            it = (int)(q2 / nr);
        } else {  // memory order (z, t, r)
            ir = (int)(q % nr);
            const int64_t q2 = q / nr;
            it = (int)(q2 % nt);
            iz = (int)(q2 / nt);
        }
        const int64_t i = iz * sz + it * st + ir * sr;
        double vol;
        if (w2d) vol = __ldg(w2d + (int64_t)it * nr + ir);           // explicit (theta, r) weight
        else vol = (__ldg(wr + ir) * __ldg(wz + iz)) * __ldg(fr + ir);
        double t = __ldg(a + i);
        if (b) t = t * __ldg(b + i);
        t = t * vol;
        if (!isnan(t)) acc += t;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(256)
k_sum_partials(const double *__restrict__ partial, int n, double *__restrict__ out)
{
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// ---- 1-c-1 smoother, one Jacobi pass ---------------------------------------------------------------
// C-order (n0, n1, n2) array.  `cur` holds the previous pass, `v0` the ORIGINAL field: a neighbour
// outside the array is the frozen ghost cell = v0 at the edge point itself (cartesian_grid.py:159-164,
// :186-194, :216-228).  Sum order of the reference:
//   NS=1 (axis A):      (prev + next) + cur*cw                                   / (2 + cw)   :165-166
//   NS=2 (axes A, B):   (((A- + A+) + B-) + B+) + cur*cw                         / (4 + cw)   :195-197
//   NS=3:               (((((0- + 0+) + 1-) + 1+) + 2+) + 2-) + cur*cw           / (6 + cw)   :224-228
// Thread <-> i2 (contiguous), blockIdx.y <-> i1, blockIdx.z (+ a stride loop of MT_SMOOTH_U planes per trip) <-> i0.
// One plane per trip at eight resident CTAs of 32 registers measured best (U = 1 / 2 / 4 / 8: 0.61 / 0.52 / 0.52 / 0.41 of
// the measured peak on the contiguous axis of 80 x 1000 x 1000: occupancy beats loads in flight here): no integer division, no indexed local arrays (the first version decoded a flat index with
// three 64-bit divisions and kept idx[] / n[] / str[] in local memory: 0.41-0.68 ms per pass on 80 x 1000 x 1000).
#ifndef MT_SMOOTH_U
#define MT_SMOOTH_U 1
#endif
#ifndef MT_SMOOTH_MINB
#define MT_SMOOTH_MINB 8
#endif
template <int NS>
__global__ void __launch_bounds__(256, MT_SMOOTH_MINB)
k_smooth_pass(const double *__restrict__ v0, const double *__restrict__ cur, double *__restrict__ out, int n0,
              int n1, int n2, int axA, int axB, double cw, double denom, double rdenom)
{
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i2 >= n2) return;
    const int64_t s0 = (int64_t)n1 * n2, s1 = n2;
    const mt::CDiv dd(denom, rdenom);
    constexpr int U = MT_SMOOTH_U;
    for (int i1 = blockIdx.y; i1 < n1; i1 += gridDim.y)
    for (int ib = blockIdx.z; ib < n0; ib += U * gridDim.z) {
        double res[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i0 = ib + u * gridDim.z;
            res[u] = 0.0;
            if (i0 >= n0) continue;
            const int64_t q = i0 * s0 + i1 * s1 + i2;
            // neighbour along `ax`, d = -1 / +1; outside the array: the frozen ghost cell = v0 at the edge point
            auto nb = [&](int ax, int d) {
                const int pos = ax == 0 ? i0 : (ax == 1 ? i1 : i2), lim = ax == 0 ? n0 : (ax == 1 ? n1 : n2);
                const int64_t str = ax == 0 ? s0 : (ax == 1 ? s1 : 1);
                const int j = pos + d;
                return (j < 0 || j >= lim) ? __ldg(v0 + q) : __ldg(cur + q + d * str);
            };
            double s;
            if (NS == 1) {
                s = nb(axA, -1) + nb(axA, +1);
            } else if (NS == 2) {
                s = nb(axA, -1) + nb(axA, +1);
                s = s + nb(axB, -1);
                s = s + nb(axB, +1);
            } else {
                s = nb(0, -1) + nb(0, +1);
                s = s + nb(1, -1);
                s = s + nb(1, +1);
                s = s + nb(2, +1);
                s = s + nb(2, -1);
            }
            res[u] = s + __ldg(cur + q) * cw;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i0 = ib + u * gridDim.z;
            if (i0 < n0) mt::st_stream(out + i0 * s0 + i1 * s1 + i2, dd(res[u]));
        }
    }
}

// planes per thread = n0 / gridDim.z: MT_SMOOTH_PPT (default 80: a thread walks up to 80 planes, four per trip, so the
// axis-0 neighbours of a trip are the planes the previous trip just touched; 4 / 8 / 16 / 40 measured no better)
unsigned smooth_gz(int64_t n0)
{
    static const int ppt = [] {
        const char *e = getenv("MT_SMOOTH_PPT");
        const int v = e ? atoi(e) : 80;
        return v < 1 ? 1 : v;
    }();
    const int64_t gz = (n0 + ppt - 1) / ppt;
    return (unsigned)(gz < 1 ? 1 : (gz > 64 ? 64 : gz));
}

}  // namespace

extern "C" {

int mt_axis_wsum(const double *var, int64_t n_a, int64_t n_b, int64_t n_k, int64_t s_a, int64_t s_b,
                 int64_t s_k, const double *w, int mode, double denom, double *out, int64_t o_a,
                 int64_t o_b, mt_stream_t stream)
{
    MT_REQUIRE(var && w && out && n_a >= 1 && n_b >= 1 && n_k >= 1, "mt_axis_wsum: bad arguments");
    MT_REQUIRE(mode >= 0 && mode <= 3, "mt_axis_wsum: mode = (0: fixed denominator | 1: weight sum) + MT_WSUM_PAIRWISE");
    const bool pairwise = (mode & MT_WSUM_PAIRWISE) != 0;
    mode &= 1;
    cudaStream_t s = (cudaStream_t)stream;
    // the kept axis with the smaller stride becomes the thread-fastest one
    int64_t n_f = n_a, s_f = s_a, o_f = o_a, n_s = n_b, s_s = s_b, o_s = o_b;
    if (s_b < s_a) {
        n_f = n_b, s_f = s_b, o_f = o_b;
        n_s = n_a, s_s = s_a, o_s = o_a;
    }
    mt::ProfScope prof("k_axis_wsum", stream);
    if (pairwise) {   // the reference reduces along its contiguous axis: NumPy's pairwise order
        const unsigned grid = mt::grid_for(n_a * n_b, 128, 8);
        if (mode == 0)
            k_axis_wsum_pairwise<0><<<grid, 128, 0, s>>>(var, n_f, n_s, n_k, s_f, s_s, s_k, w, denom, out, o_f, o_s);
        else
            k_axis_wsum_pairwise<1><<<grid, 128, 0, s>>>(var, n_f, n_s, n_k, s_f, s_s, s_k, w, denom, out, o_f, o_s);
    } else {
        const unsigned grid = mt::grid_for(n_a * n_b, 256, 8);
        if (mode == 0)
            k_axis_wsum<0><<<grid, 256, 0, s>>>(var, n_f, n_s, n_k, s_f, s_s, s_k, w, denom, out, o_f, o_s);
        else
            k_axis_wsum<1><<<grid, 256, 0, s>>>(var, n_f, n_s, n_k, s_f, s_s, s_k, w, denom, out, o_f, o_s);
    }
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_azimuthal_stats(const double *var, const mt_grid_t *g, double *sym, double *asy, double *axisym,
                       mt_stream_t stream)
{
    MT_REQUIRE(var && g && g->nz >= 1 && g->nt >= 1 && g->nr >= 1 && (sym || asy || axisym),
               "mt_azimuthal_stats: bad arguments");
    MT_REQUIRE(g->layout == MT_LAYOUT_ZTR || g->layout == MT_LAYOUT_TRZ, "mt_azimuthal_stats: unknown layout");
    const int nz = (int)g->nz, nt = (int)g->nt, nr = (int)g->nr;
    const bool trz = g->layout == MT_LAYOUT_TRZ;
    const int64_t sz = trz ? 1 : (int64_t)nt * nr, st = trz ? (int64_t)nr * nz : nr, sr = trz ? nz : 1;
    mt::ProfScope prof("k_azimuthal_stats", stream);
    k_azimuthal_stats<<<mt::grid_for((int64_t)nz * nr, 128, 8), 128, 0, (cudaStream_t)stream>>>(
        var, nz, nt, nr, sz, st, sr, trz ? 1 : 0, g->dt, sym, asy, axisym);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_nansum_weighted(const double *a, const double *b, const mt_grid_t *g, const double *wr, const double *wz,
                       const double *fr, const double *w2d, double *scratch, double *out, mt_stream_t stream)
{
    MT_REQUIRE(a && g && scratch && out && g->nz >= 1 && g->nt >= 1 && g->nr >= 1, "mt_nansum_weighted: bad arguments");
    MT_REQUIRE(w2d || (wr && wz && fr), "mt_nansum_weighted: give either the (theta, r) weight or wr, wz, fr");
    MT_REQUIRE(g->layout == MT_LAYOUT_ZTR || g->layout == MT_LAYOUT_TRZ, "mt_nansum_weighted: unknown layout");
    const int nz = (int)g->nz, nt = (int)g->nt, nr = (int)g->nr;
    const bool trz = g->layout == MT_LAYOUT_TRZ;
    const int64_t sz = trz ? 1 : (int64_t)nt * nr, st = trz ? (int64_t)nr * nz : nr, sr = trz ? nz : 1;
    unsigned grid = mt::grid_for((int64_t)nz * nt * nr, 256, 4);
    if (grid > MT_NANSUM_SCRATCH) grid = MT_NANSUM_SCRATCH;
    cudaStream_t s = (cudaStream_t)stream;
    mt::ProfScope prof("k_nansum_sep", stream);
    k_nansum_sep<<<grid, 256, 0, s>>>(a, b, nz, nt, nr, sz, st, sr, trz ? 1 : 0, wr, wz, fr, w2d, scratch);
    MT_LAUNCH_CHECK();
    k_sum_partials<<<1, 256, 0, s>>>(scratch, (int)grid, out);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_smooth_pass(const double *v0, const double *cur, double *out, int64_t n0, int64_t n1, int64_t n2,
                   int naxes, int ax_a, int ax_b, double center_weight, double denom, mt_stream_t stream)
{
    MT_REQUIRE(v0 && cur && out && cur != out && n0 >= 1 && n1 >= 1 && n2 >= 1 && n0 < 2147483647LL &&
                   n1 < 2147483647LL && n2 < 2147483647LL,
               "mt_smooth_pass: bad arguments");
    MT_REQUIRE(naxes >= 1 && naxes <= 3, "mt_smooth_pass: naxes must be 1, 2 or 3");
    MT_REQUIRE(naxes == 3 || (ax_a >= 0 && ax_a <= 2), "mt_smooth_pass: bad axis");
    MT_REQUIRE(naxes != 2 || (ax_b >= 0 && ax_b <= 2 && ax_b != ax_a), "mt_smooth_pass: bad second axis");
    // threads walk the last axis: an array whose last axis has one element and is not smoothed (smooth1d along the
    // last axis of the caller's array arrives as (outer, n, 1)) is the same array with its axes shifted by one
    if (n2 == 1 && naxes < 3 && ax_a != 2 && (naxes == 1 || ax_b != 2)) {
        n2 = n1, n1 = n0, n0 = 1;
        ax_a += 1, ax_b += 1;
    }
    const int threads = n2 >= 256 ? 256 : (n2 >= 128 ? 128 : (n2 >= 64 ? 64 : 32));
    const dim3 grid((unsigned)((n2 + threads - 1) / threads), (unsigned)(n1 < 65535 ? n1 : 65535),
                    (unsigned)smooth_gz(n0));
    cudaStream_t s = (cudaStream_t)stream;
    mt::ProfScope prof("k_smooth_pass", stream);
    const double rd = 1.0 / denom;
    if (naxes == 1)
        k_smooth_pass<1><<<grid, threads, 0, s>>>(v0, cur, out, (int)n0, (int)n1, (int)n2, ax_a, 0, center_weight, denom, rd);
    else if (naxes == 2)
        k_smooth_pass<2><<<grid, threads, 0, s>>>(v0, cur, out, (int)n0, (int)n1, (int)n2, ax_a, ax_b, center_weight, denom, rd);
    else
        k_smooth_pass<3><<<grid, threads, 0, s>>>(v0, cur, out, (int)n0, (int)n1, (int)n2, 0, 1, center_weight, denom, rd);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

}  // extern "C"
