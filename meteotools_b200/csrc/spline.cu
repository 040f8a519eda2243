// Spline vertical interpolation: the `ver_interp_order = 2 / 3` branch of
// cylindrical_grid.set_data (grids/cylindrical_grid.py:83-93, 123-127) and cartesian_grid.set_data
// (grids/cartesian_grid.py:61-74, 108-112).  Per column the reference calls
//     cs = scipy.interpolate.splrep(vert[:, j, i], var[:, j, i], k=order);  out = splev(levels, cs)
// i.e. FITPACK's interpolating spline (s = 0, unit weights): knots by the rule of fpcurf (data
// points for odd k, midpoints for even k), observation matrix built row by row with fpbspl and
// reduced by Givens rotations (fpgivs / fprota), back substitution (fpback), evaluation with splev
// (ext = 0: the end polynomials extrapolate).  This file follows that published algorithm
// (P. Dierckx, FITPACK: fpcurf.f, fpbspl.f, fpgivs.f, fprota.f, fpback.f, splev.f) operation by
// operation in float64 with -fmad=false, so the results agree with scipy to rounding.
//
// One thread per (column, variable); the band matrix a(m, k+1), the knots t(m+k+1) and the
// right-hand side z(m) of a thread live in a global workspace laid out [element][thread], so that
// the threads of a warp (adjacent columns) touch adjacent words.  The reference spends minutes
// here (a Python double loop over all columns); this path only works on the columns under the
// targets' stencils and is far from any roofline that matters -- it is built for coverage.
#include "common.cuh"

namespace {

struct SplPtrs {
    const void *field[MT_MAX_VARS];
    double *out[MT_MAX_VARS];
};

__device__ int g_spline_status;   // bit 0: a column's coordinate decreases (scipy: ValueError)

template <typename T, int K>
__global__ void __launch_bounds__(128)
k_spline_levels(SplPtrs p, int nvar, const T *__restrict__ vert, int m, int64_t ny, int64_t nx, int64_t y0,
                int64_t x0, int by, int bx, const double *__restrict__ levels, int nlev, int64_t o_lev,
                int64_t o_y, int64_t o_x, double *__restrict__ ws)
{
    constexpr int K1 = K + 1, K2 = K + 2;
    const int64_t NT = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int n = m + K1, nk1 = m;   // number of knots; n - k1 coefficients
    double *wt = ws + gtid, *wa = wt + (int64_t)n * NT, *wz = wa + (int64_t)m * K1 * NT;
    // 1-based accessors in the notation of the Fortran sources
#define T_(j) wt[(int64_t)((j)-1) * NT]
#define A_(i, j) wa[(int64_t)(((i)-1) * K1 + ((j)-1)) * NT]
#define Z_(i) wz[(int64_t)((i)-1) * NT]
    const int64_t plane = ny * nx, ncol = (int64_t)by * bx, items = ncol * nvar;

    auto fpbspl = [&](double x, int l, double *h) {   // fpbspl.f: the k+1 non-zero B-splines at x
        double hh[K1];
        h[0] = 1.0;
#pragma unroll
        for (int j = 1; j <= K; ++j) {
#pragma unroll
            for (int i = 0; i < j; ++i) hh[i] = h[i];
            h[0] = 0.0;
#pragma unroll
            for (int i = 1; i <= j; ++i) {
                const int li = l + i, lj = li - j;
                const double tli = T_(li), tlj = T_(lj);
                if (tli == tlj) {
                    h[i] = 0.0;
                } else {
                    const double f = hh[i - 1] / (tli - tlj);
                    h[i - 1] = h[i - 1] + f * (tli - x);
                    h[i] = f * (x - tlj);
                }
            }
        }
    };

    for (int64_t item = gtid; item < items; item += NT) {
        const int v = (int)(item / ncol);
        const int64_t col = item % ncol;
        const int64_t yy = col / bx, xx = col % bx;
        const T *xc = vert + (y0 + yy) * nx + (x0 + xx);
        const T *yc = reinterpret_cast<const T *>(p.field[v]) + (y0 + yy) * nx + (x0 + xx);
#define X_(i) ((double)xc[(int64_t)((i)-1) * plane])
#define Y_(i) ((double)yc[(int64_t)((i)-1) * plane])
        // curfit.f input check: x must not decrease
        bool bad = false;
        for (int i = 2; i <= m; ++i) bad = bad || (X_(i - 1) > X_(i));
        if (bad) atomicOr(&g_spline_status, 1);
        // knots (fpcurf.f, s = 0): k+1 copies of the end points, interior knots at the data points
        // (odd k) or half way between them (even k)
        {
            const double xb = X_(1), xe = X_(m);
            int i = n;
            for (int j = 1; j <= K1; ++j) {
                T_(j) = xb;
                T_(i) = xe;
                --i;
            }
            constexpr int K3 = K / 2;
            int jj = K3 + 2;
            for (int q = K2; q <= m; ++q, ++jj)
                T_(q) = (K3 * 2 == K) ? (X_(jj) + X_(jj - 1)) * 0.5 : X_(jj);
        }
        for (int i = 1; i <= nk1; ++i) {
            Z_(i) = 0.0;
#pragma unroll
            for (int j = 1; j <= K1; ++j) A_(i, j) = 0.0;
        }
        // observation matrix, row by row, rotated into the triangle (fpcurf.f)
        int l = K1;
        for (int it = 1; it <= m; ++it) {
            const double xi = X_(it);
            double yi = Y_(it);
            while (!(xi < T_(l + 1) || l == nk1)) ++l;
            double h[K1];
            fpbspl(xi, l, h);
            int j = l - K1;
#pragma unroll
            for (int i = 1; i <= K1; ++i) {
                ++j;
                const double piv = h[i - 1];
                if (piv == 0.0) continue;
                // fpgivs.f
                double ww = A_(j, 1);
                const double store = fabs(piv);
                double dd = mt::nan_f64();
                if (store >= ww) {
                    const double r = ww / piv;
                    dd = store * sqrt(1.0 + r * r);
                }
                if (store < ww) {
                    const double r = piv / ww;
                    dd = ww * sqrt(1.0 + r * r);
                }
                const double c = ww / dd, s = piv / dd;
                A_(j, 1) = dd;
                {   // fprota.f on the right-hand side
                    const double s1 = yi, s2 = Z_(j);
                    Z_(j) = c * s2 + s * s1;
                    yi = c * s1 - s * s2;
                }
                if (i == K1) break;
                int i2 = 1;
#pragma unroll
                for (int i1 = i + 1; i1 <= K1; ++i1) {
                    ++i2;
                    const double s1 = h[i1 - 1], s2 = A_(j, i2);
                    A_(j, i2) = c * s2 + s * s1;
                    h[i1 - 1] = c * s1 - s * s2;
                }
            }
        }
        // fpback.f: the B-spline coefficients overwrite z
        Z_(nk1) = Z_(nk1) / A_(nk1, 1);
        {
            int i = nk1 - 1;
            for (int j = 2; j <= nk1; ++j) {
                double store = Z_(i);
                const int i1 = j <= K ? j - 1 : K;
                int mm = i;
                for (int q = 1; q <= i1; ++q) {
                    ++mm;
                    store = store - Z_(mm) * A_(i, q + 1);
                }
                Z_(i) = store / A_(i, 1);
                --i;
            }
        }
        // splev.f, ext = 0
        double *o = p.out[v] + yy * o_y + xx * o_x;
        int ls = K1, l1 = K1 + 1;
        for (int lv = 0; lv < nlev; ++lv) {
            const double arg = levels[lv];
            while (!(arg >= T_(ls) || l1 == K2)) {
                l1 = ls;
                ls = ls - 1;
            }
            while (!(arg < T_(l1) || ls == nk1)) {
                ls = l1;
                l1 = ls + 1;
            }
            double h[K1];
            fpbspl(arg, ls, h);
            double sp = 0.0;
            int ll = ls - K1;
#pragma unroll
            for (int j = 1; j <= K1; ++j) {
                ++ll;
                sp = sp + Z_(ll) * h[j - 1];
            }
            o[lv * o_lev] = sp;
        }
#undef X_
#undef Y_
    }
#undef T_
#undef A_
#undef Z_
}

__global__ void k_spline_status_clear() { g_spline_status = 0; }

}  // namespace

extern "C" {

int mt_spline_levels(const void *const *fields, int nvar, const void *vert, int dtype, int64_t nz, int64_t ny,
                     int64_t nx, int64_t y0, int64_t y1, int64_t x0, int64_t x1, const double *levels,
                     int64_t nlev, int order, double *const *out, int64_t o_lev, int64_t o_y, int64_t o_x,
                     mt_stream_t stream)
{
    MT_REQUIRE(fields && vert && levels && out, "mt_spline_levels: null pointer");
    MT_REQUIRE(nvar >= 1 && nvar <= MT_MAX_VARS, "mt_spline_levels: nvar must be 1..%d", MT_MAX_VARS);
    MT_REQUIRE(dtype == MT_F64 || dtype == MT_F32, "mt_spline_levels: dtype must be MT_F64 or MT_F32");
    MT_REQUIRE(order >= 1 && order <= 3, "mt_spline_levels: spline order must be 1, 2 or 3");
    MT_REQUIRE(nz > order && nz < 32000 && nlev >= 1 && nlev < 2147483647LL,
               "mt_spline_levels: m > k must hold (scipy.interpolate.splrep), got %lld levels for k = %d",
               (long long)nz, order);
    MT_REQUIRE(0 <= y0 && y0 < y1 && y1 <= ny && 0 <= x0 && x0 < x1 && x1 <= nx && (y1 - y0) < 2147483647LL &&
                   (x1 - x0) < 2147483647LL,
               "mt_spline_levels: column box outside the array");
    SplPtrs p;
    for (int v = 0; v < nvar; ++v) {
        MT_REQUIRE(fields[v] && out[v], "mt_spline_levels: null field %d", v);
        p.field[v] = fields[v];
        p.out[v] = out[v];
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int by = (int)(y1 - y0), bx = (int)(x1 - x0);
    const int64_t items = (int64_t)by * bx * nvar;
    const unsigned grid = mt::grid_for(items, 128, 4);
    const int64_t nthreads = (int64_t)grid * 128;
    const int64_t per_thread = (nz + order + 1) + nz * (order + 1) + nz;   // t, a, z
    double *ws = nullptr;
    MT_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ws), (size_t)(nthreads * per_thread) * sizeof(double), s));
    {
        mt::ProfScope prof("k_spline_levels", stream);
#define SPL_LAUNCH(T, K)                                                                                          \
    k_spline_levels<T, K><<<grid, 128, 0, s>>>(p, nvar, reinterpret_cast<const T *>(vert), (int)nz, ny, nx, y0, x0, \
                                               by, bx, levels, (int)nlev, o_lev, o_y, o_x, ws)
#define SPL_DISPATCH(T)                 \
    do {                                \
        if (order == 1) SPL_LAUNCH(T, 1); \
        else if (order == 2) SPL_LAUNCH(T, 2); \
        else SPL_LAUNCH(T, 3);          \
    } while (0)
        if (dtype == MT_F64) SPL_DISPATCH(double);
        else SPL_DISPATCH(float);
#undef SPL_DISPATCH
#undef SPL_LAUNCH
    }
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(ws, s);
    mt::count_launch();
    if (le != cudaSuccess) return mt::cuda_fail(le, "kernel launch", __FILE__, __LINE__);
    return MT_OK;
}

int mt_spline_status(mt_stream_t stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    int h = 0;
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(&h, g_spline_status, sizeof(int)) != cudaSuccess) return -1;
    if (h) {
        k_spline_status_clear<<<1, 1, 0, s>>>();
        cudaStreamSynchronize(s);
    }
    return h;
}

}  // extern "C"
