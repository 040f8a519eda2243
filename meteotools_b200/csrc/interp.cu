// Regular-grid interpolation kernels: B200 replacements of the eight
// OpenMP routines of the reference's fastcompute.f90.  Per-point semantics are those of the
// Fortran (file:line cited at each rule); the parallel decomposition is not:
//   * the hot kernel (bilinear gather, layer-fastest "TRZ" layout) runs one WARP per target
//     point with the lanes spread over the contiguous layer index, two layers per lane
//     (16-byte loads), so each of the four corner gathers is one fully coalesced segment
//     and the index/weight arithmetic is done once per point (a reusable "plan");
//   * C-order sources use one thread per target point marching over a chunk of layers.
// Everything is IEEE double with -fmad=false: index and weight arithmetic is bit-identical
// to the reference (x/dx with a true division, truncation by (int)).
#include "common.cuh"

namespace {

using mt::nan_f64;

// ------------------------------------------------------------------------------------ plans
// Per-point stencil of the 2-D routines (fastcompute.f90:80-93): 0-based x0,x1,y0,y1 and the
// two weights.  ix.x == -1: sentinel target (-> 1.7e308), ix.x == -2: NaN target (-> NaN).
__device__ __forceinline__ void plan2d_point(double xl, double yl, double dx, double dy, int lenx,
                                             int leny, int4 &ix, double2 &w)
{
    if (xl == MT_SENTINEL || yl == MT_SENTINEL) {
        ix = make_int4(-1, 0, 0, 0);
        w = make_double2(0.0, 0.0);
        return;
    }
    if (isnan(xl) || isnan(yl)) {
        ix = make_int4(-2, 0, 0, 0);
        w = make_double2(0.0, 0.0);
        return;
    }
    double xt = xl / dx, yt = yl / dy;                 // :80-81 (true IEEE divisions)
    int x0 = (int)xt, y0 = (int)yt;                    // Fortran int() truncation, 0-based
    int x1 = (xt == (double)x0) ? x0 : x0 + 1;         // :83-87 exact-node rule
    int y1 = (yt == (double)y0) ? y0 : y0 + 1;         // :89-93
    w = make_double2(xt - (double)x0, yt - (double)y0);  // x/dx-(x0-1) with 1-based x0
    // clamp (the reference would read out of bounds; unreachable for validated targets)
    x0 = min(max(x0, 0), lenx - 1);
    x1 = min(max(x1, 0), lenx - 1);
    y0 = min(max(y0, 0), leny - 1);
    y1 = min(max(y1, 0), leny - 1);
    ix = make_int4(x0, x1, y0, y1);
}

__global__ void k_plan2d(const double *__restrict__ xloc, const double *__restrict__ yloc, int64_t n,
                         double dx, double dy, int lenx, int leny, int4 *__restrict__ ix,
                         double2 *__restrict__ w)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int4 a;
        double2 b;
        plan2d_point(xloc[i], yloc[i], dx, dy, lenx, leny, a, b);
        ix[i] = a;
        w[i] = b;
    }
}

__device__ __forceinline__ double bilerp(double Z00, double Z10, double Z01, double Z11, double fx,
                                         double fy)
{
    double t0 = (Z10 - Z00) * fx + Z00;  // :100
    double t1 = (Z11 - Z01) * fx + Z01;  // :101
    return (t1 - t0) * fy + t0;          // :102
}

struct VarPtrs {
    const double *src[MT_MAX_VARS];
    double *out[MT_MAX_VARS];
};

// ---------------------------------------------------------------------- hot kernel (a1, TRZ)
// Layer-fastest source  src[(y*lenx_s + x)*L + l]  (generic s_y, s_x, s_l == 1) and layer-fastest
// output out[i*L + l] (o_l == 1, o_n == L or any).  One warp per point; lane handles VEC layers.
// Optional terrain mask: hgt_idx[i] > l  ->  NaN  (cylindrical_grid.py:171-181).
template <int VEC>
__global__ void __launch_bounds__(256)
k_gather2d_lf(VarPtrs vp, int nvar, int layers, int64_t s_y, int64_t s_x, const int4 *__restrict__ pix,
              const double2 *__restrict__ pw, const double *__restrict__ xloc,
              const double *__restrict__ yloc, double dx, double dy, int lenx, int leny, int64_t n,
              int64_t o_n, const int32_t *__restrict__ hgt_idx)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    // targets are dealt round-robin over the grid, so that the output rows of a CTA's eight warps
    // form one contiguous 3840-byte run per variable (sequential output order measured faster
    // than any source-locality order: profiles/experiments_r01.md); a warp handles its point for
    // ALL variables, so the plan entry is read once and the gathers of different variables are
    // independent loads in flight
    const int64_t stride = (int64_t)gridDim.x * warps_per_block;
    int64_t i = blockIdx.x * (int64_t)warps_per_block + (threadIdx.x >> 5);
    if (i >= n) return;
    auto load_plan = [&](int64_t k, int4 &ix, double2 &w, int &terrain) {
        if (pix) {
            ix = pix[k];
            w = pw[k];
        } else {
            plan2d_point(xloc[k], yloc[k], dx, dy, lenx, leny, ix, w);
        }
        terrain = hgt_idx ? hgt_idx[k] : 0;
    };
    int4 ix;
    double2 w;
    int terrain;
    load_plan(i, ix, w, terrain);
    while (i < n) {
        // software pipeline: the next point's plan entry is requested before this point's gathers
        const int64_t inext = i + stride;
        int4 nix = ix;
        double2 nw = w;
        int nter = terrain;
        if (inext < n) load_plan(inext, nix, nw, nter);
        if (ix.x < 0) {
            const double fill = (ix.x == -1) ? MT_SENTINEL : nan_f64();
            for (int v = 0; v < nvar; ++v)
                for (int l = lane; l < layers; l += 32) vp.out[v][i * o_n + l] = fill;
        } else {
            const int64_t a00 = ix.z * s_y + ix.x * s_x, a10 = ix.z * s_y + ix.y * s_x;
            const int64_t a01 = ix.w * s_y + ix.x * s_x, a11 = ix.w * s_y + ix.y * s_x;
            if (VEC == 2) {
                for (int l = lane * 2; l < layers; l += 64) {
#pragma unroll 2
                    for (int v = 0; v < nvar; ++v) {
                        const double *__restrict__ src = vp.src[v] + l;
                        const double2 a = *reinterpret_cast<const double2 *>(src + a00);
                        const double2 b = *reinterpret_cast<const double2 *>(src + a10);
                        const double2 c = *reinterpret_cast<const double2 *>(src + a01);
                        const double2 d = *reinterpret_cast<const double2 *>(src + a11);
                        double2 r;
                        r.x = bilerp(a.x, b.x, c.x, d.x, w.x, w.y);
                        r.y = bilerp(a.y, b.y, c.y, d.y, w.x, w.y);
                        if (terrain > l) r.x = nan_f64();
                        if (terrain > l + 1) r.y = nan_f64();
                        __stcs(reinterpret_cast<double2 *>(vp.out[v] + i * o_n + l), r);
                    }
                }
            } else {
                for (int l = lane; l < layers; l += 32) {
                    for (int v = 0; v < nvar; ++v) {
                        const double *__restrict__ src = vp.src[v] + l;
                        double r = bilerp(src[a00], src[a10], src[a01], src[a11], w.x, w.y);
                        if (terrain > l) r = nan_f64();
                        __stcs(vp.out[v] + i * o_n + l, r);
                    }
                }
            }
        }
        i = inext, ix = nix, w = nw, terrain = nter;
    }
}

// ------------------------------------------------------------------ generic strided 2-D (a1,a2)
// One thread per target point, marching over a chunk of layers (blockIdx.y).  Used for C-order
// sources (s_x == 1: the four gathers of neighbouring points share sectors) and any other layout.
__global__ void __launch_bounds__(256)
k_interp2d_pt(const double *__restrict__ src, int64_t layers, int leny, int lenx, int64_t s_l,
              int64_t s_y, int64_t s_x, double dx, double dy, const double *__restrict__ xloc,
              const double *__restrict__ yloc, int64_t n, double *__restrict__ out, int64_t o_l,
              int64_t o_n, int layer_chunk)
{
    const int64_t l0 = (int64_t)blockIdx.y * layer_chunk;
    const int64_t l1 = min(l0 + (int64_t)layer_chunk, layers);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int4 ix;
        double2 w;
        plan2d_point(xloc[i], yloc[i], dx, dy, lenx, leny, ix, w);
        if (ix.x < 0) {
            const double fill = (ix.x == -1) ? MT_SENTINEL : nan_f64();
            for (int64_t l = l0; l < l1; ++l) out[l * o_l + i * o_n] = fill;
            continue;
        }
        const int64_t a00 = ix.z * s_y + ix.x * s_x, a10 = ix.z * s_y + ix.y * s_x;
        const int64_t a01 = ix.w * s_y + ix.x * s_x, a11 = ix.w * s_y + ix.y * s_x;
#pragma unroll 4
        for (int64_t l = l0; l < l1; ++l) {
            const double *s = src + l * s_l;
            out[l * o_l + i * o_n] = bilerp(s[a00], s[a10], s[a01], s[a11], w.x, w.y);
        }
    }
}

// ------------------------------------------------------------------------------ 3-D (a3)
// fastcompute.f90:133-172: x1 == x0 only when the target sits on the LAST node; z first, then y,
// then x.  Flat index over (point, layer); LAYER_FAST picks which of the two is contiguous.
template <bool LAYER_FAST>
__global__ void __launch_bounds__(256)
k_interp3d(const double *__restrict__ src, int64_t layers, int lenz, int leny, int lenx, int64_t s_l,
           int64_t s_z, int64_t s_y, int64_t s_x, double dx, double dy, double dz,
           const double *__restrict__ xloc, const double *__restrict__ yloc,
           const double *__restrict__ zloc, int64_t n, double *__restrict__ out, int64_t o_l,
           int64_t o_n)
{
    const int64_t total = layers * n;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t l, i;
        if (LAYER_FAST) {
            i = t / layers;
            l = t - i * layers;
        } else {
            l = t / n;
            i = t - l * n;
        }
        const double xl = xloc[i], yl = yloc[i], zl = zloc[i];
        double res;
        if (xl == MT_SENTINEL || yl == MT_SENTINEL || zl == MT_SENTINEL) {
            res = MT_SENTINEL;
        } else if (isnan(xl) || isnan(yl) || isnan(zl)) {
            res = nan_f64();
        } else {
            int x0 = (int)(xl / dx), y0 = (int)(yl / dy), z0 = (int)(zl / dz);
            int x1 = (xl == (double)(lenx - 1) * dx) ? x0 : x0 + 1;  // :135-139
            int y1 = (yl == (double)(leny - 1) * dy) ? y0 : y0 + 1;
            int z1 = (zl == (double)(lenz - 1) * dz) ? z0 : z0 + 1;
            const double zf = zl / dz - (double)z0, yf = yl / dy - (double)y0,
                         xf = xl / dx - (double)x0;  // :163-165
            x0 = min(max(x0, 0), lenx - 1), x1 = min(max(x1, 0), lenx - 1);
            y0 = min(max(y0, 0), leny - 1), y1 = min(max(y1, 0), leny - 1);
            z0 = min(max(z0, 0), lenz - 1), z1 = min(max(z1, 0), lenz - 1);
            const double *s = src + l * s_l;
#define W3(zz, yy, xx) s[(zz) * s_z + (yy) * s_y + (xx) * s_x]
            const double w000 = W3(z0, y0, x0), w010 = W3(z0, y1, x0), w001 = W3(z0, y0, x1),
                         w011 = W3(z0, y1, x1), w100 = W3(z1, y0, x0), w110 = W3(z1, y1, x0),
                         w101 = W3(z1, y0, x1), w111 = W3(z1, y1, x1);
#undef W3
            const double t00 = (w100 - w000) * zf + w000;  // :166-172
            const double t10 = (w110 - w010) * zf + w010;
            const double t01 = (w101 - w001) * zf + w001;
            const double t11 = (w111 - w011) * zf + w011;
            const double t0 = (t10 - t00) * yf + t00;
            const double t1 = (t11 - t01) * yf + t01;
            res = (t1 - t0) * xf + t0;
        }
        out[l * o_l + i * o_n] = res;
    }
}

// ------------------------------------------------------------------------------ 1-D (a4)
template <bool LAYER_FAST>
__global__ void __launch_bounds__(256)
k_interp1d(const double *__restrict__ src, int64_t layers, int lenx, int64_t s_l, int64_t s_x,
           double dx, const double *__restrict__ xloc, int64_t n, double *__restrict__ out,
           int64_t o_l, int64_t o_n)
{
    const int64_t total = layers * n;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t l, i;
        if (LAYER_FAST) {
            i = t / layers;
            l = t - i * layers;
        } else {
            l = t / n;
            i = t - l * n;
        }
        const double xl = xloc[i];
        double res;
        if (xl == MT_SENTINEL) {
            res = MT_SENTINEL;
        } else if (isnan(xl)) {
            res = nan_f64();
        } else {
            int x0 = (int)(xl / dx);
            int x1 = (xl == (double)(lenx - 1) * dx) ? x0 : x0 + 1;  // fastcompute.f90:270-274
            const double f = xl / dx - (double)x0;                  // :279
            x0 = min(max(x0, 0), lenx - 1), x1 = min(max(x1, 0), lenx - 1);
            const double Z0 = src[l * s_l + x0 * s_x], Z1 = src[l * s_l + x1 * s_x];
            res = (Z1 - Z0) * f + Z0;
        }
        out[l * o_l + i * o_n] = res;
    }
}

// Unequal axis: bisection of fastcompute.f90:340-351 (works for increasing or decreasing axes).
template <bool LAYER_FAST>
__global__ void __launch_bounds__(256)
k_interp1d_nonequal(const double *__restrict__ xin, const double *__restrict__ src, int64_t layers,
                    int lenx, int64_t s_l, int64_t s_x, const double *__restrict__ xout, int64_t n,
                    double *__restrict__ out, int64_t o_l, int64_t o_n)
{
    const int64_t total = layers * n;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t l, i;
        if (LAYER_FAST) {
            i = t / layers;
            l = t - i * layers;
        } else {
            l = t / n;
            i = t - l * n;
        }
        const double xo = xout[i];
        double res;
        if (xo == MT_SENTINEL) {
            res = MT_SENTINEL;
        } else if (isnan(xo)) {
            res = nan_f64();
        } else {
            int x0 = 0, x1 = lenx - 1;  // 0-based images of the 1-based (1, lenx)
            while (x1 - x0 > 1) {
                const int mid = (x0 + x1 + 2) / 2 - 1;  // (x0+x1)/2 on the 1-based indices
                if ((xo - xin[mid]) * (xo - xin[x0]) <= 0)
                    x1 = mid;
                else
                    x0 = mid;
            }
            const double Z0 = src[l * s_l + x0 * s_x], Z1 = src[l * s_l + x1 * s_x];
            res = (Z1 - Z0) * ((xo - xin[x0]) / (xin[x1] - xin[x0])) + Z0;  // :354
        }
        out[l * o_l + i * o_n] = res;
    }
}

inline bool fits_int(int64_t v) { return v > 0 && v < 2147483647LL; }

}  // namespace

// =====================================================================================
// device API
// =====================================================================================
extern "C" {

int mt_interp2d_plan(const double *xloc, const double *yloc, int64_t n, double dx, double dy,
                     int64_t lenx, int64_t leny, int32_t *ix4, double *w2, mt_stream_t stream)
{
    MT_REQUIRE(xloc && yloc && ix4 && w2, "mt_interp2d_plan: null pointer");
    MT_REQUIRE(n >= 0 && fits_int(lenx) && fits_int(leny), "mt_interp2d_plan: bad sizes");
    if (n == 0) return MT_OK;
    mt::ProfScope prof("k_plan2d", stream);
    k_plan2d<<<mt::grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        xloc, yloc, n, dx, dy, (int)lenx, (int)leny, reinterpret_cast<int4 *>(ix4),
        reinterpret_cast<double2 *>(w2));
    MT_LAUNCH_CHECK();
    return MT_OK;
}

// Multi-variable layer-fastest gather with an optional plan and optional terrain mask.
int mt_gather2d_lf(const double *const *src, double *const *out, int nvar, int64_t layers,
                   int64_t leny, int64_t lenx, int64_t s_y, int64_t s_x, double dx, double dy,
                   const double *xloc, const double *yloc, const int32_t *plan_ix,
                   const double *plan_w, int64_t n, int64_t o_n, const int32_t *hgt_idx,
                   mt_stream_t stream)
{
    MT_REQUIRE(src && out && nvar >= 1 && nvar <= MT_MAX_VARS, "mt_gather2d_lf: nvar must be 1..%d",
               MT_MAX_VARS);
    MT_REQUIRE((plan_ix && plan_w) || (xloc && yloc), "mt_gather2d_lf: need a plan or xloc/yloc");
    MT_REQUIRE(fits_int(layers) && fits_int(lenx) && fits_int(leny) && n >= 0 && o_n >= layers,
               "mt_gather2d_lf: bad sizes");
    if (n == 0) return MT_OK;
    VarPtrs vp;
    bool vec = (layers % 2 == 0) && (s_y % 2 == 0) && (s_x % 2 == 0) && (o_n % 2 == 0);
    for (int v = 0; v < nvar; ++v) {
        MT_REQUIRE(src[v] && out[v], "mt_gather2d_lf: null field %d", v);
        vp.src[v] = src[v];
        vp.out[v] = out[v];
        vec = vec && (((uintptr_t)src[v] | (uintptr_t)out[v]) % 16 == 0);
    }
    // grid: 16 CTAs per SM was the best of {5,8,16,24,32}; one point per warp iteration beat 2 and 4
    // points in flight and a cp.async.bulk (TMA) ring -- profiles/experiments_r01.md
    const unsigned grid = mt::grid_for(n, 8, 4);  // 4 CTAs/SM best of {2,4,5,8,16}
    mt::ProfScope prof("k_gather2d_lf", stream);
    auto pix = reinterpret_cast<const int4 *>(plan_ix);
    auto pw = reinterpret_cast<const double2 *>(plan_w);
    if (vec)
        k_gather2d_lf<2><<<grid, 256, 0, (cudaStream_t)stream>>>(
            vp, nvar, (int)layers, s_y, s_x, pix, pw, xloc, yloc, dx, dy, (int)lenx, (int)leny, n, o_n, hgt_idx);
    else
        k_gather2d_lf<1><<<grid, 256, 0, (cudaStream_t)stream>>>(
            vp, nvar, (int)layers, s_y, s_x, pix, pw, xloc, yloc, dx, dy, (int)lenx, (int)leny, n, o_n, hgt_idx);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_interp2d_layers(const double *src, int64_t layers, int64_t leny, int64_t lenx, int64_t s_l,
                       int64_t s_y, int64_t s_x, double dx, double dy, const double *xloc,
                       const double *yloc, int64_t n, double *out, int64_t o_l, int64_t o_n,
                       mt_stream_t stream)
{
    MT_REQUIRE(src && xloc && yloc && out, "mt_interp2d_layers: null pointer");
    MT_REQUIRE(fits_int(layers) && fits_int(lenx) && fits_int(leny) && n >= 0,
               "mt_interp2d_layers: sizes must be in (0, 2^31)");
    if (n == 0) return MT_OK;
    if (s_l == 1 && o_l == 1 && layers >= 16)
        return mt_gather2d_lf(&src, &out, 1, layers, leny, lenx, s_y, s_x, dx, dy, xloc, yloc,
                              nullptr, nullptr, n, o_n, nullptr, stream);
    const int chunk = layers >= 64 ? 16 : (layers >= 8 ? 8 : (int)layers);
    mt::ProfScope prof("k_interp2d_pt", stream);
    dim3 grid(mt::grid_for(n, 256), (unsigned)((layers + chunk - 1) / chunk));
    k_interp2d_pt<<<grid, 256, 0, (cudaStream_t)stream>>>(src, layers, (int)leny, (int)lenx, s_l, s_y,
                                                          s_x, dx, dy, xloc, yloc, n, out, o_l, o_n,
                                                          chunk);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_interp3d_layers(const double *src, int64_t layers, int64_t lenz, int64_t leny, int64_t lenx,
                       int64_t s_l, int64_t s_z, int64_t s_y, int64_t s_x, double dx, double dy,
                       double dz, const double *xloc, const double *yloc, const double *zloc,
                       int64_t n, double *out, int64_t o_l, int64_t o_n, mt_stream_t stream)
{
    MT_REQUIRE(src && xloc && yloc && zloc && out, "mt_interp3d_layers: null pointer");
    MT_REQUIRE(fits_int(layers) && fits_int(lenx) && fits_int(leny) && fits_int(lenz) && n >= 0,
               "mt_interp3d_layers: sizes must be in (0, 2^31)");
    if (n == 0) return MT_OK;
    const unsigned grid = mt::grid_for(layers * n, 256);
    mt::ProfScope prof("k_interp3d", stream);
    if (s_l == 1 && layers > 1)
        k_interp3d<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
            src, layers, (int)lenz, (int)leny, (int)lenx, s_l, s_z, s_y, s_x, dx, dy, dz, xloc, yloc,
            zloc, n, out, o_l, o_n);
    else
        k_interp3d<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
            src, layers, (int)lenz, (int)leny, (int)lenx, s_l, s_z, s_y, s_x, dx, dy, dz, xloc, yloc,
            zloc, n, out, o_l, o_n);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_interp1d_layers(const double *src, int64_t layers, int64_t lenx, int64_t s_l, int64_t s_x,
                       double dx, const double *xloc, int64_t n, double *out, int64_t o_l,
                       int64_t o_n, mt_stream_t stream)
{
    MT_REQUIRE(src && xloc && out, "mt_interp1d_layers: null pointer");
    MT_REQUIRE(fits_int(layers) && fits_int(lenx) && n >= 0, "mt_interp1d_layers: bad sizes");
    if (n == 0) return MT_OK;
    const unsigned grid = mt::grid_for(layers * n, 256);
    mt::ProfScope prof("k_interp1d", stream);
    if (s_l == 1 && layers > 1)
        k_interp1d<true><<<grid, 256, 0, (cudaStream_t)stream>>>(src, layers, (int)lenx, s_l, s_x, dx,
                                                                 xloc, n, out, o_l, o_n);
    else
        k_interp1d<false><<<grid, 256, 0, (cudaStream_t)stream>>>(src, layers, (int)lenx, s_l, s_x, dx,
                                                                  xloc, n, out, o_l, o_n);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_interp1d_nonequal_layers(const double *x_input, const double *src, int64_t layers,
                                int64_t lenx, int64_t s_l, int64_t s_x, const double *x_output,
                                int64_t n, double *out, int64_t o_l, int64_t o_n,
                                mt_stream_t stream)
{
    MT_REQUIRE(x_input && src && x_output && out, "mt_interp1d_nonequal_layers: null pointer");
    MT_REQUIRE(fits_int(layers) && fits_int(lenx) && lenx >= 2 && n >= 0,
               "mt_interp1d_nonequal_layers: bad sizes");
    if (n == 0) return MT_OK;
    const unsigned grid = mt::grid_for(layers * n, 256);
    mt::ProfScope prof("k_interp1d_nonequal", stream);
    if (s_l == 1 && layers > 1)
        k_interp1d_nonequal<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
            x_input, src, layers, (int)lenx, s_l, s_x, x_output, n, out, o_l, o_n);
    else
        k_interp1d_nonequal<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
            x_input, src, layers, (int)lenx, s_l, s_x, x_output, n, out, o_l, o_n);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

}  // extern "C"
