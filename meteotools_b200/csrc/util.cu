// Staging helpers for the host-facing wrappers: exact f32 -> f64 upcast and a dense 3-D
// permutation copy (layout changes between the user's C-order arrays and the library's
// level-fastest "TRZ" order) through a padded shared-memory tile so that both the read and the
// write are coalesced.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_cast_f32_f64(const float *__restrict__ in, double *__restrict__ out, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)__ldg(in + i);
}

// float64 -> float32 (round to nearest, what NumPy's astype and the reference's netCDF cache 'f4'
// write do): results that leave the device as float32 take half the PCIe bytes.
__global__ void __launch_bounds__(256)
k_cast_f64_f32(const double *__restrict__ in, float *__restrict__ out, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)__ldg(in + i);
}

// dst is dense (n0, n1, n2); dst[i,j,k] = src[i*s0 + j*s1 + k*s2].
// Tiles of 32x32 over (the axis that is contiguous in src) x (k); generic fallback otherwise.
__global__ void __launch_bounds__(256)
k_permute3_generic(const double *__restrict__ src, double *__restrict__ dst, int64_t n0, int64_t n1,
                   int64_t n2, int64_t s0, int64_t s1, int64_t s2)
{
    const int64_t total = n0 * n1 * n2;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t % n2, q = t / n2;
        const int64_t j = q % n1, i = q / n1;
        dst[t] = __ldg(src + i * s0 + j * s1 + k * s2);
    }
}

// Case s0 == 1 (src contiguous along dst axis 0), dst contiguous along axis 2: transpose (i,k)
// planes for each j through shared memory.
__global__ void __launch_bounds__(256)
k_permute3_t02(const double *__restrict__ src, double *__restrict__ dst, int64_t n0, int64_t n1,
               int64_t n2, int64_t s1, int64_t s2)
{
    __shared__ double tile[32][33];
    const int64_t tiles_i = (n0 + 31) / 32, tiles_k = (n2 + 31) / 32;
    const int64_t ntiles = tiles_i * tiles_k * n1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int64_t tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const int64_t ti = tile_id % tiles_i, q = tile_id / tiles_i;
        const int64_t tk = q % tiles_k, j = q / tiles_k;
        const int64_t i0 = ti * 32, k0 = tk * 32;
        for (int r = ty; r < 32; r += 8) {  // read: i contiguous
            const int64_t i = i0 + tx, k = k0 + r;
            if (i < n0 && k < n2) tile[r][tx] = __ldg(src + i + j * s1 + k * s2);
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {  // write: k contiguous
            const int64_t i = i0 + r, k = k0 + tx;
            if (i < n0 && k < n2) dst[(i * n1 + j) * n2 + k] = tile[tx][r];
        }
        __syncthreads();
    }
}


// Batched (targets, levels) -> (levels, targets) transpose of the rows mt_setdata_tma writes, for a
// (z, y, x) grid, with the Cartesian terrain mask (z[k] <= hgt[i] -> NaN, cartesian_grid.py:127-136)
// applied on the way: ONE launch for all variables of a set_data call instead of one transpose and
// one masking pass per variable.  A tile = 32 consecutive targets x all levels: contiguous in the
// source (flat coalesced read), 256-byte segments per level in the destination.
struct RowsToLevels {
    const double *rows[MT_MAX_VARS];
    double *out[MT_MAX_VARS];
};
constexpr int RTL_TGT = 32, RTL_MAXLEV = 96;

__global__ void __launch_bounds__(256)
k_rows_to_levels(RowsToLevels P, int nvar, int64_t n, int nlev, const double *__restrict__ z,
                 const double *__restrict__ hgt)
{
    __shared__ double tile[RTL_TGT][RTL_MAXLEV + 1];
    __shared__ double z_s[RTL_MAXLEV];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (z && tid < nlev) z_s[tid] = z[tid];
    const int64_t ntile = (n + RTL_TGT - 1) / RTL_TGT, nwork = ntile * nvar;
    for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int v = (int)(w % nvar);           // the variables of a tile back to back: hgt stays in L1
        const int64_t i0 = (w / nvar) * RTL_TGT;
        const int cnt = (int)min((int64_t)RTL_TGT, n - i0);
        const double *__restrict__ src = P.rows[v] + i0 * nlev;
        __syncthreads();
        for (int e = tid; e < cnt * nlev; e += 256) {
            const int r = e / nlev;
            tile[r][e - r * nlev] = __ldcs(src + e);
        }
        __syncthreads();
        if (lane < cnt) {
            const double h = z ? __ldg(hgt + i0 + lane) : 0.0;
            double *__restrict__ dst = P.out[v] + i0 + lane;
            for (int k = warp; k < nlev; k += 8) {
                double val = tile[lane][k];
                if (z && z_s[k] <= h) val = mt::nan_f64();
                __stcs(dst + (int64_t)k * n, val);
            }
        }
    }
}
// wrfout_grid.correct_map_scale (grids/wrfout_grid.py:48-55): ua *= MAPFAC_MX ... -- a (z, y, x) field times
// a (y, x) factor IN PLACE, in NumPy's arithmetic for the two dtypes: float32 *= float32 is a float
// product, every other pairing is the double product cast back to the field's type.
template <typename TF, typename TM>
__global__ void __launch_bounds__(256)
k_scale_box_2d(TF *__restrict__ box, int64_t nz, int by, int bx, int64_t pitch, const TM *__restrict__ fac,
               int64_t fac_pitch, int64_t org_y, int64_t org_x)
{
    const int64_t total = nz * by * (int64_t)bx;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % bx);
        const int64_t q = e / bx;
        const int j = (int)(q % by);
        const int64_t k = q / by;
        TF *ptr = box + (k * by + j) * pitch + i;
        const TM m = __ldg(fac + (org_y + j) * fac_pitch + org_x + i);
        if (sizeof(TF) == 4 && sizeof(TM) == 4) *ptr = (TF)((float)*ptr * (float)m);
        else *ptr = (TF)((double)*ptr * (double)m);
    }
}
}  // namespace

extern "C" {

int mt_scale_box_2d(void *box, int dtype, int64_t nz, int64_t by, int64_t bx, int64_t row_pitch, const void *fac,
                    int fac_dtype, int64_t fac_pitch, int64_t org_y, int64_t org_x, mt_stream_t stream)
{
    MT_REQUIRE(box && fac && nz >= 1 && by >= 1 && bx >= 1 && row_pitch >= bx && fac_pitch >= 1 && org_y >= 0 &&
                   org_x >= 0 && by < 2147483647LL && bx < 2147483647LL,
               "mt_scale_box_2d: bad arguments");
    MT_REQUIRE((dtype == MT_F64 || dtype == MT_F32) && (fac_dtype == MT_F64 || fac_dtype == MT_F32),
               "mt_scale_box_2d: bad dtype");
    const unsigned grid = mt::grid_for(nz * by * bx, 256, 8);
    cudaStream_t s = (cudaStream_t)stream;
    mt::ProfScope prof("k_scale_box_2d", stream);
#define SC_LAUNCH(TF, TM)                                                                                        \
    k_scale_box_2d<TF, TM><<<grid, 256, 0, s>>>((TF *)box, nz, (int)by, (int)bx, row_pitch, (const TM *)fac, fac_pitch, \
                                                org_y, org_x)
    if (dtype == MT_F32 && fac_dtype == MT_F32) SC_LAUNCH(float, float);
    else if (dtype == MT_F32) SC_LAUNCH(float, double);
    else if (fac_dtype == MT_F32) SC_LAUNCH(double, float);
    else SC_LAUNCH(double, double);
#undef SC_LAUNCH
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_cast_f32_f64(const float *in, double *out, int64_t n, mt_stream_t stream)
{
    MT_REQUIRE(in && out && n >= 0, "mt_cast_f32_f64: bad arguments");
    if (n == 0) return MT_OK;
    mt::ProfScope prof("k_cast_f32_f64", stream);
    k_cast_f32_f64<<<mt::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(in, out, n);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_cast_f64_f32(const double *in, float *out, int64_t n, mt_stream_t stream)
{
    MT_REQUIRE(in && out && n >= 0, "mt_cast_f64_f32: bad arguments");
    if (n == 0) return MT_OK;
    mt::ProfScope prof("k_cast_f64_f32", stream);
    k_cast_f64_f32<<<mt::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(in, out, n);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_rows_to_levels(const double *const *rows, double *const *out, int nvar, int64_t n, int64_t nlev,
                      const double *z, const double *hgt, mt_stream_t stream)
{
    MT_REQUIRE(rows && out && nvar >= 1 && nvar <= MT_MAX_VARS && n > 0 && nlev >= 1 && nlev <= RTL_MAXLEV,
               "mt_rows_to_levels: need 1..%d variables and 1..%d levels", MT_MAX_VARS, RTL_MAXLEV);
    MT_REQUIRE((z == nullptr) == (hgt == nullptr), "mt_rows_to_levels: z and hgt go together");
    RowsToLevels P;
    for (int v = 0; v < nvar; ++v) {
        MT_REQUIRE(rows[v] && out[v] && rows[v] != out[v], "mt_rows_to_levels: variable %d: null or in place", v);
        P.rows[v] = rows[v], P.out[v] = out[v];
    }
    const int64_t nwork = (n + RTL_TGT - 1) / RTL_TGT * nvar;
    mt::ProfScope prof("k_rows_to_levels", stream);
    k_rows_to_levels<<<mt::grid_for(nwork, 1, 8), 256, 0, (cudaStream_t)stream>>>(P, nvar, n, (int)nlev, z, hgt);
    MT_LAUNCH_CHECK();
    return MT_OK;
}

int mt_permute3(const double *src, double *dst, int64_t n0, int64_t n1, int64_t n2, int64_t s0,
                int64_t s1, int64_t s2, mt_stream_t stream)
{
    MT_REQUIRE(src && dst && src != dst && n0 > 0 && n1 > 0 && n2 > 0, "mt_permute3: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    mt::ProfScope prof("k_permute3", stream);
    if (s0 == 1 && s2 != 1 && n0 >= 8 && n2 >= 8) {
        const int64_t ntiles = ((n0 + 31) / 32) * ((n2 + 31) / 32) * n1;
        k_permute3_t02<<<mt::grid_for(ntiles, 1, 16), 256, 0, s>>>(src, dst, n0, n1, n2, s1, s2);
    } else {
        k_permute3_generic<<<mt::grid_for(n0 * n1 * n2, 256, 8), 256, 0, s>>>(src, dst, n0, n1, n2, s0, s1, s2);
    }
    MT_LAUNCH_CHECK();
    return MT_OK;
}

}  // extern "C"

// (The source may also be a DEVICE array: set_data stages the box of a resident source whose rows are
// not 16-byte aligned -- an odd nx of float64 -- so that the TMA kernel can read it.)
// Strided host->device upload of the column box [y0,y1)x[x0,x1) of a C-order (nz,ny,nx) HOST
// array into a dense (nz, y1-y0, x1-x0) device array: one cudaMemcpy3DAsync (DMA straight from
// the caller's buffer -- at PCIe speed when it is pinned), no host-side crop copy.
extern "C" int mt_upload_box(const void *host, int elem_bytes, int64_t nz, int64_t ny, int64_t nx,
                             int64_t y0, int64_t y1, int64_t x0, int64_t x1, void *dev,
                             mt_stream_t stream)
{
    MT_REQUIRE(host && dev && (elem_bytes == 4 || elem_bytes == 8), "mt_upload_box: bad arguments");
    MT_REQUIRE(nz > 0 && 0 <= y0 && y0 < y1 && y1 <= ny && 0 <= x0 && x0 < x1 && x1 <= nx,
               "mt_upload_box: box outside the array");
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(host), (size_t)nx * elem_bytes, (size_t)nx * elem_bytes, (size_t)ny);
    p.srcPos = make_cudaPos((size_t)x0 * elem_bytes, (size_t)y0, 0);
    p.dstPtr = make_cudaPitchedPtr(dev, (size_t)(x1 - x0) * elem_bytes, (size_t)(x1 - x0) * elem_bytes, (size_t)(y1 - y0));
    p.dstPos = make_cudaPos(0, 0, 0);
    p.extent = make_cudaExtent((size_t)(x1 - x0) * elem_bytes, (size_t)(y1 - y0), (size_t)nz);
    p.kind = cudaMemcpyDefault;   // host (pinned or pageable) or device source: decided from the pointer
    MT_CUDA(cudaMemcpy3DAsync(&p, (cudaStream_t)stream));
    return MT_OK;
}

extern "C" int mt_upload_box_pitched(const void *host, int elem_bytes, int64_t nz, int64_t ny, int64_t nx,
                                     int64_t y0, int64_t y1, int64_t x0, int64_t x1, void *dev,
                                     int64_t dev_row_pitch, mt_stream_t stream)
{
    MT_REQUIRE(host && dev && (elem_bytes == 4 || elem_bytes == 8), "mt_upload_box_pitched: bad arguments");
    MT_REQUIRE(nz > 0 && 0 <= y0 && y0 < y1 && y1 <= ny && 0 <= x0 && x0 < x1 && x1 <= nx &&
                   dev_row_pitch >= x1 - x0,
               "mt_upload_box_pitched: box outside the array or pitch too small");
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(host), (size_t)nx * elem_bytes, (size_t)nx * elem_bytes, (size_t)ny);
    p.srcPos = make_cudaPos((size_t)x0 * elem_bytes, (size_t)y0, 0);
    p.dstPtr = make_cudaPitchedPtr(dev, (size_t)dev_row_pitch * elem_bytes, (size_t)dev_row_pitch * elem_bytes,
                                   (size_t)(y1 - y0));
    p.dstPos = make_cudaPos(0, 0, 0);
    p.extent = make_cudaExtent((size_t)(x1 - x0) * elem_bytes, (size_t)(y1 - y0), (size_t)nz);
    p.kind = cudaMemcpyDefault;   // host (pinned or pageable) or device source: decided from the pointer
    MT_CUDA(cudaMemcpy3DAsync(&p, (cudaStream_t)stream));
    return MT_OK;
}
