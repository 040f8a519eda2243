// Generic finite differences along one axis of a dense array (SURVEY 8a rows a8, a9;
// calc/differential.py:6-180 through calc/check_arrays.py:52-63).  The array is viewed as
// (outer, n, inner).  No integer division per element (the first version spent more issue slots on
// two 64-bit div/mod pairs than on the stencil and ran at 51 % of the HBM peak):
//   inner > 1 : blockIdx.x/threadIdx.x walk the contiguous `inner` index, blockIdx.y a chunk of the
//               differentiated axis, blockIdx.z (+ a stride loop) the outer index; neighbours are
//               re-read through L1/L2 (same rows, just loaded by this block);
//   inner == 1: the differentiated axis itself is contiguous: threads walk it, blockIdx.y the rows.
#include <cstdlib>

#include "fd.cuh"

namespace {

template <int KIND, class V>
__device__ __forceinline__ double fd_apply(V v, int i, int n, const mt::CDiv &delta)
{
    if (KIND == MT_FD2) return mt::fd2(v, i, n, delta);
    if (KIND == MT_FD4) return mt::fd4(v, i, n, delta);
    if (KIND == MT_FD2_FRONT) return mt::fd2_front(v, i, n, delta);
    if (KIND == MT_FD2_BACK) return mt::fd2_back(v, i, n, delta);
    if (KIND == MT_FD2_2) return mt::fd2_2(v, i, n, delta);
    if (KIND == MT_FD2_2_FRONT) return mt::fd2_2_front(v, i, n, delta);
    return mt::fd2_2_back(v, i, n, delta);
}

// Points of the differentiated axis per thread (inner > 1) and resident CTAs per SM.  Eight points at 40
// registers: with sixteen (round 1) the eighteen operands did not fit the 80 registers of three resident CTAs and
// ptxas kept them in LOCAL memory (168 bytes of stack): 0.33 ms; 8 / 6 CTAs: 0.22 / 0.21 ms along axes 0 / 1 of
// (80, 1000, 1000) (profiles/experiments_r02.md section 3).  -DMT_FD_CHUNK / -DMT_FD_MINB for A/B builds.
#ifndef MT_FD_CHUNK
#define MT_FD_CHUNK 8
#endif
#ifndef MT_FD_MINB
#define MT_FD_MINB 6
#endif
constexpr int FD_CHUNK = MT_FD_CHUNK;

template <int KIND>
__global__ void __launch_bounds__(256, MT_FD_MINB)
k_fd(const double *__restrict__ in, double *__restrict__ out, int64_t outer, int n, int64_t inner,
     double delta)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= inner) return;
    const int i0 = blockIdx.y * FD_CHUNK, i1 = min(i0 + FD_CHUNK, n);
    const mt::CDiv dd(delta);
    for (int64_t o = blockIdx.z; o < outer; o += gridDim.z) {
        const double *base = in + o * n * inner + j;
        double *ob = out + o * n * inner + j;
        auto v = [&](int k) { return __ldg(base + (int64_t)k * inner); };
        if (KIND == MT_FD2) {
            // the hot kind: the chunk's FD_CHUNK + 2 operands are loaded FIRST -- ten independent loads in flight
            // per thread (a rolling prev / cur / next loop keeps one) -- then differenced from registers; the two
            // halo operands of a chunk are the neighbouring chunks' own, i.e. L2 hits
            double val[FD_CHUNK + 2];
#pragma unroll
            for (int k = 0; k < FD_CHUNK + 2; ++k) {
                const int idx = i0 - 1 + k;
                val[k] = (idx >= 0 && idx < n && idx <= i1) ? v(idx) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < FD_CHUNK; ++k) {
                const int i = i0 + k;
                if (i < i1) {
                    const double r = (i == 0 || i == n - 1) ? mt::fd2(v, i, n, dd) : dd((val[k + 2] - val[k]) / 2);
                    mt::st_stream(ob + (int64_t)i * inner, r);
                }
            }
        } else {
#pragma unroll 4
            for (int i = i0; i < i1; ++i) mt::st_stream(ob + (int64_t)i * inner, fd_apply<KIND>(v, i, n, dd));
        }
    }
}

#ifndef MT_FD_FLAT_U
#define MT_FD_FLAT_U 6
#endif
#ifndef MT_FD_FLAT_MINB
#define MT_FD_FLAT_MINB 4
#endif
template <int KIND>
__global__ void __launch_bounds__(256)
k_fd_row(const double *__restrict__ in, double *__restrict__ out, int64_t rows, int n, double delta, double rdelta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const mt::CDiv dd(delta, rdelta);   // 1 / delta comes from the host: one point per thread, no division per thread
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {   // (several rows per trip measured slower: 0.41 vs 0.29 ms)
        const double *base = in + r * n;
        auto v = [&](int k) { return __ldg(base + k); };
        mt::st_stream(out + r * n + i, fd_apply<KIND>(v, i, n, dd));
    }
}

// The contiguous axis in FLAT order: thread <-> element of the whole array, so that every warp stores 32
// consecutive doubles starting on a 256-byte boundary whatever the row length (rows of 1000 doubles start at
// multiples of 64 bytes; a warp store that straddles a 256-byte block costs ~20 % of the write bandwidth on
// B200: profiles/experiments_r02.md section 2).  The row of an element comes from one exact double-precision
// quotient: floor((idx + 0.5) / n) has no rounding hazard for idx < 2^32.
template <int KIND>
__global__ void __launch_bounds__(256, MT_FD_FLAT_MINB)
k_fd_row_flat(const double *__restrict__ in, double *__restrict__ out, uint32_t total, int n, double rn, double delta,
              double rdelta)
{
    // U elements per thread and trip, a grid stride apart: their neighbour loads are issued together (one element
    // per trip keeps 16 bytes per thread in flight: latency bound at full occupancy, ncu long_scoreboard 32 per issue)
    constexpr int U = MT_FD_FLAT_U;
    const mt::CDiv dd(delta, rdelta);
    const uint64_t stride = (uint64_t)gridDim.x * 256u;
    for (uint64_t base = blockIdx.x * 256u + threadIdx.x; base < total; base += U * stride) {
        uint32_t row[U];
        int i[U];
        double hi[U], lo[U];
        bool act[U], inner[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const uint64_t idx = base + k * stride;
            act[k] = idx < total;
            const uint32_t ix = act[k] ? (uint32_t)idx : 0u;
            row[k] = (uint32_t)__double2uint_rz(((double)ix + 0.5) * rn);
            i[k] = (int)(ix - row[k] * (uint32_t)n);
            inner[k] = KIND == MT_FD2 && act[k] && i[k] > 0 && i[k] < n - 1;
            hi[k] = inner[k] ? __ldg(in + ix + 1) : 0.0;
            lo[k] = inner[k] ? __ldg(in + ix - 1) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (!act[k]) continue;
            const uint32_t ix = (uint32_t)(base + k * stride);
            double r;
            if (inner[k]) {
                r = dd((hi[k] - lo[k]) / 2);   // the interior expression of mt::fd2
            } else {
                const double *b = in + (size_t)row[k] * n;
                auto v = [&](int j) { return __ldg(b + j); };
                r = fd_apply<KIND>(v, i[k], n, dd);
            }
            mt::st_stream(out + ix, r);
        }
    }
}

}  // namespace

extern "C" int mt_fd(const double *in, double *out, int64_t outer, int64_t n, int64_t inner,
                     double delta, int kind, mt_stream_t stream)
{
    MT_REQUIRE(in && out && in != out, "mt_fd: null or aliased pointers");
    MT_REQUIRE(outer > 0 && inner > 0 && n > 0 && n < 2147483647LL, "mt_fd: bad extents");
    const int need = (kind == MT_FD4) ? 5 : (kind == MT_FD2_2 || kind == MT_FD2_2_FRONT || kind == MT_FD2_2_BACK) ? 4 : 3;
    MT_REQUIRE(n >= need, "mt_fd: axis length %lld < %d", (long long)n, need);
    cudaStream_t s = (cudaStream_t)stream;
    mt::ProfScope prof("k_fd", stream);
    static const bool flat_rows = [] {   // MT_FD_FLAT=0: the row-tiled kernel (A/B runs)
        const char *e = getenv("MT_FD_FLAT");
        return !(e && e[0] == '0');
    }();
    if (inner == 1 && flat_rows && outer * n < (int64_t)0xffffffffLL && n % 32 != 0) {
        // rows that do not start on 256-byte boundaries: flat order (rows that do are aligned either way)
        const uint32_t total = (uint32_t)(outer * n);
        const unsigned grid = (unsigned)mt::grid_for((int64_t)total, 256, 8);
        switch (kind) {
#define FD_CASE(K) \
    case K: k_fd_row_flat<K><<<grid, 256, 0, s>>>(in, out, total, (int)n, 1.0 / (double)n, delta, 1.0 / delta); break;
            FD_CASE(MT_FD2)
            FD_CASE(MT_FD4)
            FD_CASE(MT_FD2_FRONT)
            FD_CASE(MT_FD2_BACK)
            FD_CASE(MT_FD2_2)
            FD_CASE(MT_FD2_2_FRONT)
            FD_CASE(MT_FD2_2_BACK)
#undef FD_CASE
            default: mt::set_error("mt_fd: unknown kind %d", kind); return MT_EINVAL;
        }
    } else if (inner == 1) {
        const dim3 grid((unsigned)((n + 255) / 256), (unsigned)(outer < 32768 ? outer : 32768));
        switch (kind) {
#define FD_CASE(K) \
    case K: k_fd_row<K><<<grid, 256, 0, s>>>(in, out, outer, (int)n, delta, 1.0 / delta); break;
            FD_CASE(MT_FD2)
            FD_CASE(MT_FD4)
            FD_CASE(MT_FD2_FRONT)
            FD_CASE(MT_FD2_BACK)
            FD_CASE(MT_FD2_2)
            FD_CASE(MT_FD2_2_FRONT)
            FD_CASE(MT_FD2_2_BACK)
#undef FD_CASE
            default: mt::set_error("mt_fd: unknown kind %d", kind); return MT_EINVAL;
        }
    } else {
        const int threads = inner >= 256 ? 256 : (inner >= 128 ? 128 : (inner >= 64 ? 64 : 32));
        const int64_t gx = (inner + threads - 1) / threads;
        const dim3 grid((unsigned)gx, (unsigned)((n + FD_CHUNK - 1) / FD_CHUNK), (unsigned)(outer < 4096 ? outer : 4096));
        MT_REQUIRE(grid.y <= 65535u, "mt_fd: axis too long for one launch (n = %lld)", (long long)n);
        switch (kind) {
#define FD_CASE(K) \
    case K: k_fd<K><<<grid, threads, 0, s>>>(in, out, outer, (int)n, inner, delta); break;
            FD_CASE(MT_FD2)
            FD_CASE(MT_FD4)
            FD_CASE(MT_FD2_FRONT)
            FD_CASE(MT_FD2_BACK)
            FD_CASE(MT_FD2_2)
            FD_CASE(MT_FD2_2_FRONT)
            FD_CASE(MT_FD2_2_BACK)
#undef FD_CASE
            default: mt::set_error("mt_fd: unknown kind %d", kind); return MT_EINVAL;
        }
    }
    MT_LAUNCH_CHECK();
    return MT_OK;
}
