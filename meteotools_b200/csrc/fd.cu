// Generic finite differences along one axis of a dense array (SURVEY 8a rows a8, a9;
// calc/differential.py:6-180 through calc/check_arrays.py:52-63).  The array is viewed as
// (outer, n, inner); one thread per element, neighbours come through L1/L2 (stride `inner`).
#include "fd.cuh"

namespace {

template <int KIND>
__global__ void __launch_bounds__(256)
k_fd(const double *__restrict__ in, double *__restrict__ out, int64_t outer, int n, int64_t inner,
     double delta)
{
    const int64_t total = outer * n * inner;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t % inner;
        const int64_t oi = t / inner;
        const int i = (int)(oi % n);
        const int64_t o = oi / n;
        const double *base = in + o * n * inner + j;
        auto v = [&](int k) { return __ldg(base + (int64_t)k * inner); };
        double r;
        if (KIND == MT_FD2) r = mt::fd2(v, i, n, delta);
        else if (KIND == MT_FD4) r = mt::fd4(v, i, n, delta);
        else if (KIND == MT_FD2_FRONT) r = mt::fd2_front(v, i, n, delta);
        else if (KIND == MT_FD2_BACK) r = mt::fd2_back(v, i, n, delta);
        else if (KIND == MT_FD2_2) r = mt::fd2_2(v, i, n, delta);
        else if (KIND == MT_FD2_2_FRONT) r = mt::fd2_2_front(v, i, n, delta);
        else r = mt::fd2_2_back(v, i, n, delta);
        mt::st_stream(out + t, r);
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
    const unsigned grid = mt::grid_for(outer * n * inner, 256, 16);
    cudaStream_t s = (cudaStream_t)stream;
    mt::ProfScope prof("k_fd", stream);
    switch (kind) {
#define FD_CASE(K) \
    case K: k_fd<K><<<grid, 256, 0, s>>>(in, out, outer, (int)n, inner, delta); break;
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
    MT_LAUNCH_CHECK();
    return MT_OK;
}
