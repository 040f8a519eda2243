// micro-benchmark: what limits a streaming FD2 stencil on B200?   (80,1000,1000) f64
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int MODE, int CHUNK>   // MODE 0: IEEE divide, st.cs   1: multiply by reciprocal   2: IEEE divide, plain store  3: copy only
__global__ void __launch_bounds__(256) k0(const double* __restrict__ in, double* __restrict__ out, int n, int64_t inner, double delta)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= inner) return;
    const int i0 = blockIdx.y * CHUNK, i1 = min(i0 + CHUNK, n);
    const double rd = 1.0 / delta;
    #pragma unroll 4
    for (int i = i0; i < i1; ++i) {
        double r;
        if (MODE == 3) r = __ldg(in + (int64_t)i * inner + j);
        else {
            double a, b;
            if (i == 0) { a = -3 * __ldg(in + j) + 4 * __ldg(in + inner + j); b = __ldg(in + 2 * inner + j); a = a - b; }
            else if (i == n - 1) { a = 3 * __ldg(in + (int64_t)(n - 1) * inner + j) - 4 * __ldg(in + (int64_t)(n - 2) * inner + j) + __ldg(in + (int64_t)(n - 3) * inner + j); }
            else a = __ldg(in + (int64_t)(i + 1) * inner + j) - __ldg(in + (int64_t)(i - 1) * inner + j);
            r = MODE == 1 ? a * 0.5 * rd : a / 2 / delta;
        }
        if (MODE == 2) out[(int64_t)i * inner + j] = r; else __stcs(out + (int64_t)i * inner + j, r);
    }
}
// rolling-register march: one load per point
template <int CHUNK>
__global__ void __launch_bounds__(256) k_roll(const double* __restrict__ in, double* __restrict__ out, int n, int64_t inner, double delta)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= inner) return;
    const int i0 = max(blockIdx.y * CHUNK, 1), i1 = min(blockIdx.y * CHUNK + CHUNK, n - 1);
    double prev = __ldg(in + (int64_t)(i0 - 1) * inner + j), cur = __ldg(in + (int64_t)i0 * inner + j);
    #pragma unroll 4
    for (int i = i0; i < i1; ++i) {
        const double nxt = __ldg(in + (int64_t)(i + 1) * inner + j);
        __stcs(out + (int64_t)i * inner + j, (nxt - prev) / 2 / delta);
        prev = cur; cur = nxt;
    }
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a); for (int i = 0; i < 20; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20;
}
int main()
{
    const int n0 = 80; const int64_t inner = 1000 * 1000; const int64_t tot = n0 * inner;
    double *in, *out; cudaMalloc(&in, tot * 8); cudaMalloc(&out, tot * 8); cudaMemset(in, 0, tot * 8);
    const double gb = 16.0 * tot / 1e9;
#define RUN(NAME, KERNEL, CH) { dim3 g((unsigned)((inner + 255) / 256), (n0 + CH - 1) / CH); float ms = timeit([&] { KERNEL<<<g, 256>>>(in, out, n0, inner, 250.0); }); printf("%-34s chunk %3d: %.4f ms  %.0f GB/s  %s\n", NAME, CH, ms, gb / ms * 1e3 / 1, cudaGetErrorString(cudaGetLastError())); }
    RUN("ieee divide, st.cs", (k0<0, 16>), 16)
    RUN("ieee divide, st.cs", (k0<0, 4>), 4)
    RUN("ieee divide, st.cs", (k0<0, 80>), 80)
    RUN("reciprocal multiply, st.cs", (k0<1, 16>), 16)
    RUN("ieee divide, plain store", (k0<2, 16>), 16)
    RUN("copy only", (k0<3, 16>), 16)
    RUN("copy only", (k0<3, 1>), 1)
    RUN("rolling registers, ieee divide", (k_roll<16>), 16)
    RUN("rolling registers, ieee divide", (k_roll<80>), 80)
    float ms = timeit([&] { cudaMemcpyAsync(out, in, tot * 8, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D: %.4f ms %.0f GB/s\n", ms, gb / ms * 1e3);
    return 0;
}
