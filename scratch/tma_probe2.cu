// usage: tma_probe2 b0 b1 b2   (3-D f64 tensor map of a (14,70,84) array, box {b0,b1,b2})
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void kern(const __grid_constant__ CUtensorMap M, double* outp, int nelem, int x, int y, int rank)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((nelem*8 + 127) & ~127));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nelem*8) : "memory");
        if (rank == 2)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(smem_u32(smem)), "l"(&M), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
        else
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_u32(smem)), "l"(&M), "r"(x), "r"(y), "r"(0), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    } while (!ok);
    const double* t = reinterpret_cast<const double*>(smem);
    for (int i = threadIdx.x; i < nelem; i += blockDim.x) outp[i] = t[i];
}
int main(int argc, char** argv)
{
    const int b0 = atoi(argv[1]), b1 = atoi(argv[2]), b2 = atoi(argv[3]);
    const int rank = argc > 4 ? atoi(argv[4]) : 3, f32 = argc > 5 ? atoi(argv[5]) : 0, promo = argc > 6 ? atoi(argv[6]) : 1;
    const int nz = 14, ny = 70, nx = 84;
    std::vector<double> h((size_t)nz*ny*nx);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) h[((size_t)k*ny+j)*nx+i] = k*10000 + j*100 + i;
    const int nelem = f32 ? b0*b1*b2/2 : b0*b1*b2;
    double *d, *out; cudaMalloc(&d, h.size()*8); cudaMalloc(&out, nelem*8);
    cudaMemcpy(d, h.data(), h.size()*8, cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fp;
    CUtensorMap M;
    cuuint64_t gdim[3] = {(cuuint64_t)(f32 ? 2*nx : nx), (cuuint64_t)(rank == 2 ? ny*nz : ny), nz}; cuuint64_t gstr[2] = {(cuuint64_t)nx*8, (cuuint64_t)nx*ny*8};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2}; cuuint32_t es[3] = {1,1,1};
    CUresult r = enc(&M, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024);
    int x = argc > 7 ? atoi(argv[7]) : 7, y = 3;
    kern<<<1, 128, nelem*8 + 512, 0>>>(M, out, nelem, f32 ? 2*x : x, y, rank);
    cudaError_t e = cudaDeviceSynchronize();
    printf("box {%d,%d,%d} rank %d f32 %d promo %d encode %d: %s", b0, b1, b2, rank, f32, promo, (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    if (rank == 2 || f32) { printf(" (no check)\n"); return 0; }
    std::vector<double> o(nelem); cudaMemcpy(o.data(), out, o.size()*8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int kk = 0; kk < b2; ++kk) for (int j = 0; j < b1; ++j) for (int i = 0; i < b0; ++i) {
        double want = (x+i < nx && y+j < ny && kk < nz) ? kk*10000 + (y+j)*100 + (x+i) : 0.0;
        if (o[(kk*b1+j)*b0+i] != want) ++bad;
    }
    printf(" mismatches %d\n", bad);
    return 0;
}
