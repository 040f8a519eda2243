// probe: mbarrier / 1-D bulk / 3-D FP64 TMA box load, checked on the host.  usage: tma_probe MODE
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void kern(const __grid_constant__ CUtensorMap M, const double* src, double* outp, const CUtensorMap* gmap, int mode, int nz, int x, int y)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 80*nz*8);
    if (mode == 0) { for (int i = threadIdx.x; i < 80*nz; i += blockDim.x) outp[i] = 1.0; return; }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mode == 1) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        } else if (mode == 2) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(1024) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(smem_u32(smem)), "l"(src), "r"(1024), "r"(smem_u32(bar)) : "memory");
        } else if (mode == 8) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(32*8*4) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(smem_u32(smem)), "l"(gmap), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
        } else {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(80*nz*8) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(smem)), "l"(mode >= 7 ? gmap : &M), "r"((mode == 5 || mode == 6) ? 2*x : x), "r"(y), "r"(0), "r"(smem_u32(bar)) : "memory");
        }
    }
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    } while (!ok);
    const double* t = reinterpret_cast<const double*>(smem);
    for (int i = threadIdx.x; i < 80*nz; i += blockDim.x) outp[i] = t[i];
}
int main(int argc, char** argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 3;
    const int nz = 14, ny = 70, nx = 84;
    std::vector<double> h((size_t)nz*ny*nx);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) h[((size_t)k*ny+j)*nx+i] = k*10000 + j*100 + i;
    double *d, *out; cudaMalloc(&d, h.size()*8); cudaMalloc(&out, 80*nz*8);
    cudaMemcpy(d, h.data(), h.size()*8, cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fp;
    CUtensorMap M;
    cuuint64_t gdim[3] = {nx, ny, nz}; cuuint64_t gstr[2] = {(cuuint64_t)nx*8, (cuuint64_t)nx*ny*8};
    cuuint32_t box[3] = {16, 5, nz}; cuuint32_t es[3] = {1,1,1};
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    if (mode == 4) dt = CU_TENSOR_MAP_DATA_TYPE_UINT64;
    if (mode == 5) { dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; gdim[0] = 2*nx; box[0] = 32; }
    if (mode == 6) { dt = CU_TENSOR_MAP_DATA_TYPE_INT32; gdim[0] = 2*nx; box[0] = 32; }
    int rank = 3;
    if (mode == 8) { dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; rank = 2; gdim[0] = 2*nx; gdim[1] = ny*nz; box[0] = 32; box[1] = 8; }
    CUresult r = enc(&M, dt, rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("mode %d entry %d %d encode %d\n", mode, (int)e, (int)q, (int)r);
    CUtensorMap* gmap; cudaMalloc(&gmap, 128); cudaMemcpy(gmap, &M, 128, cudaMemcpyHostToDevice);
    { const unsigned* w = (const unsigned*)&M; for (int i = 0; i < 32; ++i) printf("%08x%c", w[i], i%8==7?'\n':' '); }
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100*1024);
    for (int trial = 0; trial < 2; ++trial) {
        int x = trial == 0 ? 7 : 75, y = trial == 0 ? 3 : 68;
        if (mode == 9) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 80*nz*8 + 256;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, kern, M, (const double*)d, out, (const CUtensorMap*)gmap, 3, nz, x, y);
        } else
        kern<<<1, 128, 80*nz*8 + 256, 0>>>(M, d, out, gmap, mode, nz, x, y);
        e = cudaDeviceSynchronize();
        printf("trial %d: %s\n", trial, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        if (mode < 3 || mode == 8) continue;
        std::vector<double> o(80*nz); cudaMemcpy(o.data(), out, o.size()*8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int kk = 0; kk < nz; ++kk) for (int j = 0; j < 5; ++j) for (int i = 0; i < 16; ++i) {
            double want = (x+i < nx && y+j < ny) ? kk*10000 + (y+j)*100 + (x+i) : 0.0;
            if (o[(kk*5+j)*16+i] != want) { if (bad < 5) printf("  mismatch k%d j%d i%d got %g want %g\n", kk,j,i,o[(kk*5+j)*16+i],want); ++bad; }
        }
        printf("trial %d mismatches: %d\n", trial, bad);
    }
    return 0;
}
