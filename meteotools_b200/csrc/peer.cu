// Halo planes over NVLink / NVSwitch PEER MEMORY (the fast path next to csrc/halo.cu's NCCL exchange).
//
// NCCL's send / recv kernels need SMs and move the three 32 MB planes of a 2000 x 2000 level shard at
// ~160 GB/s per direction; launched beside the persistent compute kernels they are starved and the "overlap"
// is lost (profiles/experiments_r02.md: 3.0 ms per step at 8 GPUs against 2.1 ms for the kernels alone).
// Here the planes travel by the copy engines, straight into a mailbox in the NEIGHBOUR's memory, and the two
// processes order themselves with stream memory operations -- no kernel, no SM, no host synchronisation:
//
//   box (one per neighbour, allocated by its owner, opened by the neighbour through CUDA IPC):
//       [0]   data_seq   written by the NEIGHBOUR after its planes have landed in `data`
//       [64]  ack_seq    written by the NEIGHBOUR after it has consumed the planes I put into ITS box
//       [256] data       the planes
//   step s, per neighbour:   wait(own box.ack_seq >= s - 1)  ->  cudaMemcpyAsync(planes -> neighbour's box.data)
//                            ->  write(neighbour's box.data_seq = s)
//                            wait(own box.data_seq >= s)  ->  copy own box.data into the halo levels
//                            ->  write(neighbour's box.ack_seq = s)
//   (cuStreamWaitValue32 / cuStreamWriteValue32: every flag a rank WAITS on lives in its own memory.)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>

#include "common.cuh"

namespace {

typedef CUresult (*WaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*WriteFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

void *driver_fn(const char *name)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    return p;
}
WaitFn wait_fn()
{
    static WaitFn f = (WaitFn)driver_fn("cuStreamWaitValue32");
    return f;
}
WriteFn write_fn()
{
    static WriteFn f = (WriteFn)driver_fn("cuStreamWriteValue32");
    return f;
}

}  // namespace

extern "C" {

int mt_peer_supported(void) { return wait_fn() && write_fn() ? 1 : 0; }

int mt_peer_box_create(int64_t bytes, void **box, void *ipc_handle64)
{
    MT_REQUIRE(bytes > 0 && box && ipc_handle64, "mt_peer_box_create: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC memory handles are 64 bytes");
    void *p = nullptr;
    MT_CUDA(cudaMalloc(&p, (size_t)bytes));
    MT_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return mt::cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
    }
    memcpy(ipc_handle64, &h, sizeof(h));
    *box = p;
    return MT_OK;
}

int mt_peer_box_open(const void *ipc_handle64, void **peer_ptr)
{
    MT_REQUIRE(ipc_handle64 && peer_ptr, "mt_peer_box_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, sizeof(h));
    void *p = nullptr;
    MT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr = p;
    return MT_OK;
}

int mt_peer_box_close(void *peer_ptr)
{
    if (peer_ptr) MT_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return MT_OK;
}

int mt_peer_box_destroy(void *box)
{
    if (box) MT_CUDA(cudaFree(box));
    return MT_OK;
}

int mt_peer_copy(void *dst, const void *src, int64_t bytes, mt_stream_t stream)
{
    MT_REQUIRE(dst && src && bytes > 0, "mt_peer_copy: bad arguments");
    MT_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return MT_OK;
}

int mt_stream_write32(void *addr, uint32_t value, mt_stream_t stream)
{
    MT_REQUIRE(addr && ((uintptr_t)addr & 3) == 0, "mt_stream_write32: bad address");
    WriteFn f = write_fn();
    MT_REQUIRE(f, "mt_stream_write32: cuStreamWriteValue32 is not available in this driver");
    const CUresult r = f((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) {
        mt::set_error("cuStreamWriteValue32 failed (%d)", (int)r);
        return MT_ECUDA;
    }
    return MT_OK;
}

int mt_stream_wait_geq32(void *addr, uint32_t value, mt_stream_t stream)
{
    MT_REQUIRE(addr && ((uintptr_t)addr & 3) == 0, "mt_stream_wait_geq32: bad address");
    WaitFn f = wait_fn();
    MT_REQUIRE(f, "mt_stream_wait_geq32: cuStreamWaitValue32 is not available in this driver");
    const CUresult r = f((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) {
        mt::set_error("cuStreamWaitValue32 failed (%d)", (int)r);
        return MT_ECUDA;
    }
    return MT_OK;
}

}  // extern "C"
