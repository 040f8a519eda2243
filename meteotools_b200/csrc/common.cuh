// Shared helpers for the meteotools_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "meteotools_b200.h"

#define MT_SENTINEL 1.7e308

namespace mt {

// thread-local error text + status plumbing ---------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
void count_launch(int n = 1);
int sm_count();

#define MT_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return mt::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define MT_REQUIRE(cond, ...)             \
    do {                                  \
        if (!(cond)) {                    \
            mt::set_error(__VA_ARGS__);   \
            return MT_EINVAL;             \
        }                                 \
    } while (0)

#define MT_LAUNCH_CHECK()                                                              \
    do {                                                                               \
        mt::count_launch();                                                            \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess) return mt::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// optional per-launch CUDA-event timing (bench.py's roofline leg): mt_profile_enable(1), run,
// mt_profile_fetch().  Disabled (the default) it costs one relaxed atomic load per scope.
struct ProfScope {
    ProfScope(const char *name, mt_stream_t stream);
    ~ProfScope();
    int slot;
    cudaStream_t s;
};

// grid sizing: enough CTAs to cover `n` items with `per_block` items each, capped at a multiple
// of the SM count (grid-stride loops take the rest) so that every wave is full on 148 SMs.
inline unsigned grid_for(int64_t n, int per_block, int ctas_per_sm = 8)
{
    int64_t need = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

// streaming (evict-first) store for outputs that are written once and not re-read by the
// same kernel: keeps the L2 for the gathered / stencilled inputs.
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

__device__ __forceinline__ double nan_f64() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double inf_f64() { return __longlong_as_double(0x7ff0000000000000LL); }

}  // namespace mt
