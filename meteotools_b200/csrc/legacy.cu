// The 8 legacy C symbols of the reference's fastcompute.f90 (include/meteotools_b200.h, group 1).
// HOST pointers in Fortran (layer-fastest) layout, synchronous: stage to the device, run the
// kernel of interp.cu on the library's own stream, copy back, synchronise.  This is the entry
// point the UNMODIFIED reference Python binds through ctypes (interp_{1d,2d,3d}.py:13-40).
// The reference ABI has no error return: failures are printed to stderr and the output is
// filled with NaN so that a broken device can never pass for a result.
#include <mutex>
#include <string.h>

#include "common.cuh"

extern "C" {
int mt_interp2d_layers(const double *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, double,
                       double, const double *, const double *, int64_t, double *, int64_t, int64_t,
                       mt_stream_t);
int mt_interp3d_layers(const double *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                       int64_t, double, double, double, const double *, const double *,
                       const double *, int64_t, double *, int64_t, int64_t, mt_stream_t);
int mt_interp1d_layers(const double *, int64_t, int64_t, int64_t, int64_t, double, const double *,
                       int64_t, double *, int64_t, int64_t, mt_stream_t);
int mt_interp1d_nonequal_layers(const double *, const double *, int64_t, int64_t, int64_t, int64_t,
                                const double *, int64_t, double *, int64_t, int64_t, mt_stream_t);
}

namespace {

// Grow-only device arena + stream, one per process, serialised by a mutex (the reference never
// calls these concurrently; SURVEY 8b "threading").
struct Arena {
    std::mutex mu;
    int device = -1;
    char *base = nullptr;
    size_t cap = 0, used = 0;
    cudaStream_t stream = nullptr;

    cudaError_t reset(size_t need)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev != device) {  // device switched: drop the old arena lazily
            base = nullptr, cap = 0, stream = nullptr, device = dev;
        }
        if (!stream && (e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess)
            return e;
        if (need > cap) {
            if (base) cudaFree(base);
            base = nullptr, cap = 0;
            size_t want = need + need / 8 + (1u << 20);
            if ((e = cudaMalloc(&base, want)) != cudaSuccess) return e;
            cap = want;
        }
        used = 0;
        return cudaSuccess;
    }
    double *take(size_t n_doubles)
    {
        size_t bytes = (n_doubles * sizeof(double) + 255) & ~size_t(255);
        char *p = base + used;
        used += bytes;
        return reinterpret_cast<double *>(p);
    }
};
Arena g_arena;

inline size_t pad(size_t n_doubles) { return (n_doubles * sizeof(double) + 255) & ~size_t(255); }

void fail(const char *name, double *out, size_t n_out)
{
    fprintf(stderr, "meteotools_b200: %s failed: %s\n", name, mt_last_error());
    for (size_t i = 0; i < n_out; ++i) out[i] = nan("");
}

#define L_CUDA(call)                                                         \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) {                                             \
            mt::cuda_fail(_e, #call, __FILE__, __LINE__);                    \
            fail(name, out_h, n_out);                                        \
            return;                                                          \
        }                                                                    \
    } while (0)

struct In {
    const double *host;
    size_t n;
};

// Generic driver: upload `nin` arrays, run `launch(dev_in[], dev_out, stream)`, download.
template <class Launch>
void run(const char *name, const In *ins, int nin, double *out_h, size_t n_out, Launch launch)
{
    std::lock_guard<std::mutex> lock(g_arena.mu);
    size_t need = pad(n_out);
    for (int i = 0; i < nin; ++i) need += pad(ins[i].n);
    L_CUDA(g_arena.reset(need));
    cudaStream_t s = g_arena.stream;
    const double *dev_in[8];
    for (int i = 0; i < nin; ++i) {
        double *d = g_arena.take(ins[i].n);
        L_CUDA(cudaMemcpyAsync(d, ins[i].host, ins[i].n * sizeof(double), cudaMemcpyHostToDevice, s));
        dev_in[i] = d;
    }
    double *dev_out = g_arena.take(n_out);
    if (launch(dev_in, dev_out, (mt_stream_t)s) != MT_OK) {
        cudaStreamSynchronize(s);
        fail(name, out_h, n_out);
        return;
    }
    L_CUDA(cudaMemcpyAsync(out_h, dev_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, s));
    L_CUDA(cudaStreamSynchronize(s));
}

}  // namespace

extern "C" {

void interp2D_fast(int lenx, int leny, double dx, double dy, int N, const double *xNewLoc,
                   const double *yNewLoc, const double *xyData, double *RThetaData)
{
    if (N <= 0) return;
    In ins[3] = {{xyData, (size_t)lenx * leny}, {xNewLoc, (size_t)N}, {yNewLoc, (size_t)N}};
    run("interp2D_fast", ins, 3, RThetaData, (size_t)N,
        [&](const double **d, double *o, mt_stream_t s) {
            // xyData(leny,lenx): y contiguous
            return mt_interp2d_layers(d[0], 1, leny, lenx, 0, 1, leny, dx, dy, d[1], d[2], N, o, 0, 1, s);
        });
}

void interp2D_fast_layers(int lenx, int leny, int layers, double dx, double dy, int N,
                          const double *xNewLoc, const double *yNewLoc, const double *xyData,
                          double *RThetaData)
{
    if (N <= 0 || layers <= 0) return;
    In ins[3] = {{xyData, (size_t)layers * lenx * leny}, {xNewLoc, (size_t)N}, {yNewLoc, (size_t)N}};
    run("interp2D_fast_layers", ins, 3, RThetaData, (size_t)layers * N,
        [&](const double **d, double *o, mt_stream_t s) {
            // xyData(layers,leny,lenx) / RThetaData(layers,N): layer index contiguous
            return mt_interp2d_layers(d[0], layers, leny, lenx, 1, layers, (int64_t)layers * leny, dx,
                                      dy, d[1], d[2], N, o, 1, layers, s);
        });
}

void interp3D_fast(int lenx, int leny, int lenz, double dx, double dy, double dz, int N,
                   const double *xNewLoc, const double *yNewLoc, const double *zNewLoc,
                   const double *xyzData, double *NewData)
{
    if (N <= 0) return;
    In ins[4] = {{xyzData, (size_t)lenx * leny * lenz}, {xNewLoc, (size_t)N}, {yNewLoc, (size_t)N},
                 {zNewLoc, (size_t)N}};
    run("interp3D_fast", ins, 4, NewData, (size_t)N, [&](const double **d, double *o, mt_stream_t s) {
        return mt_interp3d_layers(d[0], 1, lenz, leny, lenx, 0, 1, lenz, (int64_t)lenz * leny, dx, dy,
                                  dz, d[1], d[2], d[3], N, o, 0, 1, s);
    });
}

void interp3D_fast_layers(int lenx, int leny, int lenz, int layers, double dx, double dy, double dz,
                          int N, const double *xNewLoc, const double *yNewLoc,
                          const double *zNewLoc, const double *xyzData, double *NewData)
{
    if (N <= 0 || layers <= 0) return;
    In ins[4] = {{xyzData, (size_t)layers * lenx * leny * lenz}, {xNewLoc, (size_t)N},
                 {yNewLoc, (size_t)N}, {zNewLoc, (size_t)N}};
    run("interp3D_fast_layers", ins, 4, NewData, (size_t)layers * N,
        [&](const double **d, double *o, mt_stream_t s) {
            return mt_interp3d_layers(d[0], layers, lenz, leny, lenx, 1, layers,
                                      (int64_t)layers * lenz, (int64_t)layers * lenz * leny, dx, dy,
                                      dz, d[1], d[2], d[3], N, o, 1, layers, s);
        });
}

void interp1D_fast(int lenx, double dx, int Nr, const double *xNewLoc, const double *xData,
                   double *RThetaData)
{
    if (Nr <= 0) return;
    In ins[2] = {{xData, (size_t)lenx}, {xNewLoc, (size_t)Nr}};
    run("interp1D_fast", ins, 2, RThetaData, (size_t)Nr,
        [&](const double **d, double *o, mt_stream_t s) {
            return mt_interp1d_layers(d[0], 1, lenx, 0, 1, dx, d[1], Nr, o, 0, 1, s);
        });
}

void interp1D_fast_layers(int lenx, int layers, double dx, int Nr, const double *xNewLoc,
                          const double *xData, double *NewData)
{
    if (Nr <= 0 || layers <= 0) return;
    In ins[2] = {{xData, (size_t)layers * lenx}, {xNewLoc, (size_t)Nr}};
    run("interp1D_fast_layers", ins, 2, NewData, (size_t)layers * Nr,
        [&](const double **d, double *o, mt_stream_t s) {
            return mt_interp1d_layers(d[0], layers, lenx, 1, layers, dx, d[1], Nr, o, 1, layers, s);
        });
}

void interp1D_nonequal(int lenx, const double *x_input, int N, const double *x_output,
                       const double *inData, double *outData)
{
    if (N <= 0) return;
    In ins[3] = {{x_input, (size_t)lenx}, {inData, (size_t)lenx}, {x_output, (size_t)N}};
    run("interp1D_nonequal", ins, 3, outData, (size_t)N,
        [&](const double **d, double *o, mt_stream_t s) {
            return mt_interp1d_nonequal_layers(d[0], d[1], 1, lenx, 0, 1, d[2], N, o, 0, 1, s);
        });
}

void interp1D_nonequal_layers(int lenx, const double *x_input, int layers, int N,
                              const double *x_output, const double *inData, double *outData)
{
    if (N <= 0 || layers <= 0) return;
    In ins[3] = {{x_input, (size_t)lenx}, {inData, (size_t)layers * lenx}, {x_output, (size_t)N}};
    run("interp1D_nonequal_layers", ins, 3, outData, (size_t)layers * N,
        [&](const double **d, double *o, mt_stream_t s) {
            return mt_interp1d_nonequal_layers(d[0], d[1], layers, lenx, 1, layers, d[2], N, o, 1,
                                               layers, s);
        });
}

}  // extern "C"
