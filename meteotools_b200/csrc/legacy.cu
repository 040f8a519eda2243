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

// Grow-only device arena + stream, one per DEVICE (a process that alternates between GPUs keeps one arena
// on each: nothing is dropped or leaked on a switch), serialised by one mutex (the reference never calls
// these concurrently; SURVEY 8b "threading").
struct Arena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    cudaStream_t stream = nullptr;

    cudaError_t reset(size_t need)
    {
        cudaError_t e;
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
constexpr int MAX_DEVICES = 64;
std::mutex g_mu;
Arena g_arenas[MAX_DEVICES];

// the arena of the CURRENT device (nullptr + error text when the device cannot be determined)
Arena *current_arena()
{
    int dev = 0;
    const cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        mt::cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
        return nullptr;
    }
    if (dev < 0 || dev >= MAX_DEVICES) {
        mt::set_error("device index %d out of range", dev);
        return nullptr;
    }
    return &g_arenas[dev];
}

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

// One input of a call: a dense host array of n doubles, or (rows > 0) a 2-D window of a larger host array --
// `rows` rows of `n` contiguous doubles each, `pitch` doubles apart -- that lands densely on the device.
struct In {
    const double *host;
    size_t n;
    size_t rows = 0, pitch = 0;
    size_t total() const { return rows ? rows * n : n; }
};

// Generic driver: upload `nin` arrays, run `launch(dev_in[], dev_out, stream)`, download.
template <class Launch>
void run(const char *name, const In *ins, int nin, double *out_h, size_t n_out, Launch launch)
{
    std::lock_guard<std::mutex> lock(g_mu);
    Arena *arena = current_arena();
    if (!arena) {
        fail(name, out_h, n_out);
        return;
    }
    size_t need = pad(n_out);
    for (int i = 0; i < nin; ++i) need += pad(ins[i].total());
    L_CUDA(arena->reset(need));
    cudaStream_t s = arena->stream;
    const double *dev_in[8];
    for (int i = 0; i < nin; ++i) {
        double *d = arena->take(ins[i].total());
        if (ins[i].rows)
            L_CUDA(cudaMemcpy2DAsync(d, ins[i].n * sizeof(double), ins[i].host, ins[i].pitch * sizeof(double),
                                     ins[i].n * sizeof(double), ins[i].rows, cudaMemcpyHostToDevice, s));
        else
            L_CUDA(cudaMemcpyAsync(d, ins[i].host, ins[i].n * sizeof(double), cudaMemcpyHostToDevice, s));
        dev_in[i] = d;
    }
    double *dev_out = arena->take(n_out);
    if (launch(dev_in, dev_out, (mt_stream_t)s) != MT_OK) {
        cudaStreamSynchronize(s);
        fail(name, out_h, n_out);
        return;
    }
    L_CUDA(cudaMemcpyAsync(out_h, dev_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, s));
    L_CUDA(cudaStreamSynchronize(s));
}

// Bounding box [x0, x1) x [y0, y1) of the source columns the bilinear stencils of all targets touch
// (fastcompute.f90:80-93: x0 = int(x/dx), x1 <= x0 + 1).  The 2-D routines then upload ONLY that window of
// the source (cart -> cyl of a 1000 x 1000 WRF field touches 9 % of it: 43 MB instead of 480 MB of pageable
// host memory per call).  `whole` when the window would not pay, or when a target lies outside the array (the
// kernel clamps such indices into the FULL array, which must then be resident).
struct Box {
    int x0, x1, y0, y1;
    bool whole;
};
Box target_box(const double *x, const double *y, int n, double dx, double dy, int lenx, int leny)
{
    Box b{0, lenx, 0, leny, true};
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    int valid = 0;
    for (int i = 0; i < n; ++i) {
        const double xi = x[i], yi = y[i];
        if (!(xi == xi) || !(yi == yi) || xi == MT_SENTINEL || yi == MT_SENTINEL) continue;   // NaN / sentinel targets
        const double tx = xi / dx, ty = yi / dy;
        xmin = tx < xmin ? tx : xmin, xmax = tx > xmax ? tx : xmax;
        ymin = ty < ymin ? ty : ymin, ymax = ty > ymax ? ty : ymax;
        ++valid;
    }
    if (!valid || xmin < 0 || ymin < 0 || xmax > lenx - 1 || ymax > leny - 1) return b;
    b.x0 = (int)xmin, b.y0 = (int)ymin;
    b.x1 = (int)xmax + 2 < lenx ? (int)xmax + 2 : lenx;
    b.y1 = (int)ymax + 2 < leny ? (int)ymax + 2 : leny;
    b.whole = (double)(b.x1 - b.x0) * (b.y1 - b.y0) > 0.6 * (double)lenx * leny;
    if (b.whole) b = Box{0, lenx, 0, leny, true};
    return b;
}

}  // namespace

extern "C" {

void interp2D_fast(int lenx, int leny, double dx, double dy, int N, const double *xNewLoc,
                   const double *yNewLoc, const double *xyData, double *RThetaData)
{
    interp2D_fast_layers(lenx, leny, 1, dx, dy, N, xNewLoc, yNewLoc, xyData, RThetaData);
}

void interp2D_fast_layers(int lenx, int leny, int layers, double dx, double dy, int N,
                          const double *xNewLoc, const double *yNewLoc, const double *xyData,
                          double *RThetaData)
{
    if (N <= 0 || layers <= 0) return;
    // xyData(layers,leny,lenx) / RThetaData(layers,N): layer index contiguous, x slowest.  Only the window of
    // the source under the targets travels: rows = x, each row the y-range of that x, `layers` per column.
    const Box b = target_box(xNewLoc, yNewLoc, N, dx, dy, lenx, leny);
    const int64_t by = b.y1 - b.y0, bx = b.x1 - b.x0;
    In src{xyData + ((int64_t)b.y0 + (int64_t)leny * b.x0) * layers, (size_t)(by * layers), (size_t)bx,
           (size_t)leny * layers};
    if (b.whole) src = In{xyData, (size_t)layers * lenx * leny};
    In ins[3] = {src, {xNewLoc, (size_t)N}, {yNewLoc, (size_t)N}};
    const char *name = layers == 1 ? "interp2D_fast" : "interp2D_fast_layers";
    run(name, ins, 3, RThetaData, (size_t)layers * N, [&](const double **d, double *o, mt_stream_t s) {
        // the window is addressed with FULL-array indices: element (l, y, x) of the dense window sits at
        // l + layers*((y - y0) + by*(x - x0)), i.e. at the virtual base  d[0] - layers*(y0 + by*x0)
        const double *base = d[0] - (int64_t)layers * ((int64_t)b.y0 + by * b.x0);
        return mt_interp2d_layers(base, layers, leny, lenx, 1, layers, (int64_t)layers * by, dx, dy, d[1], d[2], N,
                                  o, 1, layers, s);
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
