// Memory-pattern probe for the stencil stage of car_d: NR read streams + NW write streams over
// (nz, ny, nx) float64 arrays, no arithmetic.  Which ACCESS ORDER reaches the copy bandwidth?
//   linear   grid-stride over the flat index (what k_stage_a does: 98 % of peak)
//   march    one thread per (y, x) point, loop over z (what the marching / TMA kernels do: every step moves
//            all streams to another plane, 8-32 MB further)
//   slab     like march, but the grid walks z-slabs of `zs` levels: all CTAs stay inside one slab
//   rowmarch one thread per (z-chunk, x) column, loop over y: steps move by one row (8-16 KB)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scratch/stream_probe scratch/stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int MAXS = 12;
struct Ptrs { const double *r[MAXS]; double *w[MAXS]; int nr, nw; };

__device__ int g_store_kind = 0;   // 0 = st.cs, 1 = plain st, 2 = st.wt, 3 = st.cg
__device__ __forceinline__ void touch(const Ptrs &P, long long i)
{
    double s = 0;
    for (int k = 0; k < P.nr; ++k) s += __ldg(P.r[k] + i);
    const int kind = g_store_kind;
    for (int k = 0; k < P.nw; ++k) {
        if (kind == 0) __stcs(P.w[k] + i, s + k);
        else if (kind == 1) P.w[k][i] = s + k;
        else if (kind == 2) __stwt(P.w[k] + i, s + k);
        else __stcg(P.w[k] + i, s + k);
    }
}

__global__ void k_linear(Ptrs P, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) touch(P, i);
}
// march 32x8 with the column window of every ROW shifted so that the warp's 32 doubles start on a 256-byte
// boundary: x = tx*32 - s(y) + lane with s(y) = (y*nx) mod 32 (one more tile per row covers the tail)
__global__ void k_march_shift(Ptrs P, int nz, int ny, int nx)
{
    const int TY = blockDim.x / 32;
    const int tiles_x = (nx + 31 + 31) / 32, tiles_y = (ny + TY - 1) / TY;
    const long long ntask = (long long)tiles_x * tiles_y;
    for (long long t = blockIdx.x; t < ntask; t += gridDim.x) {
        const int ty = (int)(t / tiles_x), tx = (int)(t - (long long)ty * tiles_x);
        const int y = ty * TY + (threadIdx.x >> 5);
        if (y >= ny) continue;
        const int sh = (int)(((long long)y * nx) & 31);
        const int x = tx * 32 - sh + (threadIdx.x & 31);
        if (x < 0 || x >= nx) continue;
        for (int z = 0; z < nz; ++z) touch(P, ((long long)z * ny + y) * nx + x);
    }
}
// CTA = W x TY tile of the plane (TY = blockDim.x / W), marches z in [z0, z1)
template <int W>
__global__ void k_march_w(Ptrs P, int nz, int ny, int nx, int zs)
{
    const int TY = blockDim.x / W;
    const int tiles_x = (nx + W - 1) / W, tiles_y = (ny + TY - 1) / TY;
    const int nslab = (nz + zs - 1) / zs;
    const long long ntask = (long long)tiles_x * tiles_y * nslab;
    for (long long t = blockIdx.x; t < ntask; t += gridDim.x) {
        const int slab = (int)(t / ((long long)tiles_x * tiles_y));
        const int rem = (int)(t - (long long)slab * tiles_x * tiles_y);
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const int x = tx * W + (threadIdx.x % W), y = ty * TY + (threadIdx.x / W);
        if (x >= nx || y >= ny) continue;
        const int z0 = slab * zs, z1 = min(nz, z0 + zs);
        for (int z = z0; z < z1; ++z) touch(P, ((long long)z * ny + y) * nx + x);
    }
}
// CTA = 32 x TY tile of the plane (TY = blockDim.x / 32), marches z in [z0, z1)
__global__ void k_march(Ptrs P, int nz, int ny, int nx, int zs)
{
    const int TY = blockDim.x / 32;
    const int tiles_x = (nx + 31) / 32, tiles_y = (ny + TY - 1) / TY;
    const int nslab = (nz + zs - 1) / zs;
    const long long ntask = (long long)tiles_x * tiles_y * nslab;
    for (long long t = blockIdx.x; t < ntask; t += gridDim.x) {
        const int slab = (int)(t / ((long long)tiles_x * tiles_y));
        const int rem = (int)(t - (long long)slab * tiles_x * tiles_y);
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const int x = tx * 32 + (threadIdx.x & 31), y = ty * TY + (threadIdx.x >> 5);
        if (x >= nx || y >= ny) continue;
        const int z0 = slab * zs, z1 = min(nz, z0 + zs);
        for (int z = z0; z < z1; ++z) touch(P, ((long long)z * ny + y) * nx + x);
    }
}
// CTA = W x-points x TZ levels (TZ = blockDim.x / W), marches y; task order: x-tile fastest, then y-segment, then z-chunk
template <int W>
__global__ void k_rowmarch_w(Ptrs P, int nz, int ny, int nx, int ys)
{
    const int TZ = blockDim.x / W;
    const int tiles_x = (nx + W - 1) / W, tiles_z = (nz + TZ - 1) / TZ;
    const int nseg = (ny + ys - 1) / ys;
    const long long ntask = (long long)tiles_x * tiles_z * nseg;
    for (long long t = blockIdx.x; t < ntask; t += gridDim.x) {
        const int tz = (int)(t / ((long long)tiles_x * nseg));
        const int rem = (int)(t - (long long)tz * tiles_x * nseg);
        const int seg = rem / tiles_x, tx = rem - seg * tiles_x;
        const int x = tx * W + (threadIdx.x % W), z = tz * TZ + (threadIdx.x / W);
        if (x >= nx || z >= nz) continue;
        const int y0 = seg * ys, y1 = min(ny, y0 + ys);
        for (int y = y0; y < y1; ++y) touch(P, ((long long)z * ny + y) * nx + x);
    }
}
// CTA = 32 x-points x TZ levels (TZ = blockDim.x / 32), marches y
__global__ void k_rowmarch(Ptrs P, int nz, int ny, int nx, int ys)
{
    const int TZ = blockDim.x / 32;
    const int tiles_x = (nx + 31) / 32, tiles_z = (nz + TZ - 1) / TZ;
    const int nseg = (ny + ys - 1) / ys;
    const long long ntask = (long long)tiles_x * tiles_z * nseg;
    for (long long t = blockIdx.x; t < ntask; t += gridDim.x) {
        const int tz = (int)(t / ((long long)tiles_x * nseg));
        const int rem = (int)(t - (long long)tz * tiles_x * nseg);
        const int seg = rem / tiles_x, tx = rem - seg * tiles_x;
        const int x = tx * 32 + (threadIdx.x & 31), z = tz * TZ + (threadIdx.x >> 5);
        if (x >= nx || z >= nz) continue;
        const int y0 = seg * ys, y1 = min(ny, y0 + ys);
        for (int y = y0; y < y1; ++y) touch(P, ((long long)z * ny + y) * nx + x);
    }
}

// the store pattern of k_setdata_tma: rows of 60 doubles, one warp per row, 30 lanes x 16 bytes, 12 streams.
// pitch 60 (dense rows: 480-byte rows start at multiples of 32 bytes) or 64 (padded); `perm`: rows visited in
// a scattered order (as the source-tile bucketing does) instead of sequentially
__global__ void k_rows(Ptrs P, long long nrows, int pitch, int perm)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp0; i < nrows; i += nwarp) {
        long long row = i;
        if (perm) row = (i * 7919LL) % nrows;   // nrows is not a multiple of 7919
        if (lane < 30)
            for (int k = 0; k < P.nw; ++k)
                __stcs(reinterpret_cast<double2 *>(P.w[k] + row * pitch) + lane, make_double2((double)k, (double)lane));
    }
}
// the same bytes written as aligned 512-byte chunks (what an aligned-group store order would issue)
__global__ void k_chunks512(Ptrs P, long long n)
{
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 2; i < n; i += (long long)gridDim.x * blockDim.x * 2)
        for (int k = 0; k < P.nw; ++k) __stcs(reinterpret_cast<double2 *>(P.w[k] + i), make_double2((double)k, 1.0));
}

int main(int argc, char **argv)
{
    const int nz = argc > 1 ? atoi(argv[1]) : 80, ny = argc > 2 ? atoi(argv[2]) : 1000, nx = argc > 3 ? atoi(argv[3]) : 1000;
    const long long n = (long long)nz * ny * nx;
    std::vector<double *> bufs(17);
    for (auto &b : bufs) { CK(cudaMalloc(&b, n * 8)); CK(cudaMemset(b, 0, n * 8)); }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    auto run = [&](const char *name, int nr, int nw, auto launch) {
        Ptrs P;
        P.nr = nr, P.nw = nw;
        for (int k = 0; k < nr; ++k) P.r[k] = bufs[k];
        for (int k = 0; k < nw; ++k) P.w[k] = bufs[5 + k];
        for (int i = 0; i < 2; ++i) launch(P);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = 5;
        for (int i = 0; i < reps; ++i) launch(P);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        printf("%-34s %2dr %2dw  %8.3f ms  %7.1f GB/s\n", name, nr, nw, ms, (nr + nw) * n * 8.0 / ms / 1e6);
        fflush(stdout);
    };
    if (argc > 8) {   // setdata store pattern
        const long long nrows = n / 64;
        for (int perm = 0; perm < 2; ++perm) {
            char nm[64];
            snprintf(nm, sizeof nm, "rows of 60, pitch 60, perm %d", perm);
            run(nm, 0, 12, [&](Ptrs P) { k_rows<<<sms * 8, 256>>>(P, nrows, 60, perm); });
            snprintf(nm, sizeof nm, "rows of 60, pitch 64, perm %d", perm);
            run(nm, 0, 12, [&](Ptrs P) { k_rows<<<sms * 8, 256>>>(P, nrows, 64, perm); });
        }
        run("aligned 512-byte chunks", 0, 12, [&](Ptrs P) { k_chunks512<<<sms * 8, 256>>>(P, nrows * 60); });
        printf("(rows variants move nrows*60*8*12 bytes: scale GB/s by 60/64 = 0.9375 against the printed value)\n");
        return 0;
    }
    const bool quick = argc > 4;
    auto set_kind = [&](int k) { CK(cudaMemcpyToSymbol(g_store_kind, &k, sizeof(int))); };
    const int combos[][2] = {{5, 12}, {0, 12}, {9, 5}, {5, 0}};
    for (auto &c : combos) {
        const int nr = c[0], nw = c[1];
        set_kind(0);
        run("linear", nr, nw, [&](Ptrs P) { k_linear<<<sms * 8, 256>>>(P, n); });
        run("march 32x8", nr, nw, [&](Ptrs P) { k_march<<<sms * 8, 256>>>(P, nz, ny, nx, nz); });
        if (quick && nw != 12) continue;
        if (argc > 7) {
            for (int off : {0, 8, 16, 24}) {
                char nm[64];
                snprintf(nm, sizeof nm, "linear, arrays offset by %d doubles", off);
                run(nm, nr, nw, [&](Ptrs P) {
                    for (int k = 0; k < P.nw; ++k) P.w[k] += off;
                    for (int k = 0; k < P.nr; ++k) P.r[k] += off;
                    k_linear<<<sms * 8, 256>>>(P, n - 64);
                });
            }
            run("march 32x8, rows shifted to 256 B", nr, nw, [&](Ptrs P) { k_march_shift<<<sms * 8, 256>>>(P, nz, ny, nx); });
            run("march 32x8, rows shifted, 2 CTAs/SM", nr, nw, [&](Ptrs P) { k_march_shift<<<sms * 2, 256>>>(P, nz, ny, nx); });
            continue;
        }
        if (argc > 6) {
            for (int zs : {1, 2, 4, 8, 16, 40}) {
                char nm[64];
                snprintf(nm, sizeof nm, "march 32x8, slabs of %d levels", zs);
                run(nm, nr, nw, [&](Ptrs P) { k_march<<<sms * 8, 256>>>(P, nz, ny, nx, zs); });
            }
            run("march 256x1, slabs of 1", nr, nw, [&](Ptrs P) { k_march_w<256><<<sms * 8, 256>>>(P, nz, ny, nx, 1); });
            run("march 32x8, slabs of 1, 2 CTAs/SM", nr, nw, [&](Ptrs P) { k_march<<<sms * 2, 256>>>(P, nz, ny, nx, 1); });
            continue;
        }
        if (argc > 5) {
            const int ys = ny / 8;
            run("rowmarch 32x x 8z", nr, nw, [&](Ptrs P) { k_rowmarch_w<32><<<sms * 8, 256>>>(P, nz, ny, nx, ys); });
            run("rowmarch 64x x 4z", nr, nw, [&](Ptrs P) { k_rowmarch_w<64><<<sms * 8, 256>>>(P, nz, ny, nx, ys); });
            run("rowmarch 128x x 2z", nr, nw, [&](Ptrs P) { k_rowmarch_w<128><<<sms * 8, 256>>>(P, nz, ny, nx, ys); });
            run("rowmarch 256x x 1z", nr, nw, [&](Ptrs P) { k_rowmarch_w<256><<<sms * 8, 256>>>(P, nz, ny, nx, ys); });
            run("rowmarch 256x x 1z, whole column of rows", nr, nw, [&](Ptrs P) { k_rowmarch_w<256><<<sms * 8, 256>>>(P, nz, ny, nx, ny); });
            run("rowmarch 32x x 8z, 2 CTAs/SM", nr, nw, [&](Ptrs P) { k_rowmarch_w<32><<<sms * 2, 256>>>(P, nz, ny, nx, ys); });
            run("rowmarch 32x x 1z (32 thr), 32 CTAs/SM", nr, nw, [&](Ptrs P) { k_rowmarch_w<32><<<sms * 32, 32>>>(P, nz, ny, nx, ys); });
            continue;
        }
        run("march 64x4", nr, nw, [&](Ptrs P) { k_march_w<64><<<sms * 8, 256>>>(P, nz, ny, nx, nz); });
        run("march 128x2", nr, nw, [&](Ptrs P) { k_march_w<128><<<sms * 8, 256>>>(P, nz, ny, nx, nz); });
        run("march 256x1", nr, nw, [&](Ptrs P) { k_march_w<256><<<sms * 8, 256>>>(P, nz, ny, nx, nz); });
        run("march 32x8, 4 CTAs/SM", nr, nw, [&](Ptrs P) { k_march<<<sms * 4, 256>>>(P, nz, ny, nx, nz); });
        run("march 32x4 (128 thr), 16 CTAs/SM", nr, nw, [&](Ptrs P) { k_march<<<sms * 16, 128>>>(P, nz, ny, nx, nz); });
        run("march 32x32 (1024 thr), 2 CTAs/SM", nr, nw, [&](Ptrs P) { k_march<<<sms * 2, 1024>>>(P, nz, ny, nx, nz); });
        if (nw) {
            for (int kind = 1; kind <= 3; ++kind) {
                set_kind(kind);
                char nm[64];
                snprintf(nm, sizeof nm, "march 32x8, store kind %d", kind);
                run(nm, nr, nw, [&](Ptrs P) { k_march<<<sms * 8, 256>>>(P, nz, ny, nx, nz); });
                snprintf(nm, sizeof nm, "linear, store kind %d", kind);
                run(nm, nr, nw, [&](Ptrs P) { k_linear<<<sms * 8, 256>>>(P, n); });
            }
            set_kind(0);
        }
    }
    return 0;
}
