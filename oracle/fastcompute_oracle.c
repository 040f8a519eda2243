/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
 * product path (meteotools_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement (plain C + OpenMP) of the reference's only native code,
 * /root/reference/meteotools/fastcompute.f90 (8 bind(C) subroutines), plus a
 * restatement of the third-party vertical interpolation the reference calls
 * (wrf-python 1.3.4.1 `wrf.interplevel` -> Fortran DINTERP3DZ; source NOT under
 * /root/reference, see the header of oi_interplevel below: PARITY UNPINNED).
 *
 * gfortran is not available in this image, so the reference library itself
 * cannot be built (oracle/_ref stays empty; DESIGN.md "oracle").  This file
 * follows the Fortran statement by statement:
 *   - Fortran int() truncation            -> (int) cast
 *   - 1-based (layers,leny,lenx) F-order  -> explicit column-major offsets
 *   - the two different "x1" rules (2-D: exact-node test on x/dx;
 *     1-D/3-D: test against the last node) are kept as written
 *   - sentinel 1.7D308 for NaN targets is kept
 *   - built with gcc -O3 -fopenmp and NO -march / -ffast-math, so like the
 *     reference build line (build.py:85-94) no FMA contraction happens.
 *
 * Pinned against: test_interpolation.py:70-296, test_calc.py:58-78 and the
 * pure-Python loop oracles of test_interp.py:10-189 (tests/test_oracle_*.py).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define SENTINEL 1.7e308

/* column-major helpers: first index fastest, all indices 1-based like the Fortran */
#define IDX2(a, n1, i1, i2) (a)[((size_t)(i1) - 1) + (size_t)(n1) * ((size_t)(i2) - 1)]
#define IDX3(a, n1, n2, i1, i2, i3) \
    (a)[((size_t)(i1) - 1) + (size_t)(n1) * (((size_t)(i2) - 1) + (size_t)(n2) * ((size_t)(i3) - 1))]
#define IDX4(a, n1, n2, n3, i1, i2, i3, i4)                                                        \
    (a)[((size_t)(i1) - 1) +                                                                       \
        (size_t)(n1) * (((size_t)(i2) - 1) +                                                       \
                        (size_t)(n2) * (((size_t)(i3) - 1) + (size_t)(n3) * ((size_t)(i4) - 1)))]

/* fastcompute.f90:9-58 */
void interp2D_fast(int lenx, int leny, double dx, double dy, int N, const double *xNewLoc,
                   const double *yNewLoc, const double *xyData, double *RThetaData)
{
    (void)lenx;
#pragma omp parallel for
    for (int i = 1; i <= N; ++i) {
        double xl = xNewLoc[i - 1], yl = yNewLoc[i - 1];
        if (xl == SENTINEL || yl == SENTINEL) { /* :26-29 */
            RThetaData[i - 1] = SENTINEL;
            continue;
        }
        int x0 = (int)(xl / dx) + 1; /* :31 */
        int x1 = ((xl / dx) == (double)(int)(xl / dx)) ? x0 : x0 + 1; /* :32-36 */
        int y0 = (int)(yl / dy) + 1;
        int y1 = ((yl / dy) == (double)(int)(yl / dy)) ? y0 : y0 + 1;
        double Z00 = IDX2(xyData, leny, y0, x0); /* :45-48 */
        double Z10 = IDX2(xyData, leny, y0, x1);
        double Z01 = IDX2(xyData, leny, y1, x0);
        double Z11 = IDX2(xyData, leny, y1, x1);
        double t0 = (Z10 - Z00) * (xl / dx - (x0 - 1)) + Z00; /* :49-51 */
        double t1 = (Z11 - Z01) * (xl / dx - (x0 - 1)) + Z01;
        RThetaData[i - 1] = (t1 - t0) * (yl / dy - (y0 - 1)) + t0;
    }
}

/* fastcompute.f90:60-109 */
void interp2D_fast_layers(int lenx, int leny, int layers, double dx, double dy, int N,
                          const double *xNewLoc, const double *yNewLoc, const double *xyData,
                          double *RThetaData)
{
    (void)lenx;
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= layers; ++k) {
        for (int i = 1; i <= N; ++i) {
            double xl = xNewLoc[i - 1], yl = yNewLoc[i - 1];
            if (xl == SENTINEL || yl == SENTINEL) { /* :75-78 */
                IDX2(RThetaData, layers, k, i) = SENTINEL;
                continue;
            }
            double xtmp = xl / dx, ytmp = yl / dy; /* :80-81 */
            int x0 = (int)xtmp + 1;
            int x1 = (xtmp == (double)(int)xtmp) ? x0 : x0 + 1;
            int y0 = (int)ytmp + 1;
            int y1 = (ytmp == (double)(int)ytmp) ? y0 : y0 + 1;
            double Z00 = IDX3(xyData, layers, leny, k, y0, x0); /* :96-99 */
            double Z10 = IDX3(xyData, layers, leny, k, y0, x1);
            double Z01 = IDX3(xyData, layers, leny, k, y1, x0);
            double Z11 = IDX3(xyData, layers, leny, k, y1, x1);
            double t0 = (Z10 - Z00) * (xl / dx - (x0 - 1)) + Z00; /* :100-102 */
            double t1 = (Z11 - Z01) * (xl / dx - (x0 - 1)) + Z01;
            IDX2(RThetaData, layers, k, i) = (t1 - t0) * (yl / dy - (y0 - 1)) + t0;
        }
    }
}

/* shared body of fastcompute.f90:127-173 / :196-243 */
static inline double trilinear_at(int lenx, int leny, int lenz, double dx, double dy, double dz,
                                  double xl, double yl, double zl, const double *base,
                                  size_t stride /* =layers, elements between consecutive z */)
{
    int x0 = (int)(xl / dx) + 1;
    int x1 = (xl == (lenx - 1) * dx) ? x0 : x0 + 1; /* :135-139 */
    int y0 = (int)(yl / dy) + 1;
    int y1 = (yl == (leny - 1) * dy) ? y0 : y0 + 1;
    int z0 = (int)(zl / dz) + 1;
    int z1 = (zl == (lenz - 1) * dz) ? z0 : z0 + 1;
#define W(zi, yi, xi) \
    base[stride * (((size_t)(zi) - 1) + (size_t)lenz * (((size_t)(yi) - 1) + (size_t)leny * ((size_t)(xi) - 1)))]
    double w000 = W(z0, y0, x0), w010 = W(z0, y1, x0), w001 = W(z0, y0, x1), w011 = W(z0, y1, x1);
    double w100 = W(z1, y0, x0), w110 = W(z1, y1, x0), w101 = W(z1, y0, x1), w111 = W(z1, y1, x1);
#undef W
    double zfrac = zl / dz - (z0 - 1); /* :163-165 */
    double yfrac = yl / dy - (y0 - 1);
    double xfrac = xl / dx - (x0 - 1);
    double t00 = (w100 - w000) * zfrac + w000; /* :166-172: z first, then y, then x */
    double t10 = (w110 - w010) * zfrac + w010;
    double t01 = (w101 - w001) * zfrac + w001;
    double t11 = (w111 - w011) * zfrac + w011;
    double t0 = (t10 - t00) * yfrac + t00;
    double t1 = (t11 - t01) * yfrac + t01;
    return (t1 - t0) * xfrac + t0;
}

/* fastcompute.f90:111-177 */
void interp3D_fast(int lenx, int leny, int lenz, double dx, double dy, double dz, int N,
                   const double *xNewLoc, const double *yNewLoc, const double *zNewLoc,
                   const double *xyzData, double *NewData)
{
#pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        double xl = xNewLoc[i], yl = yNewLoc[i], zl = zNewLoc[i];
        if (xl == SENTINEL || yl == SENTINEL || zl == SENTINEL) {
            NewData[i] = SENTINEL;
            continue;
        }
        NewData[i] = trilinear_at(lenx, leny, lenz, dx, dy, dz, xl, yl, zl, xyzData, 1);
    }
}

/* fastcompute.f90:179-248 (OpenMP over layers only, :192) */
void interp3D_fast_layers(int lenx, int leny, int lenz, int layers, double dx, double dy, double dz,
                          int N, const double *xNewLoc, const double *yNewLoc,
                          const double *zNewLoc, const double *xyzData, double *NewData)
{
#pragma omp parallel for
    for (int k = 0; k < layers; ++k) {
        for (int i = 0; i < N; ++i) {
            double xl = xNewLoc[i], yl = yNewLoc[i], zl = zNewLoc[i];
            double *o = &NewData[(size_t)k + (size_t)layers * (size_t)i];
            if (xl == SENTINEL || yl == SENTINEL || zl == SENTINEL) {
                *o = SENTINEL;
                continue;
            }
            *o = trilinear_at(lenx, leny, lenz, dx, dy, dz, xl, yl, zl, xyzData + k,
                              (size_t)layers);
        }
    }
}

/* fastcompute.f90:251-284 */
void interp1D_fast(int lenx, double dx, int Nr, const double *xNewLoc, const double *xData,
                   double *RThetaData)
{
#pragma omp parallel for
    for (int i = 0; i < Nr; ++i) {
        double xl = xNewLoc[i];
        if (xl == SENTINEL) {
            RThetaData[i] = SENTINEL;
            continue;
        }
        int x0 = (int)(xl / dx) + 1;
        int x1 = (xl == (lenx - 1) * dx) ? x0 : x0 + 1; /* :270-274 */
        double Z0 = xData[x0 - 1], Z1 = xData[x1 - 1];
        RThetaData[i] = (Z1 - Z0) * (xl / dx - (x0 - 1)) + Z0; /* :279 */
    }
}

/* fastcompute.f90:287-323 */
void interp1D_fast_layers(int lenx, int layers, double dx, int Nr, const double *xNewLoc,
                          const double *xData, double *NewData)
{
#pragma omp parallel for
    for (int k = 1; k <= layers; ++k) {
        for (int i = 1; i <= Nr; ++i) {
            double xl = xNewLoc[i - 1];
            if (xl == SENTINEL) {
                IDX2(NewData, layers, k, i) = SENTINEL;
                continue;
            }
            int x0 = (int)(xl / dx) + 1;
            int x1 = (xl == (lenx - 1) * dx) ? x0 : x0 + 1;
            double Z0 = IDX2(xData, layers, k, x0), Z1 = IDX2(xData, layers, k, x1);
            IDX2(NewData, layers, k, i) = (Z1 - Z0) * (xl / dx - (x0 - 1)) + Z0;
        }
    }
}

/* bisection of fastcompute.f90:340-351: returns 1-based x0, x1 */
static inline void bisect(int lenx, const double *x_input, double xo, int *px0, int *px1)
{
    int x0 = 1, x1 = lenx;
    while ((x1 - x0) > 1) {
        int mid = (x0 + x1) / 2;
        if ((xo - x_input[mid - 1]) * (xo - x_input[x0 - 1]) <= 0)
            x1 = mid;
        else
            x0 = mid;
    }
    *px0 = x0;
    *px1 = x1;
}

/* fastcompute.f90:325-358 */
void interp1D_nonequal(int lenx, const double *x_input, int N, const double *x_output,
                       const double *inData, double *outData)
{
#pragma omp parallel for
    for (int k = 0; k < N; ++k) {
        double xo = x_output[k];
        if (xo == SENTINEL) {
            outData[k] = SENTINEL;
            continue;
        }
        int x0, x1;
        bisect(lenx, x_input, xo, &x0, &x1);
        double Z0 = inData[x0 - 1], Z1 = inData[x1 - 1];
        outData[k] = (Z1 - Z0) * ((xo - x_input[x0 - 1]) / (x_input[x1 - 1] - x_input[x0 - 1])) + Z0; /* :354 */
    }
}

/* fastcompute.f90:360-396 */
void interp1D_nonequal_layers(int lenx, const double *x_input, int layers, int N,
                              const double *x_output, const double *inData, double *outData)
{
#pragma omp parallel for
    for (int s = 1; s <= layers; ++s) {
        for (int k = 1; k <= N; ++k) {
            double xo = x_output[k - 1];
            if (xo == SENTINEL) {
                IDX2(outData, layers, s, k) = SENTINEL;
                continue;
            }
            int x0, x1;
            bisect(lenx, x_input, xo, &x0, &x1);
            double Z0 = IDX2(inData, layers, s, x0), Z1 = IDX2(inData, layers, s, x1);
            IDX2(outData, layers, s, k) =
                (Z1 - Z0) * ((xo - x_input[x0 - 1]) / (x_input[x1 - 1] - x_input[x0 - 1])) + Z0;
        }
    }
}

/*
 * Vertical interpolation to constant levels -- restatement of what the
 * reference obtains from `wrf.interplevel(var, vert, levels)` at
 * grids/cylindrical_grid.py:113 and grids/cartesian_grid.py:104.
 *
 * THIRD-PARTY, PARITY UNPINNED: wrf-python (README.md:32 names 1.3.4.1) is not
 * vendored under /root/reference, not installed here, and no reference test
 * holds a vector for it.  This restates the published algorithm of
 * wrf-python's fortran/wrf_user.f90::DINTERP3DZ as recalled:
 *   - direction decided once from column (1,1): vert[0,0,0] > vert[nz-1,0,0]
 *     means the coordinate decreases with k (pressure);
 *   - per column and level the output starts as `missing`; kp scans nz..2
 *     (1-based) FROM THE TOP; the first pair with
 *         vert[lo] < level < vert[hi]        (strict on both sides)
 *     gives  w2=(level-vert[lo])/(vert[hi]-vert[lo]); w1=1-w2;
 *            out = w1*field[lo] + w2*field[hi];
 *     increasing coordinate: lo=kp-1, hi=kp ; decreasing: lo=kp, hi=kp-1;
 *   - the Python wrapper turns `missing` into NaN (masked -> xarray -> np.array).
 * `inclusive` != 0 replaces the strict test by <= on both sides (a level that
 * coincides with a model level then yields that level's value); default strict.
 *
 * Layout: field, vert are C-order (nz, ny, nx); out is C-order (nlev, ny, nx).
 */
void oi_interplevel(const double *field, const double *vert, int nz, int ny, int nx,
                    const double *levels, int nlev, int inclusive, double *out)
{
    const size_t plane = (size_t)ny * (size_t)nx;
    int ip = 0, im = 1;
    if (vert[0] > vert[(size_t)(nz - 1) * plane]) {
        ip = 1;
        im = 0;
    }
#pragma omp parallel for collapse(2)
    for (int lev = 0; lev < nlev; ++lev) {
        for (size_t c = 0; c < plane; ++c) {
            double want = levels[lev];
            double res = NAN;
            for (int kp = nz; kp >= 2; --kp) { /* 1-based kp */
                double vlo = vert[(size_t)(kp - im - 1) * plane + c];
                double vhi = vert[(size_t)(kp - ip - 1) * plane + c];
                int hit = inclusive ? (vlo <= want && vhi >= want) : (vlo < want && vhi > want);
                if (hit) {
                    double w2 = (want - vlo) / (vhi - vlo);
                    double w1 = 1.0 - w2;
                    res = w1 * field[(size_t)(kp - im - 1) * plane + c] +
                          w2 * field[(size_t)(kp - ip - 1) * plane + c];
                    break;
                }
            }
            out[(size_t)lev * plane + c] = res;
        }
    }
}
