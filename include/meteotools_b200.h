/*
 * meteotools_b200 -- C ABI of the B200 (sm_100a) backend that replaces the reference's
 * gfortran/OpenMP `fastcompute` library (meteotools/fastcompute.f90, built by
 * meteotools/build.py:83-98 into meteotools/lib/fastcompute.so and bound with ctypes at
 * meteotools/interpolation/interp_{1d,2d,3d}.py:13-40).
 *
 * Two groups of entry points, all plain C (pointers + sizes, no C++/torch types):
 *
 *  1. The 8 LEGACY symbols, byte-for-byte the signatures of the Fortran bind(C)
 *     routines.  They take HOST pointers in the Fortran (column-major, layer-fastest)
 *     layout, do H2D -> kernel -> D2H on the current device and return synchronously,
 *     so the UNMODIFIED reference Python works when this library is installed as
 *     meteotools/lib/fastcompute.so.
 *
 *  2. The `mt_*` DEVICE-pointer API used by the rewritten internals of
 *     meteotools_b200.{interpolation,calc,grids}: every array argument is a device
 *     pointer owned by the caller (a torch tensor's data_ptr), every call is
 *     asynchronous on `stream` (a cudaStream_t passed as void*), returns 0 on success
 *     or a negative MT_E* code with a message in mt_last_error().
 *
 * Arithmetic is IEEE double, compiled with -fmad=false so that every +,-,*,/ rounds
 * exactly like the reference's gfortran -O3 / NumPy evaluation (no FMA contraction).
 */
#ifndef METEOTOOLS_B200_H
#define METEOTOOLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_ABI_VERSION 2

/* status codes */
#define MT_OK 0
#define MT_EINVAL (-1)  /* bad argument (size, layout, null pointer) */
#define MT_ECUDA (-2)   /* a CUDA runtime call failed; see mt_last_error() */
#define MT_ENOSUP (-3)  /* combination not supported */

/* ------------------------------------------------------------------------------------------
 * 1. Legacy symbols -- replace fastcompute.f90 one for one (file:line of each routine).
 *    Layout notes: `xyData(layers,leny,lenx)` etc. are Fortran order, i.e. the FIRST index is
 *    contiguous; outputs `(layers,N)` likewise.  NaN targets arrive as the sentinel 1.7e308
 *    and are answered with 1.7e308 (interpolation/check_arrays.py:43-48).
 * ---------------------------------------------------------------------------------------- */
void interp2D_fast(int lenx, int leny, double dx, double dy, int N, const double *xNewLoc,
                   const double *yNewLoc, const double *xyData,
                   double *RThetaData); /* fastcompute.f90:9 */
void interp2D_fast_layers(int lenx, int leny, int layers, double dx, double dy, int N,
                          const double *xNewLoc, const double *yNewLoc, const double *xyData,
                          double *RThetaData); /* fastcompute.f90:60 */
void interp3D_fast(int lenx, int leny, int lenz, double dx, double dy, double dz, int N,
                   const double *xNewLoc, const double *yNewLoc, const double *zNewLoc,
                   const double *xyzData, double *NewData); /* fastcompute.f90:111 */
void interp3D_fast_layers(int lenx, int leny, int lenz, int layers, double dx, double dy,
                          double dz, int N, const double *xNewLoc, const double *yNewLoc,
                          const double *zNewLoc, const double *xyzData,
                          double *NewData); /* fastcompute.f90:179 */
void interp1D_fast(int lenx, double dx, int Nr, const double *xNewLoc, const double *xData,
                   double *RThetaData); /* fastcompute.f90:251 */
void interp1D_fast_layers(int lenx, int layers, double dx, int Nr, const double *xNewLoc,
                          const double *xData, double *NewData); /* fastcompute.f90:287 */
void interp1D_nonequal(int lenx, const double *x_input, int N, const double *x_output,
                       const double *inData, double *outData); /* fastcompute.f90:325 */
void interp1D_nonequal_layers(int lenx, const double *x_input, int layers, int N,
                              const double *x_output, const double *inData,
                              double *outData); /* fastcompute.f90:360 */

/* ------------------------------------------------------------------------------------------
 * 2. Device API
 * ---------------------------------------------------------------------------------------- */
typedef void *mt_stream_t; /* cudaStream_t; NULL = legacy default stream */

int mt_abi_version(void);
const char *mt_last_error(void);  /* thread-local, valid until the next failing call */
int mt_device_count(void);        /* <0 on error (e.g. no driver) */
int mt_set_device(int device);
int mt_sm_count(void);            /* SMs of the current device */
int64_t mt_launch_count(void);    /* kernels launched by this library since load (this process) */
/* Per-launch CUDA-event timing for the roofline report: enable(1) clears the records and starts
 * bracketing every kernel launch with two events on its stream; fetch() synchronises and writes
 * "kernel_name count total_ms\n" lines into buf.  Not for use under a profiler. */
int mt_profile_enable(int on);
int mt_profile_fetch(char *buf, int64_t buf_len);

/* ---- interpolation on regular grids (a1-a4) ------------------------------------------------
 * All strides are ELEMENT strides.  `src` is addressed as src[l*s_l + (z*s_z) + y*s_y + x*s_x],
 * `out` as out[l*o_l + i*o_n].  Any layout is accepted; two are fast paths:
 *   layer-fastest (s_l == 1, the Fortran/legacy order and this library's native "TRZ" order) and
 *   C-order       (s_x == 1).
 * Semantics per point are exactly fastcompute.f90 (2-D: x1=x0 at exact nodes, :83-87;
 * 1-D/3-D: x1=x0 only at the last node, :135-139); a target equal to the legacy sentinel
 * 1.7e308 yields 1.7e308 (check_arrays.py:43-48), a NaN target yields NaN.
 * Index clamping: indices are clamped into the array (the reference reads out of bounds when
 * x/dx rounds above lenx-1; cannot happen for in-range targets with representable dx). */
int mt_interp2d_layers(const double *src, int64_t layers, int64_t leny, int64_t lenx,
                       int64_t s_l, int64_t s_y, int64_t s_x, double dx, double dy,
                       const double *xloc, const double *yloc, int64_t n, double *out,
                       int64_t o_l, int64_t o_n, mt_stream_t stream);
/* Reusable per-target stencil of the 2-D routines ("plan"): ix4[4*i..] = 0-based x0,x1,y0,y1
 * (ix4[4*i] == -1 / -2: sentinel / NaN target), w2[2*i..] = the two weights x/dx-x0, y/dy-y0.
 * Targets of a grid are fixed, so the plan is built once and reused for every variable and time
 * step; it is also the debug export through which tests check indices and weights bit for bit. */
int mt_interp2d_plan(const double *xloc, const double *yloc, int64_t n, double dx, double dy,
                     int64_t lenx, int64_t leny, int32_t *ix4, double *w2, mt_stream_t stream);
/* Multi-variable gather for layer-fastest sources (src[v][y*s_y + x*s_x + l]) into
 * layer-fastest outputs (out[v][i*o_n + l]); plan_* may be NULL (then xloc/yloc are used);
 * hgt_idx (may be NULL) applies the terrain mask of cylindrical_grid.py:171-181. */
int mt_gather2d_lf(const double *const *src, double *const *out, int nvar, int64_t layers,
                   int64_t leny, int64_t lenx, int64_t s_y, int64_t s_x, double dx, double dy,
                   const double *xloc, const double *yloc, const int32_t *plan_ix,
                   const double *plan_w, int64_t n, int64_t o_n, const int32_t *hgt_idx,
                   mt_stream_t stream);
int mt_interp3d_layers(const double *src, int64_t layers, int64_t lenz, int64_t leny,
                       int64_t lenx, int64_t s_l, int64_t s_z, int64_t s_y, int64_t s_x,
                       double dx, double dy, double dz, const double *xloc, const double *yloc,
                       const double *zloc, int64_t n, double *out, int64_t o_l, int64_t o_n,
                       mt_stream_t stream);
int mt_interp1d_layers(const double *src, int64_t layers, int64_t lenx, int64_t s_l,
                       int64_t s_x, double dx, const double *xloc, int64_t n, double *out,
                       int64_t o_l, int64_t o_n, mt_stream_t stream);
int mt_interp1d_nonequal_layers(const double *x_input, const double *src, int64_t layers,
                                int64_t lenx, int64_t s_l, int64_t s_x, const double *x_output,
                                int64_t n, double *out, int64_t o_l, int64_t o_n,
                                mt_stream_t stream);

/* ---- vertical interpolation to constant levels (a5; wrf.interplevel call sites
 *      grids/cylindrical_grid.py:113, grids/cartesian_grid.py:104; PARITY UNPINNED) ----------
 * fields[v], vert: C-order (nz,ny,nx) device arrays of `dtype` (MT_F64 / MT_F32).
 * Processes the column box [y0,y1) x [x0,x1); out[v] is addressed with box-local indices:
 *   out[v][lev*o_lev + (y-y0)*o_y + (x-x0)*o_x]     (double).
 * direction: -1 = decide from column (0,0) of `vert` as DINTERP3DZ does; 0 / 1 = the caller
 *   decided (increasing / decreasing coordinate) -- used when only a sub-box was uploaded.
 * levels_order: +1 levels ascending, -1 descending, 0 unsorted (takes the literal scan).
 * flags: MT_IL_INCLUSIVE  (<= instead of < in the bracket test),
 *        MT_IL_ROUND_F32  (round the result to float, as wrf-python does for float32 input). */
#define MT_F64 0
#define MT_F32 1
#define MT_IL_INCLUSIVE 1
#define MT_IL_ROUND_F32 2
#define MT_MAX_VARS 16
int mt_interplevel(const void *const *fields, int nvar, const void *vert, int dtype, int64_t nz,
                   int64_t ny, int64_t nx, int64_t y0, int64_t y1, int64_t x0, int64_t x1,
                   const double *levels, int64_t nlev, int levels_order, int direction,
                   double *const *out, int64_t o_lev, int64_t o_y, int64_t o_x, int flags,
                   mt_stream_t stream);

/* ---- spline vertical interpolation (`ver_interp_order = 2 / 3`:
 *      grids/cylindrical_grid.py:83-93,123-127, grids/cartesian_grid.py:61-74,108-112) -----------
 * Per column: scipy.interpolate.splrep(vert[:,j,i], field[:,j,i], k=order) followed by
 * splev(levels) -- FITPACK's interpolating spline (s = 0), end polynomials extrapolate; same
 * array / box / output conventions as mt_interplevel, float32 input is widened, the result is
 * float64 (no rounding: the reference writes into a float64 array).  A column whose coordinate
 * decreases (scipy raises ValueError "Error on input data") sets a status bit that
 * mt_spline_status() returns and clears (it synchronises the stream); the values written for such
 * a column are unspecified.  mt_setdata takes MT_IL_SPLINE2 / MT_IL_SPLINE3 in `flags` to use this
 * routine as its vertical stage. */
#define MT_IL_SPLINE2 4
#define MT_IL_SPLINE3 8
int mt_spline_levels(const void *const *fields, int nvar, const void *vert, int dtype, int64_t nz,
                     int64_t ny, int64_t nx, int64_t y0, int64_t y1, int64_t x0, int64_t x1,
                     const double *levels, int64_t nlev, int order, double *const *out,
                     int64_t o_lev, int64_t o_y, int64_t o_x, mt_stream_t stream);
int mt_spline_status(mt_stream_t stream);

/* ---- fused set_data (F1): vertical interpolation of the bounding box of the targets, then the
 *      bilinear gather, then the optional terrain mask -- the pipeline of
 *      cylindrical_grid.set_data (:95-137) / cartesian_grid.set_data (:76-125).
 * fields/vert are (nz,ny,nx) C-order arrays holding the region of the full (full_ny,full_nx)
 * source whose [0,0] column is full-source column (org_y,org_x) -- the host side uploads only the
 * bounding box of the targets.  [y0,y1)x[x0,x1) is that box in FULL-source indices.
 * Target coordinates xloc/yloc (n points, e.g. theta x r) are in source units with origin at
 * full-source index (0,0), so indices and weights are computed exactly as the reference does.  Output layout: out[v][lev*o_lev + i*o_n].
 * hgt_idx (int32, n entries) may be NULL; when given, out[lev,i] = NaN where hgt_idx[i] > lev
 * (cylindrical_grid.py:171-181).  Workspace: mt_setdata_workspace_bytes() bytes of device memory. */
int64_t mt_setdata_workspace_bytes(int nvar, int64_t nlev, int64_t y0, int64_t y1, int64_t x0,
                                   int64_t x1);
int mt_setdata(const void *const *fields, int nvar, const void *vert, int dtype, int64_t nz,
               int64_t ny, int64_t nx, int64_t org_y, int64_t org_x, int64_t full_ny,
               int64_t full_nx, double dx, double dy, int64_t y0, int64_t y1, int64_t x0, int64_t x1,
               const double *levels, int64_t nlev, int levels_order, int direction,
               const double *xloc, const double *yloc, const int32_t *plan_ix, const double *plan_w,
               int64_t n, const int32_t *hgt_idx, double *const *out, int64_t o_lev, int64_t o_n,
               int flags, void *workspace, int64_t workspace_bytes, mt_stream_t stream);

/* Fully fused variant of mt_setdata for level-fastest outputs (out[v][i*o_n + lev]): one kernel,
 * the level-interpolated columns stay in shared memory.  The caller supplies, once per plan, the
 * targets bucketed by source tile: `order` (n_valid target indices sorted by tile) and `items`
 * (nitems x 4 int32: tile_x, tile_y in full-source indices, start into `order`, count <= 128);
 * a target belongs to the tile that contains its (x0, y0) corner, tiles are tile_w x tile_h cells
 * (16x4 or 16x2).  `nflagged` = number of NaN / sentinel targets (they only get the fill value).
 * `counter`: TWO int32 of device scratch, zero before the first launch ([0]: work-item counter, [1]: the
 * status word of mt_setdata_tma, see there).  fields/vert describe the uploaded region exactly as in
 * mt_setdata. */
int mt_setdata_fused(const void *const *fields, int nvar, const void *vert, int dtype, int64_t nz,
                     int64_t ny, int64_t nx, int64_t org_y, int64_t org_x, const double *levels,
                     int64_t nlev, int direction, const int32_t *plan_ix, const double *plan_w,
                     int64_t n, const int32_t *order, const int32_t *items, int64_t nitems,
                     int tile_w, int tile_h, int64_t nflagged, const int32_t *hgt_idx,
                     double *const *out, int64_t o_n, int flags, int32_t *counter,
                     mt_stream_t stream);

/* set_data as one warp-specialised TMA kernel (csrc/setdata_tma.cu): same contract as
 * mt_setdata_fused, with these differences.  Each field is read through a 3-D TMA tensor map with
 * a {16, 5, nz} box, and the TMA unit needs every box to start on a 16-byte boundary, hence:
 *   - source rows must start 16-byte aligned: `row_pitch` (elements between consecutive rows,
 *     >= nx; the plane pitch is ny*row_pitch) times the element size must be a multiple of 16;
 *   - tiles are MT_TMA_TILE_W_F64 x MT_TMA_TILE_H source cells for float64 sources and
 *     MT_TMA_TILE_W_F32 x MT_TMA_TILE_H for float32 sources (the 16-column box covers the tile's
 *     tile_w + 1 columns), and every item's tile_x - org_x must be a multiple of 2 (float64) / 4
 *     (float32).  The kernel checks every item it draws: one that breaks this rule (the TMA unit would
 *     raise a sticky cudaErrorIllegalInstruction), or whose start / count do not fit `order` / exceed
 *     MT_TMA_MAX_TARGETS, is skipped -- its targets are not written -- and recorded in counter[1];
 *     mt_setdata_tma_status(counter, stream) synchronises, returns MT_EINVAL if anything was recorded
 *     and clears the record.  Call it once after the first launch with a new `items` array.
 * `direction` must be 0 or 1 (decided by the caller from column (0,0) of the full coordinate
 * array).  mt_setdata_tma_supported() tells whether a shape can take this path (otherwise use
 * mt_setdata: same results, two kernels).
 * At most 96 target levels (the brackets of a thread's level pairs live in registers); an odd level
 * count or output pitch is accepted (8-byte instead of 16-byte stores); at most MT_TMA_MAX_TARGETS
 * targets per item. */
#define MT_TMA_TILE_W_F64 14
#define MT_TMA_TILE_W_F32 12
#define MT_TMA_TILE_H 4
#define MT_TMA_MAX_TARGETS 128
int mt_setdata_tma_supported(int dtype, int64_t nz, int64_t nlev, int64_t row_pitch);
int mt_setdata_tma(const void *const *fields, int nvar, const void *vert, int dtype, int64_t nz,
                   int64_t ny, int64_t nx, int64_t row_pitch, int64_t org_y, int64_t org_x,
                   const double *levels, int64_t nlev, int direction, const int32_t *plan_ix,
                   const double *plan_w, int64_t n, const int32_t *order, const int32_t *items,
                   int64_t nitems, int64_t nflagged, const int32_t *hgt_idx, double *const *out,
                   int64_t o_n, int flags, int32_t *counter, mt_stream_t stream);
int mt_setdata_tma_status(int32_t *counter, mt_stream_t stream);

/* mt_upload_box with a padded device row pitch (elements): dev is (nz, y1-y0, dev_row_pitch). */
int mt_upload_box_pitched(const void *host, int elem_bytes, int64_t nz, int64_t ny, int64_t nx,
                          int64_t y0, int64_t y1, int64_t x0, int64_t x1, void *dev,
                          int64_t dev_row_pitch, mt_stream_t stream);

/* terrain index of the cylindrical grid (cylindrical_grid.py:141-170): hgt (nt x nr, C-order)
 * -> hgt_idx int32 incl. the edge/corner fix-up.  Cartesian mask (cartesian_grid.py:127-136):
 * var[k,j,i] = NaN where z[k] <= hgt[j,i]. */
int mt_terrain_index_cyl(const double *hgt, int64_t nt, int64_t nr, double z_start, double dz,
                         int32_t *hgt_idx, mt_stream_t stream);
/* The 2-D part of a cylindrical set_data step in one launch: out2d[f][i] = bilinear interpolation of the
 * C-order (ny, nx) field src2d[f] at target i through the plan of mt_interp2d_plan (the same indices and
 * weights as interp2D_fast, fastcompute.f90:9-58; flagged targets get the sentinel / NaN), f < nfield <= 4;
 * when hgt_field >= 0 that field's result also gives hgt_idx exactly as mt_terrain_index_cyl does
 * (cylindrical_grid.py:141-170; the fix-up runs in the last block to finish: `ticket` is one int32 that
 * is zero before the first launch and is left zero).  Targets are the nt x nr grid in C order. */
int mt_cyl_prelude(const double *const *src2d, double *const *out2d, int nfield, int64_t ny, int64_t nx,
                   const int32_t *plan_ix, const double *plan_w, int64_t nt, int64_t nr, int hgt_field,
                   double z_start, double dz, int32_t *hgt_idx, int32_t *ticket, mt_stream_t stream);
int mt_terrain_mask_cyl(double *var, int64_t nz, int64_t n, int64_t s_z, int64_t s_n,
                        const int32_t *hgt_idx, mt_stream_t stream);
int mt_terrain_mask_car(double *var, int64_t nz, int64_t n, int64_t s_z, int64_t s_n,
                        const double *z, const double *hgt, mt_stream_t stream);

/* ---- finite differences (a8, a9; calc/differential.py:6-180) ---------------------------------
 * `in`/`out` are dense arrays viewed as (outer, n, inner) with the differentiated axis in the
 * middle: element (o,i,j) at o*n*inner + i*inner + j. */
#define MT_FD2 0
#define MT_FD4 1
#define MT_FD2_FRONT 2
#define MT_FD2_BACK 3
#define MT_FD2_2 4
#define MT_FD2_2_FRONT 5
#define MT_FD2_2_BACK 6
int mt_fd(const double *in, double *out, int64_t outer, int64_t n, int64_t inner, double delta,
          int kind, mt_stream_t stream);

/* ---- point-wise thermodynamics (a12; calc/thermaldynamic.py) --------------------------------
 * out[i] = op(a[i], b[i], c[i]); unused inputs may be NULL; a scalar second/third operand is
 * passed through sb/sc with the pointer NULL. */
#define MT_OP_THETA 0        /* a=T, b=P                     :70  */
#define MT_OP_T 1            /* a=theta, b=P                 :87  */
#define MT_OP_SAT_VAPOR 2    /* a=T                          :104 */
#define MT_OP_VAPOR 3        /* a=T, b=RH                    :119 */
#define MT_OP_SAT_QV 4       /* a=T, b=P                     :137 */
#define MT_OP_QV_TO_VAPOR 5  /* a=P, b=qv                    :154 */
#define MT_OP_THETA_ES 6     /* a=T, b=P                     :171 */
#define MT_OP_TD 7           /* a=es                         :224 */
#define MT_OP_QV 8           /* a=P, b=vapor                 :256 */
#define MT_OP_THETA_V 9      /* a=theta, b=qv, c=ql          :309 */
#define MT_OP_TV_QV 10       /* a=T, b=qv                    :335 */
#define MT_OP_TV_VAPOR 11    /* a=T, b=vapor, c=P            :331-333 */
#define MT_OP_RHO 12         /* a=P, b=Tv                    :364 */
#define MT_OP_THETA_E 13     /* a=theta, b=qv, c=Tc          :456 */
#define MT_OP_MOIST_DTDZ 14  /* a=T, b=qvs                   :486 */
#define MT_OP_RH 15          /* a=vapor, b=T                 :529 */
#define MT_OP_P_TO_Z 16      /* a=P, sb=P0, sc=Tv            :33  */
#define MT_OP_Z_TO_P 17      /* a=Z, sb=P0, sc=Tv            :51  */
#define MT_OP_ADD 18        /* a + b (ql = qc + qr, thermaldynamic_calculation.py:70-72) */
#define MT_OP_COUNT 19
int mt_pointwise(int op, const double *a, const double *b, const double *c, double sb, double sc,
                 double *out, int64_t n, mt_stream_t stream);

/* condensation temperature (calc/thermaldynamic.py:367-402): up to 10 fixed-point iterations
 * with the reference's GLOBAL stop test np.max(|Tc-Tc2|) < 0.1 (NaN-sensitive).  `scratch`:
 * 16 + n doubles of device memory (16 control words + the sweep-invariant term per point).  Fully asynchronous (the stop test is evaluated on the device);
 * scratch[15] receives the number of iterations executed (as a double). */
int mt_calc_tc(const double *T, const double *P, const double *qv, double *Tc, int64_t n,
               double *scratch, mt_stream_t stream);

/* ---- fused diagnostic sets (F2-F4) -----------------------------------------------------------
 * All 3-D fields of one call share one layout: MT_LAYOUT_ZTR (C-order (z,t,r), r contiguous) or
 * MT_LAYOUT_TRZ ((t,r,z) order, z contiguous -- what mt_setdata produces with o_lev==1).
 * For the Cartesian sets read (z,t,r) as (z,y,x).  2-D fields are dense (t,r).
 * Output pointers that are NULL are not computed/stored. */
#define MT_LAYOUT_ZTR 0
#define MT_LAYOUT_TRZ 1

typedef struct {
    int64_t nz, nt, nr;
    int layout;
    double dz, dt, dr;           /* grid spacings (dz, dtheta|dy, dr|dx) */
    const double *r;             /* nr radii (cyl only)            */
    const double *sin_t, *cos_t; /* nt host-computed tables (cyl)  */
    /* the z range [z_lo, z_hi) owned by this call and the global nz are equal to (0,nz) unless
       the field is a level shard with halo levels (multi-GPU, car_d/cyl_d):
       arrays then hold nz = (z_hi-z_lo) + halo_lo + halo_hi levels. */
    int64_t halo_lo, halo_hi;    /* number of halo levels below / above (0 or 1)            */
    int is_bottom, is_top;       /* shard touches the global first / last level              */
    /* (ABI 2) restrict the stencil stage of mt_cyl_d / mt_car_d to the owned levels [sub_lo, sub_hi)
       (indices into the owned block; both 0 = all of it): a level shard first runs the levels that need
       no halo while the halo planes travel, then its boundary levels (pipeline.LevelShardDyn). */
    int64_t sub_lo, sub_hi;
} mt_grid_t;

typedef struct { /* inputs: thermaldynamic_calculation.py:7-133 on a grid holding T,p,qv(,qc,qr) */
    const double *T, *p, *qv, *qc, *qr;
} mt_td_in_t;
typedef struct {
    double *Th, *vapor_s, *qvs, *vapor, *RH, *Td, *Th_es, *Th_v, *Tv, *rho, *Tc, *Th_e,
        *moist_dTdz;
} mt_td_out_t;
/* scratch: 16 + n doubles (see mt_calc_tc).  Tc, Th and qv handling: Th_e needs Tc; when out->Tc is
 * NULL but Th_e is requested the call fails with MT_EINVAL. */
int mt_cyl_td(const mt_td_in_t *in, const mt_td_out_t *out, int64_t n, double *scratch,
              mt_stream_t stream);

typedef struct {
    const double *u, *v, *w, *p, *T, *Th /* may be NULL: computed from T,p */, *qv;
    const double *q[6];  /* qc,qr,qi,qs,qg,qh ; NULL-terminated prefix semantics as the reference */
    const double *f;     /* (nt,nr) Coriolis */
    /* (ABI 2, mt_cyl_d only) the caller's radial / tangential wind.  When both are set they ARE the
       horizontal wind of the call -- nothing is recomputed from u, v, exactly as the reference's calc_*
       methods use self.vr / self.vt once they exist (cylindrical_grid.py:290-508) -- and u, v are needed
       for wind_speed only; out->vr / out->vt must then be NULL. */
    const double *vr, *vt;
} mt_dyn_in_t;
typedef struct { /* cylindrical_grid.py:264-508 */
    double *vr, *vt, *wind_speed, *kinetic_energy, *Tv, *rho, *density_factor, *Th, *Th_rho;
    double *coriolis_force, *centrifugal_force, *radial_PG_force, *agradient_force;
    double *zeta_r, *zeta_t, *zeta_z, *abs_zeta_z, *div_hori, *div, *PV;
    double *shear_deform, *stretch_deform, *filamentation_time;
    uint8_t *strain_region;
    double *aam; /* absolute angular momentum vt*r + 0.5*f*r**2 (:290-297) */
} mt_cyl_d_out_t;
/* stages: bit 0 = stage A (point-wise fields), bit 1 = stage B (stencils); 3 = both (the library then
 * decides which kernel produces the wind-derived point-wise fields).
 *   MT_STAGE_B alone reads vr / vt (cyl), Th_rho and rho from the OUTPUT buffers of an earlier stage-A
 *   call (or vr / vt from in->vr / in->vt).
 *   MT_STAGE_WIND_IN_B moves the wind-derived point-wise outputs (vr, vt, wind_speed, kinetic_energy,
 *   coriolis_force, centrifugal_force, aam) from stage A to stage B, where the TMA stencil kernel has
 *   u, v, w staged in shared memory anyway.  A level-sharded run exchanges one halo level of the INPUTS
 *   u, v, w (v, w for car_d), calls MT_STAGE_A | MT_STAGE_WIND_IN_B (thermodynamic point-wise fields),
 *   exchanges one halo level of Th_rho, then calls MT_STAGE_B | MT_STAGE_WIND_IN_B -- first for the
 *   levels that need no halo (mt_grid_t.sub_lo / sub_hi), while the planes travel (mt_halo_exchange).
 * Stage B runs as a TMA plane pipeline (csrc/diag_tma.cu) for arrays whose contiguous axis has an even
 * length (16-byte aligned rows), otherwise as register-marching kernels; same results bit for bit. */
#define MT_STAGE_A 1
#define MT_STAGE_B 2
#define MT_STAGE_ALL 3
#define MT_STAGE_WIND_IN_B 4
int mt_cyl_d(const mt_grid_t *g, const mt_dyn_in_t *in, const mt_cyl_d_out_t *out, int stages,
             mt_stream_t stream);

typedef struct { /* cartesian_grid.py:232-364 */
    double *wind_speed, *kinetic_energy, *Tv, *rho, *density_factor, *Th, *Th_rho;
    double *zeta_r, *zeta_t, *zeta_z, *abs_zeta_z, *div_hori, *div, *PV;
    double *shear_deform, *stretch_deform, *filamentation_time;
} mt_car_d_out_t;
int mt_car_d(const mt_grid_t *g, const mt_dyn_in_t *in, const mt_car_d_out_t *out, int stages,
             mt_stream_t stream);

/* ---- multi-GPU: one-level vertical halo exchange over NCCL (csrc/halo.cu) -----------------
 * Vertical levels couple only through FD2(., dz, axis=0) (grids/cartesian_grid.py:250-251,275,321,
 * grids/cylindrical_grid.py:388-389,419,466): a rank that owns a contiguous block of levels of a
 * MT_LAYOUT_ZTR array needs ONE plane of each z-differentiated field from each neighbour.
 * The library never links against NCCL: it binds at run time to the libnccl.so.2 already loaded in the
 * process (torch.distributed's), or to `path` (mt_nccl_load; NULL = only look for a loaded copy).
 * Communicator: rank 0 calls mt_nccl_unique_id(128-byte buffer), the bytes reach the other ranks by any
 * means (torch.distributed broadcast), every rank calls mt_nccl_comm_init.
 * mt_halo_exchange: fields[f] are (nz_local, ny, nx) C-order device arrays whose owned levels are
 * [halo_lo, nz_local - halo_hi); for each field, in ONE NCCL group on `stream`: level halo_lo is sent to
 * rank_lo and level 0 received from it (when halo_lo = 1), level nz_local-1-halo_hi is sent to rank_hi
 * and level nz_local-1 received from it (when halo_hi = 1).  Asynchronous on `stream`: run it on a side
 * stream while mt_car_d / mt_cyl_d work on the levels that need no halo (mt_grid_t.sub_lo / sub_hi). */
int mt_nccl_load(const char *path);
int mt_nccl_version(void); /* NCCL_VERSION_CODE of the bound library, 0 when none */
int mt_nccl_unique_id(void *id128);
int mt_nccl_comm_init(void **comm, int nranks, int rank, const void *id128);
int mt_nccl_comm_destroy(void *comm);
int mt_halo_exchange(void *comm, double *const *fields, int nfields, int64_t plane_elems, int64_t nz_local,
                     int halo_lo, int halo_hi, int rank_lo, int rank_hi, mt_stream_t stream);

/* ---- the same halo planes over NVLink / NVSwitch PEER MEMORY (csrc/peer.cu): no kernel, no SM -----------------
 * A "box" is device memory its owner allocates (mt_peer_box_create, which also returns the 64-byte CUDA IPC
 * handle) and the neighbouring process maps (mt_peer_box_open).  Planes travel with mt_peer_copy
 * (cudaMemcpyAsync: the copy engines, straight into the neighbour's box) and the two processes order themselves
 * with stream memory operations: mt_stream_write32(addr, v) stores v when the stream reaches it,
 * mt_stream_wait_geq32(addr, v) holds the stream until (int32)(*addr - v) >= 0 (cuStreamWriteValue32 /
 * cuStreamWaitValue32).  pipeline.PeerHalo builds the one-level halo exchange of a level shard from these
 * (layout and protocol in csrc/peer.cu); the processes must share one node with peer access between the GPUs. */
int mt_peer_supported(void);
int mt_peer_box_create(int64_t bytes, void **box, void *ipc_handle64);
int mt_peer_box_open(const void *ipc_handle64, void **peer_ptr);
int mt_peer_box_close(void *peer_ptr);
int mt_peer_box_destroy(void *box);
int mt_peer_copy(void *dst, const void *src, int64_t bytes, mt_stream_t stream);
int mt_stream_write32(void *addr, uint32_t value, mt_stream_t stream);
int mt_stream_wait_geq32(void *addr, uint32_t value, mt_stream_t stream);

/* ---- azimuthal statistics (a13; cylindrical_grid.py:214-235) -------------------------------
 * sym[z,r] = nanmean over theta, asy = var - sym (asy may be NULL). */
int mt_azimuthal_mean(const double *var, const mt_grid_t *g, double *sym, double *asy,
                      mt_stream_t stream);

/* ---- reductions and smoothers next to the hot path (csrc/reduce.cu) ------
 * mt_axis_wsum: weighted NaN-skipping sum along one axis of a strided 3-D view (kept axes a, b;
 * reduced axis k), the kernel behind calc_Zaverage / calc_Raverage / calc_RZaverage
 * (calc/averages.py:13-114):
 *   out[a*o_a + b*o_b] = ( sum_k var[a*s_a + b*s_b + k*s_k] * w[k], NaN products skipped ) / D
 *   mode 0: D = denom;  mode 1: D = sum of w[k] over the terms that were not skipped (nanmean).
 *   mode | MT_WSUM_PAIRWISE: the terms are added in the order NumPy uses along the CONTIGUOUS axis
 *   of the reference's array (pairwise blocks of 8 accumulators, halving recursion above 128
 *   terms); without the flag they are added sequentially, which is NumPy's order along a strided
 *   axis.  With the right flag the sum is bit-identical to the reference's np.nansum. */
#define MT_WSUM_PAIRWISE 2
int mt_axis_wsum(const double *var, int64_t n_a, int64_t n_b, int64_t n_k, int64_t s_a, int64_t s_b,
                 int64_t s_k, const double *w, int mode, double denom, double *out, int64_t o_a,
                 int64_t o_b, mt_stream_t stream);
/* sym[z,r] = nanmean over theta; asy = var - sym; axisym[z,r] = sym^2 / (sym^2 + nansum(asy^2,
 * theta) * dtheta / 2 / pi) (cylindrical_grid.py:187-235).  Any of the three outputs may be NULL;
 * sym and axisym are (nz, nr) C-order, asy has the grid's layout.  Uses g->dt as dtheta. */
int mt_azimuthal_stats(const double *var, const mt_grid_t *g, double *sym, double *asy, double *axisym,
                       mt_stream_t stream);
/* out[0] = nansum( (a*b) * weight ) over the grid (b may be NULL): weight = (wr[r]*wz[z])*fr[r]
 * (integrated kinetic energy, cylindrical_grid.py:554-575: wr = r*dr*dtheta*dz evaluated on the
 * host, wz / fr the 0-1 range filters) or, when w2d is given, the explicit (nt, nr) weight
 * (calc_IKE10 :577-600, with nz = 1).  scratch: MT_NANSUM_SCRATCH doubles of device memory. */
#define MT_NANSUM_SCRATCH 2048
int mt_nansum_weighted(const double *a, const double *b, const mt_grid_t *g, const double *wr,
                       const double *wz, const double *fr, const double *w2d, double *scratch,
                       double *out, mt_stream_t stream);
/* One Jacobi pass of the 1-c-1 smoothers of cartesian_grid.py:142-229 on a C-order (n0,n1,n2)
 * array: out = (sum of the 2*naxes neighbours of `cur` + cur*center_weight) / denom, in the
 * reference's order of additions; a neighbour outside the array is the frozen ghost cell, i.e. the
 * ORIGINAL field v0 at the edge point.  naxes = 1 (axis ax_a), 2 (axes ax_a, ax_b) or 3.
 * `cur` and `out` must not alias (passes ping-pong). */
int mt_smooth_pass(const double *v0, const double *cur, double *out, int64_t n0, int64_t n1, int64_t n2,
                   int naxes, int ax_a, int ax_b, double center_weight, double denom,
                   mt_stream_t stream);

/* ---- staging helpers used by the host-facing wrappers -----------------------------------------
 * out[i] = in[i] as double (f32 -> f64 upcast on the device; exact). */
int mt_cast_f32_f64(const float *in, double *out, int64_t n, mt_stream_t stream);
/* Map-factor scaling of the ingest (wrfout_grid.correct_map_scale, grids/wrfout_grid.py:48-55:
 * ua *= MAPFAC_MX, va *= MAPFAC_MY, u10 / v10 likewise): box[k][j][i] *= fac[(org_y + j)*fac_pitch + org_x + i]
 * IN PLACE on a staged source box (nz, by, row_pitch) whose [0][0] column is column (org_y, org_x) of the full
 * grid the (., fac_pitch) factor array covers.  NumPy's arithmetic per dtype pair: float32 *= float32 is a
 * float product, every other pairing the double product cast back to the box's type. */
int mt_scale_box_2d(void *box, int dtype, int64_t nz, int64_t by, int64_t bx, int64_t row_pitch, const void *fac,
                    int fac_dtype, int64_t fac_pitch, int64_t org_y, int64_t org_x, mt_stream_t stream);
/* float64 -> float32, round to nearest (what the reference's netCDF cache stores: 'f4',
 * grids/grid_general.py:119-129): for results that leave the device as float32. */
int mt_cast_f64_f32(const double *in, float *out, int64_t n, mt_stream_t stream);
/* The rows mt_setdata_tma writes (n targets x nlev levels, level-fastest) -> the (z, y, x) layout of
 * a Cartesian grid, all variables of a set_data call in ONE launch: out[v][k*n + i] = rows[v][i*nlev + k],
 * and, when z and hgt are given, NaN where z[k] <= hgt[i] (the reference's terrain mask,
 * grids/cartesian_grid.py:127-136, otherwise a separate pass per variable).  nvar <= MT_MAX_VARS,
 * nlev <= 96; z and hgt are both NULL or both set. */
int mt_rows_to_levels(const double *const *rows, double *const *out, int nvar, int64_t n, int64_t nlev,
                      const double *z, const double *hgt, mt_stream_t stream);
/* dense 3-D permutation copy: dst (memory order b0,b1,b2) <- src (a0,a1,a2) with
 * dst[i,j,k] = src[...] given src element strides for dst's index order. */
int mt_permute3(const double *src, double *dst, int64_t n0, int64_t n1, int64_t n2, int64_t s0,
                int64_t s1, int64_t s2, mt_stream_t stream);

/* strided H2D upload of the column box [y0,y1)x[x0,x1) of a C-order (nz,ny,nx) HOST array (4- or
 * 8-byte elements) into a dense (nz,y1-y0,x1-x0) device array: the host side of the "upload only
 * what the targets touch" rule of mt_setdata.  Asynchronous when `host` is pinned. */
int mt_upload_box(const void *host, int elem_bytes, int64_t nz, int64_t ny, int64_t nx, int64_t y0,
                  int64_t y1, int64_t x0, int64_t x1, void *dev, mt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* METEOTOOLS_B200_H */
