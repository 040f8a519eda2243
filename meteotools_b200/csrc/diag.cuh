// Types shared by the two implementations of the fused dynamics diagnostics (csrc/diag.cu: point-wise
// stage + the register-marching stencil kernels; csrc/diag_tma.cu: the TMA plane-pipeline stencil kernel).
#pragma once
#include "fd.cuh"
#include "thermo.cuh"

namespace mt {
namespace dg {

template <int LAYOUT>
struct Lay {
    int nz, nt, nr;
    int64_t sz, st, sr;
    __host__ __device__ Lay(int nz_, int nt_, int nr_) : nz(nz_), nt(nt_), nr(nr_)
    {
        if (LAYOUT == MT_LAYOUT_ZTR) {
            sr = 1, st = nr, sz = (int64_t)nt * nr;
        } else {
            sz = 1, sr = nz, st = (int64_t)nr * nz;
        }
    }
    __device__ __forceinline__ void decode(int64_t m, int &iz, int &it, int &ir) const
    {
        if (LAYOUT == MT_LAYOUT_ZTR) {
            ir = (int)(m % nr);
            const int64_t q = m / nr;
            it = (int)(q % nt);
            iz = (int)(q / nt);
        } else {
            iz = (int)(m % nz);
            const int64_t q = m / nz;
            ir = (int)(q % nr);
            it = (int)(q / nr);
        }
    }
};

struct GridP {
    double dz, dt, dr;
    const double *r, *sin_t, *cos_t;
    int own_lo, own_hi;  // local level range [own_lo, own_hi) this call produces stencil outputs for
    bool bottom, top;    // own_lo / own_hi - 1 are the GLOBAL first / last level (one-sided formulas)
    int halo_lo, halo_hi;  // halo levels the arrays hold below / above the shard's owned levels
};

// horizontal wind of the point-wise stage: 0 = not needed, 1 = (u, v), rotated to (vr, vt) on a
// cylindrical grid (cylindrical_grid.py:264-270), 2 = (vr, vt) given by the caller (cylindrical only: the
// reference's calc_* methods use self.vr / self.vt as they are)
enum { WIND_NONE = 0, WIND_UV = 1, WIND_GIVEN = 2 };

struct StageAOut {
    double *vr, *vt, *wind_speed, *kinetic_energy, *Tv, *rho, *density_factor, *Th, *Th_rho,
        *coriolis_force, *centrifugal_force, *aam;
};

struct StageBIn {
    const double *a, *b, *w, *thr, *p, *rho, *f;  // cyl: a=vr, b=vt ; car: a=u, b=v
};
struct StageBOut {
    double *radial_PG_force, *agradient_force;
    double *zeta_r, *zeta_t, *zeta_z, *abs_zeta_z, *div_hori, *div, *PV;
    double *shear_deform, *stretch_deform, *filamentation_time;
    uint8_t *strain_region;
};
// wind-derived point-wise outputs the TMA stencil kernel produces while u, v, w are staged
struct WindOut {
    double *vr, *vt, *wind_speed, *kinetic_energy, *coriolis_force, *centrifugal_force, *aam;
};

// Stage B as a TMA plane pipeline.  `u`, `v` non-null: the horizontal wind is read from them (and, on a
// cylindrical grid, rotated in shared memory; wo.vr / wo.vt receive it); otherwise from bi.a / bi.b.
// Produces the stencil outputs of the INTERIOR points of the level range of `gp` (the outermost shell
// stays with k_stage_b<SHELL>) and the outputs of `wo` for every point of that range.
// Returns MT_ENOSUP when the shape cannot take this path (the caller falls back to the marching kernels).
template <bool CYL>
int stencil_tma(const mt_grid_t *g, const GridP &gp, const StageBIn &bi, const double *u, const double *v,
                const WindOut &wo, const StageBOut &bo, bool full, cudaStream_t s);
bool stencil_tma_supported(const mt_grid_t *g);

}  // namespace dg
}  // namespace mt
