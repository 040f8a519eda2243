"""ctypes binding of lib/fastcompute.so (include/meteotools_b200.h).

Loaded the way the reference loads its gfortran library (``ct.CDLL(path_of_meteotools +
'/lib/fastcompute.so')``, meteotools/interpolation/interp_2d.py:13-16), with ``argtypes`` for
every symbol of the header.  A missing library or a failing device call raises
``BackendError`` -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as ct
import os

from . import path_of_meteotools, platform_name
from .exceptions import BackendError

c_int, c_i64, c_dbl, c_vp = ct.c_int, ct.c_int64, ct.c_double, ct.c_void_p
c_dp = ct.POINTER(ct.c_double)

LIB_PATH = os.path.join(path_of_meteotools, "lib",
                        "fastcompute.dll" if platform_name == "Windows" else "fastcompute.so")

# constants of the header
MT_F64, MT_F32 = 0, 1
MT_IL_INCLUSIVE, MT_IL_ROUND_F32, MT_IL_SPLINE2, MT_IL_SPLINE3 = 1, 2, 4, 8
MT_MAX_VARS = 16
MT_NANSUM_SCRATCH = 2048
MT_WSUM_PAIRWISE = 2
MT_TMA_TILE_W_F64, MT_TMA_TILE_W_F32, MT_TMA_TILE_H = 14, 12, 4   # source cells per tile of mt_setdata_tma


MT_TMA_MAX_TARGETS = 128


def tma_tile(dtype):
    """(tile_w, tile_h, column alignment of the tile origin, targets per item) of mt_setdata_tma
    for MT_F64 / MT_F32."""
    tw, al = (MT_TMA_TILE_W_F32, 4) if dtype == MT_F32 else (MT_TMA_TILE_W_F64, 2)
    return tw, MT_TMA_TILE_H, al, MT_TMA_MAX_TARGETS
MT_LAYOUT_ZTR, MT_LAYOUT_TRZ = 0, 1
MT_STAGE_A, MT_STAGE_B, MT_STAGE_ALL, MT_STAGE_WIND_IN_B = 1, 2, 3, 4
MT_ABI_VERSION = 2
FD_KINDS = {"FD2": 0, "FD4": 1, "FD2_front": 2, "FD2_back": 3, "FD2_2": 4, "FD2_2_front": 5,
            "FD2_2_back": 6}
OPS = {name: i for i, name in enumerate(
    ["THETA", "T", "SAT_VAPOR", "VAPOR", "SAT_QV", "QV_TO_VAPOR", "THETA_ES", "TD", "QV", "THETA_V",
     "TV_QV", "TV_VAPOR", "RHO", "THETA_E", "MOIST_DTDZ", "RH", "P_TO_Z", "Z_TO_P", "ADD"])}


class mt_grid_t(ct.Structure):
    _fields_ = [("nz", c_i64), ("nt", c_i64), ("nr", c_i64), ("layout", c_int),
                ("dz", c_dbl), ("dt", c_dbl), ("dr", c_dbl),
                ("r", c_vp), ("sin_t", c_vp), ("cos_t", c_vp),
                ("halo_lo", c_i64), ("halo_hi", c_i64), ("is_bottom", c_int), ("is_top", c_int),
                ("sub_lo", c_i64), ("sub_hi", c_i64)]


def _ptr_struct(name, fields):
    return type(name, (ct.Structure,), {"_fields_": [(f, c_vp) for f in fields]})


TD_IN = ["T", "p", "qv", "qc", "qr"]
TD_OUT = ["Th", "vapor_s", "qvs", "vapor", "RH", "Td", "Th_es", "Th_v", "Tv", "rho", "Tc", "Th_e",
          "moist_dTdz"]
CYL_D_OUT = ["vr", "vt", "wind_speed", "kinetic_energy", "Tv", "rho", "density_factor", "Th", "Th_rho",
             "coriolis_force", "centrifugal_force", "radial_PG_force", "agradient_force",
             "zeta_r", "zeta_t", "zeta_z", "abs_zeta_z", "div_hori", "div", "PV",
             "shear_deform", "stretch_deform", "filamentation_time", "strain_region", "aam"]
CAR_D_OUT = ["wind_speed", "kinetic_energy", "Tv", "rho", "density_factor", "Th", "Th_rho",
             "zeta_r", "zeta_t", "zeta_z", "abs_zeta_z", "div_hori", "div", "PV",
             "shear_deform", "stretch_deform", "filamentation_time"]
HYDRO = ["qc", "qr", "qi", "qs", "qg", "qh"]

mt_td_in_t = _ptr_struct("mt_td_in_t", TD_IN)
mt_td_out_t = _ptr_struct("mt_td_out_t", TD_OUT)
mt_cyl_d_out_t = _ptr_struct("mt_cyl_d_out_t", CYL_D_OUT)
mt_car_d_out_t = _ptr_struct("mt_car_d_out_t", CAR_D_OUT)


class mt_dyn_in_t(ct.Structure):
    _fields_ = [(f, c_vp) for f in ("u", "v", "w", "p", "T", "Th", "qv")] + \
               [("q", c_vp * 6), ("f", c_vp), ("vr", c_vp), ("vt", c_vp)]


_SIGS = {
    # legacy (fastcompute.f90)
    "interp1D_fast": (None, [c_int, c_dbl, c_int, c_dp, c_dp, c_dp]),
    "interp1D_fast_layers": (None, [c_int, c_int, c_dbl, c_int, c_dp, c_dp, c_dp]),
    "interp1D_nonequal": (None, [c_int, c_dp, c_int, c_dp, c_dp, c_dp]),
    "interp1D_nonequal_layers": (None, [c_int, c_dp, c_int, c_int, c_dp, c_dp, c_dp]),
    "interp2D_fast": (None, [c_int, c_int, c_dbl, c_dbl, c_int, c_dp, c_dp, c_dp, c_dp]),
    "interp2D_fast_layers": (None, [c_int, c_int, c_int, c_dbl, c_dbl, c_int, c_dp, c_dp, c_dp, c_dp]),
    "interp3D_fast": (None, [c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_int, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "interp3D_fast_layers": (None, [c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_int, c_dp, c_dp,
                                    c_dp, c_dp, c_dp]),
    # device API
    "mt_abi_version": (c_int, []),
    "mt_last_error": (ct.c_char_p, []),
    "mt_device_count": (c_int, []),
    "mt_set_device": (c_int, [c_int]),
    "mt_sm_count": (c_int, []),
    "mt_launch_count": (c_i64, []),
    "mt_profile_enable": (c_int, [c_int]),
    "mt_profile_fetch": (c_int, [ct.c_char_p, c_i64]),
    "mt_interp2d_layers": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_dbl, c_dbl, c_vp,
                                   c_vp, c_i64, c_vp, c_i64, c_i64, c_vp]),
    "mt_interp2d_plan": (c_int, [c_vp, c_vp, c_i64, c_dbl, c_dbl, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "mt_gather2d_lf": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_dbl, c_dbl, c_vp,
                               c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "mt_interp3d_layers": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_dbl,
                                   c_dbl, c_dbl, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp]),
    "mt_interp1d_layers": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_dbl, c_vp, c_i64, c_vp, c_i64,
                                   c_i64, c_vp]),
    "mt_interp1d_nonequal_layers": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp,
                                            c_i64, c_i64, c_vp]),
    "mt_interplevel": (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                               c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_i64, c_i64, c_int, c_vp]),
    "mt_spline_levels": (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                 c_vp, c_i64, c_int, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "mt_spline_status": (c_int, [c_vp]),
    "mt_setdata_workspace_bytes": (c_i64, [c_int, c_i64, c_i64, c_i64, c_i64, c_i64]),
    "mt_setdata": (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                           c_dbl, c_dbl, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_vp,
                           c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_i64, c_vp]),
    "mt_setdata_fused": (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64,
                                 c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_vp, c_vp,
                                 c_i64, c_int, c_vp, c_vp]),
    "mt_axis_wsum": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_int, c_dbl, c_vp, c_i64,
                             c_i64, c_vp]),
    "mt_azimuthal_stats": (c_int, [c_vp, ct.POINTER(mt_grid_t), c_vp, c_vp, c_vp, c_vp]),
    "mt_nansum_weighted": (c_int, [c_vp, c_vp, ct.POINTER(mt_grid_t), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mt_smooth_pass": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_dbl, c_dbl, c_vp]),
    "mt_setdata_tma_supported": (c_int, [c_int, c_i64, c_i64, c_i64]),
    "mt_setdata_tma_status": (c_int, [c_vp, c_vp]),
    "mt_setdata_tma": (c_int, [c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64,
                               c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_int,
                               c_vp, c_vp]),
    "mt_upload_box_pitched": (c_int, [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64,
                                      c_vp]),
    "mt_terrain_index_cyl": (c_int, [c_vp, c_i64, c_i64, c_dbl, c_dbl, c_vp, c_vp]),
    "mt_terrain_mask_cyl": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "mt_terrain_mask_car": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "mt_fd": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_dbl, c_int, c_vp]),
    "mt_pointwise": (c_int, [c_int, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp, c_i64, c_vp]),
    "mt_calc_tc": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "mt_cyl_td": (c_int, [ct.POINTER(mt_td_in_t), ct.POINTER(mt_td_out_t), c_i64, c_vp, c_vp]),
    "mt_cyl_d": (c_int, [ct.POINTER(mt_grid_t), ct.POINTER(mt_dyn_in_t), ct.POINTER(mt_cyl_d_out_t),
                         c_int, c_vp]),
    "mt_car_d": (c_int, [ct.POINTER(mt_grid_t), ct.POINTER(mt_dyn_in_t), ct.POINTER(mt_car_d_out_t),
                         c_int, c_vp]),
    "mt_nccl_load": (c_int, [ct.c_char_p]),
    "mt_nccl_version": (c_int, []),
    "mt_nccl_unique_id": (c_int, [c_vp]),
    "mt_nccl_comm_init": (c_int, [ct.POINTER(c_vp), c_int, c_int, c_vp]),
    "mt_nccl_comm_destroy": (c_int, [c_vp]),
    "mt_halo_exchange": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_vp]),
    "mt_peer_supported": (c_int, []),
    "mt_peer_box_create": (c_int, [c_i64, ct.POINTER(c_vp), c_vp]),
    "mt_peer_box_open": (c_int, [c_vp, ct.POINTER(c_vp)]),
    "mt_peer_box_close": (c_int, [c_vp]),
    "mt_peer_box_destroy": (c_int, [c_vp]),
    "mt_peer_copy": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "mt_stream_write32": (c_int, [c_vp, ct.c_uint32, c_vp]),
    "mt_stream_wait_geq32": (c_int, [c_vp, ct.c_uint32, c_vp]),
    "mt_azimuthal_mean": (c_int, [c_vp, ct.POINTER(mt_grid_t), c_vp, c_vp, c_vp]),
    "mt_scale_box_2d": (c_int, [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_vp, c_int, c_i64, c_i64, c_i64, c_vp]),
    "mt_cast_f32_f64": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "mt_cast_f64_f32": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "mt_cyl_prelude": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_int, c_dbl, c_dbl,
                       c_vp, c_vp, c_vp]),
    "mt_rows_to_levels": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "mt_permute3": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp]),
    "mt_upload_box": (c_int, [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def load():
    """Open lib/fastcompute.so and type every symbol; raises BackendError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BackendError(
            f"{LIB_PATH} not found: build it with `python {os.path.join(path_of_meteotools, 'build.py')}` "
            "(nvcc, sm_100a).  meteotools_b200 has no CPU fallback.")
    try:
        lib = ct.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise BackendError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _SIGS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise BackendError(f"{LIB_PATH} does not export `{name}` (stale build?)") from e
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mt_last_error()
        raise BackendError(f"{what or 'meteotools_b200 backend call'} failed (code {rc}): "
                           f"{msg.decode(errors='replace') if msg else ''}")
