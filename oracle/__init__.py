"""CPU oracle for the meteotools hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``meteotools_b200/`` may import this package.  It is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker / CPU baseline.

Contents
--------
* ``fastcompute_oracle.c``  C/OpenMP restatement of the reference's
  ``meteotools/fastcompute.f90`` (+ wrf-python ``DINTERP3DZ``; parity unpinned).
* this module              numpy-facing wrappers that restate the reference's
  Python layer ``meteotools/interpolation/*.py`` (argument checks, NaN<->sentinel,
  layer flattening, C->F layout copies) around that library.
* ``calc_oracle.py``        numpy restatement of ``meteotools/calc`` and of the
  diagnostic methods of ``meteotools/grids``.

* ``spline_levels``         the reference's ``spline`` method (per-column
  ``scipy.interpolate.splrep`` + ``splev``): scipy IS the third-party dependency the
  reference calls and it is installed, so this row's oracle is the real thing.

gfortran is absent from this image, so ``oracle/_ref`` (the real reference
library) cannot be built; every number labelled "cpu_baseline" is therefore
``kind="port"``.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_fastcompute.so")
_REF_SO = os.path.join(_HERE, "_ref", "fastcompute.so")
SENTINEL = 1.7e308

_ci, _cd, _cp = ct.c_int, ct.c_double, ct.POINTER(ct.c_double)
_SIGS = {
    "interp1D_fast": [_ci, _cd, _ci, _cp, _cp, _cp],
    "interp1D_fast_layers": [_ci, _ci, _cd, _ci, _cp, _cp, _cp],
    "interp1D_nonequal": [_ci, _cp, _ci, _cp, _cp, _cp],
    "interp1D_nonequal_layers": [_ci, _cp, _ci, _ci, _cp, _cp, _cp],
    "interp2D_fast": [_ci, _ci, _cd, _cd, _ci, _cp, _cp, _cp, _cp],
    "interp2D_fast_layers": [_ci, _ci, _ci, _cd, _cd, _ci, _cp, _cp, _cp, _cp],
    "interp3D_fast": [_ci, _ci, _ci, _cd, _cd, _cd, _ci, _cp, _cp, _cp, _cp, _cp],
    "interp3D_fast_layers": [_ci, _ci, _ci, _ci, _cd, _cd, _cd, _ci, _cp, _cp, _cp, _cp, _cp],
}


def build(force: bool = False) -> str:
    """Compile the C restatement (and oracle/_ref when gfortran exists)."""
    src = os.path.join(_HERE, "fastcompute_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib(prefer_ref: bool = False):
    """ctypes handle of the oracle library (``prefer_ref``: the gfortran build if present)."""
    global _lib
    path = _REF_SO if (prefer_ref and os.path.exists(_REF_SO)) else build()
    if _lib is None or _lib._name != path:
        L = ct.CDLL(path)
        for name, sig in _SIGS.items():
            getattr(L, name).argtypes = sig
            getattr(L, name).restype = None
        if hasattr(L, "oi_interplevel"):
            L.oi_interplevel.argtypes = [_cp, _cp, _ci, _ci, _ci, _cp, _ci, _ci, _cp]
            L.oi_interplevel.restype = None
        _lib = L
    return _lib


def kind() -> str:
    return "reference" if os.path.exists(_REF_SO) else "port"


def _p(a):
    return a.ctypes.data_as(_cp)


class DimensionError(Exception):  # meteotools/exceptions.py:2
    pass


# ---- restatement of meteotools/interpolation/check_arrays.py -------------------------------
def _check_range(loc, d, length, name):  # check_arrays.py:5-31
    if np.nanmax(loc) > d * (length - 1) or np.nanmin(loc) < 0:
        raise ValueError(f"{name} target outside 0 ~ {d * (length - 1)}")


def _to_sentinel(v):  # check_arrays.py:43-44 (+ f64 cast: NumPy-2 hazard)
    v = np.asarray(v, dtype=np.float64)
    return np.where(np.isnan(v), SENTINEL, v)


def _from_sentinel(v):  # check_arrays.py:47-48
    return np.where(v == SENTINEL, np.nan, v)


def _f64_f(data):  # interp_2d.py:131-136: cast + Fortran-order copy
    data = np.asarray(data)
    if data.dtype != np.float64:
        data = data.astype(np.float64)
    return data if data.flags.f_contiguous else np.asfortranarray(data)


def _flat_layers(data, ndim_core):  # check_arrays.py:153-188
    lead = data.shape[:-ndim_core]
    layers = int(np.prod(lead)) if lead else 1
    return lead, data.reshape((layers,) + data.shape[-ndim_core:])


# ---- interp 1-D ---------------------------------------------------------------------------------
def interp1D_fast(dx, xLoc, data):  # interp_1d.py:43-86
    xLoc = np.array(xLoc)
    data = np.asarray(data)
    _check_range(xLoc, dx, data.shape[-1], "X")
    x = _to_sentinel(xLoc)
    if data.ndim != 1:
        raise DimensionError("data must be 1-D")
    shp = x.shape
    x = np.ascontiguousarray(x.reshape(-1))
    d = np.ascontiguousarray(data, dtype=np.float64)
    out = np.empty(x.shape[0])
    lib().interp1D_fast(d.shape[0], float(dx), x.shape[0], _p(x), _p(d), _p(out))
    return _from_sentinel(out.reshape(shp))


def interp1D_fast_layers(dx, xLoc, data):  # interp_1d.py:130-181
    xLoc = np.array(xLoc)
    data = np.asarray(data)
    _check_range(xLoc, dx, data.shape[-1], "X")
    x = _to_sentinel(xLoc)
    lead, d = _flat_layers(data, 1)
    shp = x.shape
    x = np.ascontiguousarray(x.reshape(-1))
    d = _f64_f(d)
    out = np.asfortranarray(np.empty([d.shape[0], x.shape[0]]))
    lib().interp1D_fast_layers(d.shape[1], d.shape[0], float(dx), x.shape[0], _p(x), _p(d), _p(out))
    return _from_sentinel(out.reshape(lead + shp))


def _check_nonequal(x_input, x_output):  # check_arrays.py:51-80
    diff = x_input[1:] - x_input[:-1]
    if not (np.all(diff > 0) or np.all(diff < 0)):
        raise ValueError("x_input must be strictly monotone")
    if np.nanmax(x_output) > np.nanmax(x_input) or np.nanmin(x_output) < np.nanmin(x_input):
        raise ValueError("x_output outside x_input range")


def interp1D_nonequal(x_input, x_output, data):  # interp_1d.py:89-127
    x_output = np.array(x_output)
    x_input = np.array(x_input)
    data = np.asarray(data)
    _check_nonequal(x_input, x_output)
    xo = _to_sentinel(x_output)
    if data.ndim != 1:
        raise DimensionError("data must be 1-D")
    shp = xo.shape
    xo = np.ascontiguousarray(xo.reshape(-1))
    d = np.ascontiguousarray(data, dtype=np.float64)
    xi = np.ascontiguousarray(x_input, dtype=np.float64)
    out = np.empty(xo.shape[0])
    lib().interp1D_nonequal(d.shape[0], _p(xi), xo.shape[0], _p(xo), _p(d), _p(out))
    return _from_sentinel(out.reshape(shp))


def interp1D_nonequal_layers(x_input, x_output, data):  # interp_1d.py:184-241
    x_output = np.array(x_output)
    x_input = np.array(x_input)
    data = np.asarray(data)
    _check_nonequal(x_input, x_output)
    xo = _to_sentinel(x_output)
    lead, d = _flat_layers(data, 1)
    shp = xo.shape
    xo = np.ascontiguousarray(xo.reshape(-1))
    d = _f64_f(d)
    xi = np.ascontiguousarray(x_input, dtype=np.float64)
    out = np.asfortranarray(np.empty([d.shape[0], xo.shape[0]]))
    lib().interp1D_nonequal_layers(d.shape[1], _p(xi), d.shape[0], xo.shape[0], _p(xo), _p(d), _p(out))
    return _from_sentinel(out.reshape(lead + shp))


# ---- interp 2-D ---------------------------------------------------------------------------------
def _prep2(dx, dy, xLoc, yLoc, data):
    xLoc, yLoc = np.array(xLoc), np.array(yLoc)
    if xLoc.shape != yLoc.shape:
        raise DimensionError("xLoc / yLoc shapes differ")
    _check_range(xLoc, dx, data.shape[-1], "X")
    _check_range(yLoc, dy, data.shape[-2], "Y")
    return _to_sentinel(xLoc), _to_sentinel(yLoc)


def interp2D_fast(dx, dy, xLoc, yLoc, data):  # interp_2d.py:33-86
    data = np.asarray(data)
    x, y = _prep2(dx, dy, xLoc, yLoc, data)
    if data.ndim != 2:
        raise DimensionError("data must be 2-D")
    shp = x.shape
    x, y = np.ascontiguousarray(x.reshape(-1)), np.ascontiguousarray(y.reshape(-1))
    d = _f64_f(data)
    out = np.empty(x.shape[0])
    lib().interp2D_fast(d.shape[1], d.shape[0], float(dx), float(dy), x.shape[0], _p(x), _p(y), _p(d), _p(out))
    return _from_sentinel(out.reshape(shp))


def interp2D_fast_layers(dx, dy, xLoc, yLoc, data, return_bare_seconds=None):  # interp_2d.py:89-150
    data = np.asarray(data)
    x, y = _prep2(dx, dy, xLoc, yLoc, data)
    lead, d = _flat_layers(data, 2)
    shp = x.shape
    x, y = np.ascontiguousarray(x.reshape(-1)), np.ascontiguousarray(y.reshape(-1))
    d = _f64_f(d)
    out = np.asfortranarray(np.empty([d.shape[0], x.shape[0]]))
    if return_bare_seconds is not None:
        import time
        t0 = time.perf_counter()
    lib().interp2D_fast_layers(d.shape[2], d.shape[1], d.shape[0], float(dx), float(dy), x.shape[0],
                               _p(x), _p(y), _p(d), _p(out))
    if return_bare_seconds is not None:
        return_bare_seconds.append(time.perf_counter() - t0)
    return _from_sentinel(out.reshape(lead + shp))


# ---- interp 3-D ---------------------------------------------------------------------------------
def _prep3(dx, dy, dz, xLoc, yLoc, zLoc, data):
    xLoc, yLoc, zLoc = np.array(xLoc), np.array(yLoc), np.array(zLoc)
    if xLoc.shape != yLoc.shape or zLoc.shape != yLoc.shape:
        raise DimensionError("xLoc / yLoc / zLoc shapes differ")
    _check_range(xLoc, dx, data.shape[-1], "X")
    _check_range(yLoc, dy, data.shape[-2], "Y")
    _check_range(zLoc, dz, data.shape[-3], "Z")
    return _to_sentinel(xLoc), _to_sentinel(yLoc), _to_sentinel(zLoc)


def interp3D_fast(dx, dy, dz, xLoc, yLoc, zLoc, data):  # interp_3d.py:35-95
    data = np.asarray(data)
    x, y, z = _prep3(dx, dy, dz, xLoc, yLoc, zLoc, data)
    if data.ndim != 3:
        raise DimensionError("data must be 3-D")
    shp = x.shape
    x, y, z = (np.ascontiguousarray(v.reshape(-1)) for v in (x, y, z))
    d = _f64_f(data)
    out = np.empty(x.shape[0])
    lib().interp3D_fast(d.shape[2], d.shape[1], d.shape[0], float(dx), float(dy), float(dz), x.shape[0],
                        _p(x), _p(y), _p(z), _p(d), _p(out))
    return _from_sentinel(out.reshape(shp))


def interp3D_fast_layers(dx, dy, dz, xLoc, yLoc, zLoc, data):  # interp_3d.py:98-163
    data = np.asarray(data)
    x, y, z = _prep3(dx, dy, dz, xLoc, yLoc, zLoc, data)
    lead, d = _flat_layers(data, 3)
    shp = x.shape
    x, y, z = (np.ascontiguousarray(v.reshape(-1)) for v in (x, y, z))
    d = _f64_f(d)
    out = np.asfortranarray(np.empty([d.shape[0], x.shape[0]]))
    lib().interp3D_fast_layers(d.shape[3], d.shape[2], d.shape[1], d.shape[0], float(dx), float(dy),
                               float(dz), x.shape[0], _p(x), _p(y), _p(z), _p(d), _p(out))
    return _from_sentinel(out.reshape(lead + shp))


# ---- vertical interpolation (wrf.interplevel; parity unpinned) -----------------------------
def interplevel(field3d, vert, levels, inclusive: bool = False):
    """``np.array(wrf.interplevel(field3d, vert, levels))`` as the reference's grids see it
    (cylindrical_grid.py:113, cartesian_grid.py:104): NaN where no bracket exists; the result
    carries the dtype of ``field3d`` (wrf-python casts back), computed in float64."""
    field3d, vert = np.asarray(field3d), np.asarray(vert)
    in_dtype = field3d.dtype if field3d.dtype in (np.float32, np.float64) else np.float64
    f = np.ascontiguousarray(field3d, dtype=np.float64)
    v = np.ascontiguousarray(vert, dtype=np.float64)
    lv = np.ascontiguousarray(np.atleast_1d(levels), dtype=np.float64)
    nz, ny, nx = f.shape
    out = np.empty((lv.shape[0], ny, nx))
    lib().oi_interplevel(_p(f), _p(v), nz, ny, nx, _p(lv), lv.shape[0], int(inclusive), _p(out))
    return out.astype(in_dtype, copy=False)


def spline_levels(var, vert, obj_vert, order=2):
    """``cylindrical_grid.spline`` (grids/cylindrical_grid.py:83-93, k = 2) /
    ``cartesian_grid.spline`` (grids/cartesian_grid.py:61-74, k = order): per column the
    interpolating spline of scipy (FITPACK ``curfit`` with s = 0) evaluated at ``obj_vert``
    (``splev``, ext = 0: extrapolates).  Result float64 (the reference fills ``np.zeros``)."""
    from scipy.interpolate import splev, splrep
    var, vert = np.asarray(var), np.asarray(vert)
    ny, nx = var.shape[1], var.shape[2]
    out = np.zeros([len(obj_vert), ny, nx])
    for j in range(ny):
        for i in range(nx):
            cs = splrep(vert[:, j, i], var[:, j, i], k=order)
            out[:, j, i] = splev(obj_vert, cs)
    return out
