"""Range averages -- same functions as meteotools/calc/averages.py:13-114, reduced on the device
by ``mt_axis_wsum`` (csrc/reduce.cu).  ``var`` is a NumPy array or a CUDA tensor (dense, C order);
the (small) result is always a NumPy array.

Semantics kept from the reference: the averaging axis must increase strictly (ValueError) and
match ``var``'s extent (DimensionError); NaN limits mean "from the first / to the last point";
NaNs in ``var`` are skipped in the SUM but calc_Zaverage still divides by the number of SELECTED
points (:113) and calc_Raverage by ``nansum(select*r)`` (:81-82), exactly as written there.
The additions follow NumPy's order for a C-order array (sequential along strided axes, pairwise
blocks along the last axis): the results are bit-identical to the reference's.
"""
import numpy as np

from .. import _device as D
from .. import _lib
from ..exceptions import DimensionError


def convert_nan_to_default_value(axis, min_value, max_value):  # averages.py:5-10
    if np.isnan(min_value):
        min_value = axis[0]
    if np.isnan(max_value):
        max_value = axis[-1]
    return min_value, max_value


def _check_axis(var_axis, var, axis, name):  # check_arrays.py:11-14, :33-38
    var_axis = np.array(var_axis)
    if np.min(var_axis[1:] - var_axis[:-1]) < 0:
        raise ValueError(f"{name} axis must increase strictly.")
    if var.shape[axis] != var_axis.shape[0]:
        raise DimensionError(f"The shape of axis ({var_axis.shape[0]:d}) and the shape of averaged dimension "
                             f"of var ({var.shape[axis]:d}) are not the same.")
    return var_axis


def axis_wsum(var, axis, w, mode=0, denom=1.0):
    """Device reduction of a dense C-order array along ``axis``: see mt_axis_wsum.  Returns a CUDA
    tensor with that axis removed.  The terms are added in NumPy's order for a C-order array --
    pairwise along the last (contiguous) axis, sequentially along the others -- so the result is
    bit-identical to the reference's ``np.nansum(var * w, axis)``."""
    v = D.to_device(var)
    shape = tuple(int(s) for s in v.shape)
    nd = len(shape)
    axis %= nd
    outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64)) if axis < nd - 1 else 1
    n = shape[axis]
    wd = D.to_device(np.ascontiguousarray(w, dtype=np.float64))
    out = D.empty(shape[:axis] + shape[axis + 1:] if nd > 1 else (1,))
    if axis == nd - 1:
        mode = int(mode) | _lib.MT_WSUM_PAIRWISE
    _lib.check(_lib.load().mt_axis_wsum(D.ptr(v), outer, inner, n, n * inner, 1, inner, D.ptr(wd), int(mode),
                                        float(denom), D.ptr(out), inner, 1, D.stream()), "mt_axis_wsum")
    return out


def calc_Zaverage(var, z_axis, axis, zmin=np.nan, zmax=np.nan):
    """Average over [zmin, zmax] of ``z_axis`` along ``axis`` (averages.py:87-114):
    nansum(var*select)/count(select)."""
    z_axis = _check_axis(z_axis, var, axis, 'input')
    zmin, zmax = convert_nan_to_default_value(z_axis, zmin, zmax)
    z_select = np.logical_and(z_axis >= zmin, z_axis <= zmax)
    with np.errstate(all="ignore"):
        denom = float(np.sum(z_select))
    out = D.to_host(axis_wsum(var, axis, z_select.astype(np.float64), 0, denom))
    return out if len(var.shape) > 1 else out[0]


def calc_Raverage(var, r_axis, axis, rmin=np.nan, rmax=np.nan):
    """Radially weighted average over [rmin, rmax] along ``axis`` (averages.py:56-84):
    nansum(var*select*r)/nansum(select*r)."""
    r_axis = _check_axis(r_axis, var, axis, 'input')
    rmin, rmax = convert_nan_to_default_value(r_axis, rmin, rmax)
    r_select = np.logical_and(r_axis >= rmin, r_axis <= rmax)
    with np.errstate(all="ignore"):
        denom = float(np.nansum(r_select * r_axis))
    out = D.to_host(axis_wsum(var, axis, r_select * r_axis, 0, denom))
    return out if len(var.shape) > 1 else out[0]


def calc_RZaverage(var, z_axis, r_axis, rmin=np.nan, rmax=np.nan, zmin=np.nan, zmax=np.nan, r_weight=True):
    """Average of a (z, r) array over a z range (nanmean) and an r range (radially weighted unless
    ``r_weight`` is false) (averages.py:13-53).  Returns a float."""
    z_axis = _check_axis(z_axis, var, 0, 'z')
    r_axis = _check_axis(r_axis, var, 1, 'r')
    rmin, rmax = convert_nan_to_default_value(r_axis, rmin, rmax)
    zmin, zmax = convert_nan_to_default_value(z_axis, zmin, zmax)
    z_select = np.logical_and(z_axis >= zmin, z_axis <= zmax)
    r_select = np.logical_and(r_axis >= rmin, r_axis <= rmax)
    tmp = axis_wsum(var, 0, z_select.astype(np.float64), 1)          # nanmean over the selected rows
    if r_weight == True:  # noqa: E712  (the reference's test, :49)
        with np.errstate(all="ignore"):
            denom = float(np.nansum(r_select * r_axis))
        return float(D.to_host(axis_wsum(tmp, 0, r_select * r_axis, 0, denom))[0])
    # np.nanmean(tmp[r_select]) (:51): the selection is compacted first, so the pairwise blocks of
    # the sum start at the first selected radius (a range of a monotone axis is one run)
    idx = np.flatnonzero(r_select)
    if idx.size == 0:
        return float('nan')
    run = tmp[int(idx[0]):int(idx[-1]) + 1]
    return float(D.to_host(axis_wsum(run, 0, np.ones(run.shape[0]), 1))[0])
