"""Argument validation shared by the interpolation entry points.

Same contract as the reference's ``meteotools/interpolation/check_arrays.py``:
* targets outside ``[0, d*(len-1)]`` -> ``ValueError``                      (:5-31)
* mismatching target shapes          -> ``DimensionError``                  (:38-40)
* unequal axis not strictly monotone / targets outside it -> ``ValueError`` (:51-80)
* NaN targets are legal and give NaN outputs                                (:43-48)
* leading dimensions of ``data`` are flattened into ``layers``              (:153-188)

Unlike the reference, targets are always converted to float64 first (the reference forgets to,
which breaks float32 targets under NumPy >= 2; SURVEY 8b "NumPy-2 hazard"), and NaN targets are
passed to the device as NaN (the kernels answer NaN directly; the legacy C symbols still honour
the 1.7e308 sentinel).
"""
import numpy as np

from ..exceptions import DimensionError


def as_targets(*locs):
    out = [np.array(v, dtype=np.float64) for v in locs]
    for other in out[1:]:
        if other.shape != out[0].shape:
            raise DimensionError("xLoc / yLoc / zLoc must have the same shape")
    return out


def check_destination(destination, dx, length, dimension_name):
    if destination.size == 0:
        return
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            dmax, dmin = np.nanmax(destination), np.nanmin(destination)
    upper = dx * (length - 1)
    if dmax > upper:
        raise ValueError(f"{dimension_name} target maximum {dmax} exceeds the grid range 0 ~ {upper}")
    if dmin < 0:
        raise ValueError(f"{dimension_name} target minimum {dmin} exceeds the grid range 0 ~ {upper}")


def check_monotonically(axis):
    diff = axis[1:] - axis[:-1]
    if not (np.all(diff > 0) or np.all(diff < 0)):
        raise ValueError("x_input must be strictly increasing or strictly decreasing")


def check_destination_nonequal(destination, origin):
    dmax, dmin = np.nanmax(destination), np.nanmin(destination)
    omax, omin = np.nanmax(origin), np.nanmin(origin)
    if dmax > omax:
        raise ValueError(f"x_output maximum {dmax} exceeds the data range {omin} ~ {omax}")
    if dmin < omin:
        raise ValueError(f"x_output minimum {dmin} exceeds the data range {omin} ~ {omax}")


def flatten_layers(shape, core_ndim):
    """(lead dims..., core dims) -> (layers, lead shape)."""
    lead = tuple(shape[:-core_ndim])
    layers = 1
    for s in lead:
        layers *= int(s)
    return layers, lead
