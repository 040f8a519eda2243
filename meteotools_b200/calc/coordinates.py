"""Axis builders and the Cartesian -> cylindrical entry point (meteotools/calc/coordinates.py).

Target coordinates are computed ON THE HOST with NumPy, exactly as the reference does
(:136-138): device cos/sin differ from NumPy's in the last ulp, which would flip ``int(x/dx)``
at the exact-integer nodes that the reference's tests pin with ``==`` (test_calc.py:67-78).
"""
import numpy as np

from ..exceptions import InputError, UnitError
from ..interpolation import interp2D_fast_layers
from .check_arrays import check_dimension

_RATIO = {'km': 1000, 'hPa': 100, 'deg': np.pi / 180}


def unit_ratio(in_unit):  # :8-16
    return _RATIO.get(in_unit, 1)


def _axis(start, end, step, eps, in_unit, allowed):
    if in_unit not in allowed:
        raise UnitError('unsupported unit')
    return np.arange(start, end + eps, step, dtype='float64')


def Make_vertical_axis(height_start, height_end, dz, in_unit='m'):  # :19-49
    return _axis(height_start, height_end, dz, 0.00001, in_unit, ('m', 'Pa', 'km', 'hPa')) * unit_ratio(in_unit)


def Make_tangential_axis(theta_start, theta_end, dtheta, in_unit='rad'):  # :52-83
    theta = _axis(theta_start, theta_end, dtheta, 0.00001, in_unit, ('deg', 'rad'))
    ratio = unit_ratio(in_unit)
    if (theta_end - theta_start) * ratio == 2 * np.pi:
        theta = theta[:-1]
    return theta * ratio


def Make_radial_axis(r_start, r_end, dr, in_unit='m'):  # :86-116
    return _axis(r_start, r_end, dr, 0.0001, in_unit, ('m', 'km')) * unit_ratio(in_unit)


def Make_cyclinder_coord(centerLocation, r, theta):
    """x, y of every (theta, r) node for a centre [y, x] (coordinates.py:119-140)."""
    rr, ttheta = np.meshgrid(r, theta)
    return rr * np.cos(ttheta) + centerLocation[1], rr * np.sin(ttheta) + centerLocation[0]


def cartesian2cylindrical(dx, dy, cartesian_data, centerLocation=None, r=None, theta=None,
                          xTR=None, yTR=None):
    """Horizontal interpolation of a (z, y, x) field onto (z, theta, r) (coordinates.py:143-190)."""
    check_dimension(cartesian_data, 3)
    if xTR is None or yTR is None:
        if centerLocation is None or r is None or theta is None:
            raise InputError("either (xTR, yTR) or (centerLocation, r, theta) is required")
        xTR, yTR = Make_cyclinder_coord(centerLocation, r, theta)
    return interp2D_fast_layers(dx, dy, xTR, yTR, cartesian_data)
