"""Helpers of meteotools/calc/check_arrays.py that the hot path needs."""
import numpy as np

from ..exceptions import DimensionError


def check_dimension(variable, number_of_dim):  # calc/check_arrays.py:6-8
    if len(variable.shape) != number_of_dim:
        raise DimensionError(f'Cartesian data (car_data) must be {number_of_dim:d}-dimensional')


def split_axis(shape, axis):
    """(outer, n, inner) view of a dense C-order array around ``axis`` -- the device analogue of
    the reference's np.moveaxis(axis -> 0) round trip (calc/check_arrays.py:52-63)."""
    ndim = len(shape)
    if not -ndim <= axis < ndim:
        raise np.exceptions.AxisError(axis, ndim)
    axis %= ndim
    outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64)) if axis < ndim - 1 else 1
    return outer, int(shape[axis]), inner
