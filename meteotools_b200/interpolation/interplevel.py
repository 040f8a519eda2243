"""``interplevel(field3d, vert, desiredlev)`` -- vertical interpolation of a 3-D field to constant
height / pressure levels with the semantics of ``wrf.interplevel`` (wrf-python 1.3.x, DINTERP3DZ),
the third-party routine the reference calls at ``grids/cylindrical_grid.py:113`` and
``grids/cartesian_grid.py:104`` (parity unpinned).

Provided so that scripts that call ``wrf.interplevel`` next to meteotools keep one backend:

    from meteotools_b200.interpolation import interplevel
    T_850 = interplevel(T, p, [85000.0])          # (1, ny, nx); NaN where no bracket exists

The direction of the coordinate is decided from column (0, 0) as DINTERP3DZ does; a level that has
no bracketing pair of model levels gives NaN (wrf-python's masked ``missing`` seen through
``np.array``).  ``inclusive=True`` selects ``<=`` instead of the strict ``<`` in the bracket test.
float32 input gives a float32 result (the value rounded exactly where wrf-python rounds it); CUDA
tensors are taken in place and a float64 CUDA tensor is returned.
"""
import numpy as np

from .. import _device as D
from .. import _lib
from ..exceptions import DimensionError


def interplevel(field3d, vert, desiredlev, inclusive=False):
    t = D.require_cuda()
    on_dev = D.is_tensor(field3d)
    if D.is_tensor(vert) != on_dev:
        raise TypeError("field3d and vert must both be host arrays or both CUDA tensors")
    if len(field3d.shape) != 3 or tuple(field3d.shape) != tuple(vert.shape):
        raise DimensionError("field3d and vert must be 3-D arrays of the same shape (nz, ny, nx)")
    nz, ny, nx = (int(s) for s in field3d.shape)
    if nz < 2:
        raise DimensionError("at least two model levels are needed")
    levels = np.atleast_1d(np.asarray(desiredlev, dtype=np.float64))
    if levels.ndim != 1:
        raise DimensionError("desiredlev must be a scalar or a 1-D sequence of levels")

    def is_f32(a):
        return a.dtype == (t.float32 if on_dev else np.float32)
    f32 = is_f32(field3d)
    use32 = f32 and is_f32(vert)
    dtype = _lib.MT_F32 if use32 else _lib.MT_F64

    def stage(a):
        if on_dev:
            a = a.contiguous()
            return a if (use32 and a.dtype == t.float32) else D.cast_f64(a)
        a = np.ascontiguousarray(a)
        if use32:
            return t.from_numpy(a).cuda()
        return D.to_device(a)
    fd, vd = stage(field3d), stage(vert)
    lv = D.to_device(levels)
    order = 1 if np.all(np.diff(levels) > 0) else (-1 if np.all(np.diff(levels) < 0) else 0)
    if levels.size == 1:
        order = 1
    out = D.empty((levels.size, ny, nx))
    flags = (_lib.MT_IL_INCLUSIVE if inclusive else 0) | (_lib.MT_IL_ROUND_F32 if f32 else 0)
    _lib.check(_lib.load().mt_interplevel(D.ptr_array([fd]), 1, D.ptr(vd), dtype, nz, ny, nx, 0, ny, 0, nx,
                                          D.ptr(lv), levels.size, order, -1, D.ptr_array([out]), ny * nx, nx, 1,
                                          flags, D.stream()), "mt_interplevel")
    if on_dev:     # device result: float64 tensor (holding float32-rounded values for float32 input)
        return out
    res = D.to_host(out)
    return res.astype(np.float32) if f32 else res
