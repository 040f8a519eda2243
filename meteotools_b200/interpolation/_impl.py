"""Implementation of the eight public interpolation functions on top of the mt_* device API.

Reference counterparts: meteotools/interpolation/interp_1d.py:43-241, interp_2d.py:33-150,
interp_3d.py:35-163.  ``data`` may be a NumPy array (any float dtype, any layout) -- it is
uploaded, the kernel runs, the result comes back as a NumPy float64 array -- or a CUDA
``torch.Tensor`` (float64/float32, C-order), in which case the result stays on the device.
The reference's full transposing host copy (np.asfortranarray, interp_2d.py:135-136) does not
exist here: the kernels take element strides.
"""
import numpy as np

from .. import _device as D
from .. import _lib
from ..exceptions import DimensionError
from . import check_arrays as C


def _data_shape(data):
    return tuple(data.shape)


def _as_data(data):
    return data if D.is_tensor(data) else np.asarray(data)


def _finish(out_dev, shape, want_tensor):
    out_dev = out_dev.reshape(tuple(shape))
    return out_dev if want_tensor else D.to_host(out_dev)


def _run1d(dx, xLoc, data, layers_mode):
    data = _as_data(data)
    (x,) = C.as_targets(xLoc)
    shape = _data_shape(data)
    C.check_destination(x, dx, shape[-1], "X")
    if not layers_mode and len(shape) != 1:
        raise DimensionError("source data must be 1-D")
    layers, lead = C.flatten_layers(shape, 1)
    want_tensor = D.is_tensor(data)
    d = D.to_device(data)
    xd = D.to_device(x.reshape(-1))
    n, lenx = xd.numel(), shape[-1]
    out = D.empty((layers, n))
    if n and layers:
        _lib.check(_lib.load().mt_interp1d_layers(D.ptr(d), layers, lenx, lenx, 1, float(dx), D.ptr(xd), n,
                                                  D.ptr(out), n, 1, D.stream()), "mt_interp1d_layers")
    return _finish(out, lead + x.shape, want_tensor)


def interp1D_fast(dx, xLoc, data):
    """Linear interpolation on an equally spaced 1-D axis (interp_1d.py:43-86)."""
    return _run1d(dx, xLoc, data, False)


def interp1D_fast_layers(dx, xLoc, data):
    """Same, repeated over all leading dimensions of ``data`` (interp_1d.py:130-181)."""
    return _run1d(dx, xLoc, data, True)


def _run1d_nonequal(x_input, x_output, data, layers_mode):
    data = _as_data(data)
    (xo,) = C.as_targets(x_output)
    xi = np.array(x_input)
    C.check_monotonically(xi)
    C.check_destination_nonequal(xo, xi)
    shape = _data_shape(data)
    if not layers_mode and len(shape) != 1:
        raise DimensionError("source data must be 1-D")
    layers, lead = C.flatten_layers(shape, 1)
    want_tensor = D.is_tensor(data)
    d = D.to_device(data)
    xid = D.to_device(xi.astype(np.float64))
    xod = D.to_device(xo.reshape(-1))
    n, lenx = xod.numel(), shape[-1]
    out = D.empty((layers, n))
    if n and layers:
        _lib.check(_lib.load().mt_interp1d_nonequal_layers(D.ptr(xid), D.ptr(d), layers, lenx, lenx, 1,
                                                           D.ptr(xod), n, D.ptr(out), n, 1, D.stream()),
                   "mt_interp1d_nonequal_layers")
    return _finish(out, lead + xo.shape, want_tensor)


def interp1D_nonequal(x_input, x_output, data):
    """Linear interpolation on a monotone, unequally spaced axis (interp_1d.py:89-127)."""
    return _run1d_nonequal(x_input, x_output, data, False)


def interp1D_nonequal_layers(x_input, x_output, data):
    """Same, repeated over the leading dimensions (interp_1d.py:184-241)."""
    return _run1d_nonequal(x_input, x_output, data, True)


def _run2d(dx, dy, xLoc, yLoc, data, layers_mode):
    data = _as_data(data)
    x, y = C.as_targets(xLoc, yLoc)
    shape = _data_shape(data)
    C.check_destination(x, dx, shape[-1], "X")
    C.check_destination(y, dy, shape[-2], "Y")
    if not layers_mode and len(shape) != 2:
        raise DimensionError("source data must be 2-D")
    layers, lead = C.flatten_layers(shape, 2)
    want_tensor = D.is_tensor(data)
    d = D.to_device(data)
    xd, yd = D.to_device(x.reshape(-1)), D.to_device(y.reshape(-1))
    n, leny, lenx = xd.numel(), shape[-2], shape[-1]
    out = D.empty((layers, n))
    if n and layers:
        _lib.check(_lib.load().mt_interp2d_layers(D.ptr(d), layers, leny, lenx, leny * lenx, lenx, 1,
                                                  float(dx), float(dy), D.ptr(xd), D.ptr(yd), n, D.ptr(out),
                                                  n, 1, D.stream()), "mt_interp2d_layers")
    return _finish(out, lead + x.shape, want_tensor)


def interp2D_fast(dx, dy, xLoc, yLoc, data):
    """Bilinear interpolation of one equally spaced 2-D field (interp_2d.py:33-86)."""
    return _run2d(dx, dy, xLoc, yLoc, data, False)


def interp2D_fast_layers(dx, dy, xLoc, yLoc, data):
    """Bilinear interpolation repeated over all leading dimensions (interp_2d.py:89-150):
    dim(data)=(a,b,e,f), dim(xLoc)=dim(yLoc)=(g,h) -> dim(out)=(a,b,g,h)."""
    return _run2d(dx, dy, xLoc, yLoc, data, True)


def _run3d(dx, dy, dz, xLoc, yLoc, zLoc, data, layers_mode):
    data = _as_data(data)
    x, y, z = C.as_targets(xLoc, yLoc, zLoc)
    shape = _data_shape(data)
    C.check_destination(x, dx, shape[-1], "X")
    C.check_destination(y, dy, shape[-2], "Y")
    C.check_destination(z, dz, shape[-3], "Z")
    if not layers_mode and len(shape) != 3:
        raise DimensionError("source data must be 3-D")
    layers, lead = C.flatten_layers(shape, 3)
    want_tensor = D.is_tensor(data)
    d = D.to_device(data)
    xd, yd, zd = (D.to_device(v.reshape(-1)) for v in (x, y, z))
    n, lenz, leny, lenx = xd.numel(), shape[-3], shape[-2], shape[-1]
    out = D.empty((layers, n))
    if n and layers:
        _lib.check(_lib.load().mt_interp3d_layers(D.ptr(d), layers, lenz, leny, lenx, lenz * leny * lenx,
                                                  leny * lenx, lenx, 1, float(dx), float(dy), float(dz),
                                                  D.ptr(xd), D.ptr(yd), D.ptr(zd), n, D.ptr(out), n, 1,
                                                  D.stream()), "mt_interp3d_layers")
    return _finish(out, lead + x.shape, want_tensor)


def interp3D_fast(dx, dy, dz, xLoc, yLoc, zLoc, data):
    """Trilinear interpolation of one equally spaced 3-D field (interp_3d.py:35-95)."""
    return _run3d(dx, dy, dz, xLoc, yLoc, zLoc, data, False)


def interp3D_fast_layers(dx, dy, dz, xLoc, yLoc, zLoc, data):
    """Trilinear interpolation repeated over the leading dimensions (interp_3d.py:98-163)."""
    return _run3d(dx, dy, dz, xLoc, yLoc, zLoc, data, True)
