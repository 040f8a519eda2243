"""Device-buffer plumbing.  PyTorch is used ONLY as the owner of device memory and streams
(``torch.empty(..., device='cuda')``, ``tensor.data_ptr()``, ``torch.cuda.current_stream()``);
all arithmetic happens in lib/fastcompute.so.  Nothing here computes on the CPU or with torch
operators -- if CUDA is unavailable the functions raise ``BackendError``.
"""
from __future__ import annotations

import ctypes as ct
import sys

import numpy as np

from . import _lib
from .exceptions import BackendError

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise BackendError("meteotools_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return t


def stream():
    return ct.c_void_p(torch().cuda.current_stream().cuda_stream)


def ptr(t):
    return ct.c_void_p(t.data_ptr()) if t is not None else ct.c_void_p(0)


def is_tensor(x):
    # torch may have been imported by the caller only (never through torch() above)
    t = _torch or sys.modules.get("torch")
    return t is not None and isinstance(x, t.Tensor)


def empty(shape, dtype=None):
    t = require_cuda()
    return t.empty(tuple(int(s) for s in shape), dtype=dtype or t.float64, device="cuda")


def to_device(a, dtype=np.float64):
    """Host array (any float dtype / layout) or CUDA tensor -> contiguous CUDA tensor of `dtype`.
    float32 host data is uploaded as float32 (half the PCIe bytes) and widened on the device
    (exact); other dtypes are converted on the host first."""
    t = require_cuda()
    if is_tensor(a):
        if not a.is_cuda:
            a = a.cuda()
        if a.dtype != t.float64 and dtype == np.float64:
            return cast_f64(a.contiguous())
        return a.contiguous()
    a = np.asarray(a)
    if dtype == np.float64 and a.dtype == np.float32:
        return cast_f64(_from_numpy(np.ascontiguousarray(a)).cuda())
    return _from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def _from_numpy(a):
    """torch.from_numpy without the 'array is not writable' warning: grid fields are handed out
    as read-only views on purpose and are only ever READ here (H2D copy)."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch().from_numpy(a)


def cast_f64(x):
    """f32 CUDA tensor -> f64 CUDA tensor through mt_cast_f32_f64 (exact widening)."""
    t = torch()
    if x.dtype == t.float64:
        return x
    if x.dtype != t.float32:
        raise BackendError(f"unsupported device dtype {x.dtype}: pass float64 or float32 tensors")
    out = t.empty(x.shape, dtype=t.float64, device=x.device)
    _lib.check(_lib.load().mt_cast_f32_f64(ptr(x), ptr(out), x.numel(), stream()), "mt_cast_f32_f64")
    return out


def scale_box(box, fac, org_y, org_x, width):
    """box (nz, by, pitch) *= fac[org_y:org_y+by, org_x:org_x+width] in place (mt_scale_box_2d): the map-factor
    scaling of wrfout_grid.correct_map_scale on a staged source box, in the box's own dtype."""
    t = torch()
    if not is_tensor(fac):   # keep a float32 factor float32: float32 *= float32 is a float product in NumPy
        fac = np.asarray(fac)
        fac = np.ascontiguousarray(fac, dtype=fac.dtype if fac.dtype in (np.float32, np.float64) else np.float64)
        fac = _from_numpy(fac).cuda()
    else:
        fac = fac.contiguous() if fac.dtype in (t.float32, t.float64) else fac.to(t.float64).contiguous()
    nz, by, pitch = (int(v) for v in box.shape)
    code = lambda x: _lib.MT_F32 if x.dtype == t.float32 else _lib.MT_F64  # noqa: E731
    _lib.check(_lib.load().mt_scale_box_2d(ptr(box), code(box), nz, by, int(width), pitch, ptr(fac), code(fac),
                                           int(fac.shape[-1]), int(org_y), int(org_x), stream()), "mt_scale_box_2d")
    return box


def to_host(x):
    """Device tensor -> NumPy array backed by PINNED host memory (torch's caching host allocator
    recycles the blocks, so repeated downloads run at PCIe speed without re-pinning)."""
    t = torch()
    h = t.empty(x.shape, dtype=x.dtype, pin_memory=True)
    h.copy_(x, non_blocking=True)
    t.cuda.current_stream().synchronize()
    return h.numpy()


def box_pitch(width, itemsize=None):
    """Row pitch (elements) of a staged box: a multiple of 4 elements, so that rows of float32 and
    float64 boxes alike start 16-byte aligned and the TMA path of set_data (mt_setdata_tma) can read
    the staging buffer through a tensor map; one pitch for both dtypes keeps mixed batches aligned."""
    return (int(width) + 3) // 4 * 4


def upload_box(a, y0, y1, x0, x1, out=None):
    """Column box of a C-order (nz,ny,nx) host array -> device tensor (nz, y1-y0, pitch) of the
    same dtype, pitch = box_pitch(x1-x0) (valid columns first, the padding is never read), with
    one strided DMA (mt_upload_box_pitched); no host-side crop copy."""
    t = require_cuda()
    if is_tensor(a):   # a resident source: device-to-device strided copy of its box
        a = a.contiguous()
        if a.dtype not in (t.float32, t.float64):
            a = cast_f64(a)
        f32, itemsize, src = a.dtype == t.float32, a.element_size(), ct.c_void_p(a.data_ptr())
    else:
        a = np.asarray(a)
        if a.dtype not in (np.float32, np.float64) or not a.flags.c_contiguous:
            a = np.ascontiguousarray(a, dtype=np.float64)
        f32, itemsize, src = a.dtype == np.float32, a.dtype.itemsize, ct.c_void_p(a.ctypes.data)
    nz, ny, nx = (int(v) for v in a.shape)
    pitch = box_pitch(x1 - x0, itemsize)
    if out is None:
        out = t.empty((nz, y1 - y0, pitch), dtype=t.float32 if f32 else t.float64, device="cuda")
    assert tuple(out.shape) == (nz, y1 - y0, pitch), (tuple(out.shape), (nz, y1 - y0, pitch))
    _lib.check(_lib.load().mt_upload_box_pitched(src, itemsize, nz, ny, nx, y0, y1,
                                                 x0, x1, ptr(out), pitch, stream()), "mt_upload_box_pitched")
    out._mt_keepalive = a   # the async DMA reads `a` until the stream reaches this point
    return out


def permute3(src, shape, strides):
    """Dense tensor dst of `shape` with dst[i,j,k] = src.flat[i*s0 + j*s1 + k*s2]."""
    dst = empty(shape)
    n0, n1, n2 = (int(s) for s in shape)
    s0, s1, s2 = (int(s) for s in strides)
    _lib.check(_lib.load().mt_permute3(ptr(src), ptr(dst), n0, n1, n2, s0, s1, s2, stream()),
               "mt_permute3")
    return dst


def ptr_array(tensors):
    """Host array of device pointers (const double *const *) for the multi-variable entry points."""
    arr = (ct.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr
