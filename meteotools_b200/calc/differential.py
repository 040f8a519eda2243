"""Finite differences -- same functions as meteotools/calc/differential.py:6-180, computed by the
``mt_fd`` kernel (csrc/fd.cu).  ``var`` is a NumPy array (result: NumPy float64) or a CUDA
tensor (result stays on the device).  The formulas and their order of evaluation are the
reference's; results are bit-identical for float64 input.  (float32 input is widened first, the
reference differentiates in float32 and stores to float64 -- SURVEY 8a "dtype".)

``FD2(..., cyclic=True)`` is unreachable in the reference (the wrapper signature is
``(var, delta, axis)``, calc/check_arrays.py:54, so passing ``cyclic`` raises TypeError);
the same signature is kept here.
"""
from .. import _device as D
from .. import _lib
from .check_arrays import split_axis


def _fd(kind, var, delta, axis):
    delta = float(delta)
    want_tensor = D.is_tensor(var)
    v = D.to_device(var)
    shape = tuple(v.shape)
    outer, n, inner = split_axis(shape, axis)
    need = {"FD4": 5, "FD2_2": 4, "FD2_2_front": 4, "FD2_2_back": 4}.get(kind, 3)
    if n < need:
        raise IndexError(f"{kind} needs at least {need} points along axis {axis} (got {n})")
    out = D.empty(shape)
    _lib.check(_lib.load().mt_fd(D.ptr(v), D.ptr(out), outer, n, inner, delta, _lib.FD_KINDS[kind],
                                 D.stream()), f"mt_fd({kind})")
    return out if want_tensor else D.to_host(out)


def FD2(var, delta, axis):
    """2nd-order centred difference, one-sided 2nd-order at both ends (differential.py:6-36)."""
    return _fd("FD2", var, delta, axis)


def FD4(var, delta, axis):
    """4th-order centred interior, 2nd-order near the ends (differential.py:38-74)."""
    return _fd("FD4", var, delta, axis)


def FD2_front(var, delta, axis):
    """2nd-order forward difference; the last two rows are NaN (differential.py:76-94)."""
    return _fd("FD2_front", var, delta, axis)


def FD2_back(var, delta, axis):
    """2nd-order backward difference; the first two rows are NaN (differential.py:97-115)."""
    return _fd("FD2_back", var, delta, axis)


def FD2_2(var, delta, axis):
    """Second derivative, 2nd-order (differential.py:118-138)."""
    return _fd("FD2_2", var, delta, axis)


def FD2_2_front(var, delta, axis):
    """Second derivative, forward; last three rows NaN (differential.py:141-159)."""
    return _fd("FD2_2_front", var, delta, axis)


def FD2_2_back(var, delta, axis):
    """Second derivative, backward; first three rows NaN (differential.py:162-180)."""
    return _fd("FD2_2_back", var, delta, axis)
