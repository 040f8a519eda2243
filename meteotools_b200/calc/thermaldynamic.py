"""Point-wise moist thermodynamics -- the functions of meteotools/calc/thermaldynamic.py
(:16-533) with the same names, optional-argument dispatch and errors, evaluated by the
``mt_pointwise`` / ``mt_calc_tc`` kernels (csrc/thermo.cu).

Arguments may be NumPy arrays, Python floats (mixed freely, broadcast like NumPy) or CUDA
tensors; the result is a NumPy array / float / CUDA tensor accordingly.
"""
import numpy as np

from .. import _device as D
from .. import _lib
from ..exceptions import InputError

# physical constants, thermaldynamic.py:4-13
Rd = 287.  # J/kg*K
Cp = 1004.  # J/kg*K
g = 9.8  # m/s^2
T0 = 250.  # K
P0 = 1000.  # hPa
H = Rd * T0 / g  # m
kappa = Rd / Cp
Lv = 2.5e6
A = 2.53e11
B = 5420.


def _is_scalar(x):
    return x is not None and not D.is_tensor(x) and np.ndim(x) == 0


def _prepare(args):
    """Broadcast array operands to one shape, upload; scalars stay scalars (None stays None)."""
    arrays = [a for a in args if a is not None and not _is_scalar(a)]
    if not arrays:  # all scalars: run on a 1-element device array
        shape, scalar_out = (1,), True
        first = True
        out = []
        for a in args:
            if a is not None and first:
                out.append(D.to_device(np.full(1, float(a))))
                first = False
            else:
                out.append(None if a is None else float(a))
        return out, shape, scalar_out, False
    want_tensor = any(D.is_tensor(a) for a in arrays)
    shapes = {tuple(a.shape) for a in arrays}
    if len(shapes) > 1:
        if want_tensor:
            raise InputError("device tensors of different shapes cannot be broadcast")
        bshape = np.broadcast_shapes(*shapes)
        args = [np.broadcast_to(np.asarray(a), bshape) if (a is not None and not _is_scalar(a)) else a
                for a in args]
        shape = tuple(bshape)
    else:
        shape = next(iter(shapes))
    out = [None if a is None else (float(a) if _is_scalar(a) else D.to_device(a)) for a in args]
    return out, shape, False, want_tensor


def _pw(op, a, b=None, c=None):
    """out = op(a, b, c) on the device; `a` must end up an array operand."""
    (a, b, c), shape, scalar_out, want_tensor = _prepare([a, b, c])
    if not D.is_tensor(a):  # first operand is a scalar but another one is an array: materialise it
        a = D.to_device(np.full(shape, a))
    n = a.numel()
    out = D.empty(shape)
    pb = D.ptr(b) if D.is_tensor(b) else D.ptr(None)
    pc = D.ptr(c) if D.is_tensor(c) else D.ptr(None)
    sb = 0.0 if (b is None or D.is_tensor(b)) else b
    sc = 0.0 if (c is None or D.is_tensor(c)) else c
    if n:
        _lib.check(_lib.load().mt_pointwise(_lib.OPS[op], D.ptr(a), pb, pc, sb, sc, D.ptr(out), n,
                                            D.stream()), f"mt_pointwise({op})")
    if want_tensor:
        return out
    res = D.to_host(out)
    return float(res[0]) if scalar_out else res


def _none(*xs):
    return any(x is None for x in xs)


def calc_H(Tv):
    """Scale height (thermaldynamic.py:16-30) -- one multiply/divide, evaluated where Tv lives."""
    return Rd * Tv / g


def P_to_Z(P0, Tv, P):  # :33-49
    return _pw("P_TO_Z", P, P0, Tv)


def Z_to_P(P0, Tv, Z):  # :51-67
    return _pw("Z_TO_P", Z, P0, Tv)


def calc_theta(T, P):  # :70-84
    return _pw("THETA", T, P)


def calc_T(theta, P):  # :87-101
    return _pw("T", theta, P)


def calc_saturated_vapor(T):  # :104-116
    return _pw("SAT_VAPOR", T)


def calc_vapor(T, RH=1.):  # :119-134
    return _pw("VAPOR", T, RH)


def calc_saturated_qv(T, P):  # :137-151
    return _pw("SAT_QV", T, P)


def qv_to_vapor(P, qv):  # :154-168
    return _pw("QV_TO_VAPOR", P, qv)


def calc_theta_es(T, P):  # :171-189
    return _pw("THETA_ES", T, P)


def calc_Td(es=None, P=None, qv=None, T=None, RH=None, theta=None):  # :192-225
    if es is None:
        if not _none(qv, P):
            es = qv_to_vapor(P, qv)
        elif not _none(T, RH):
            es = calc_vapor(T, RH)
        elif not _none(theta, P, RH):
            es = calc_vapor(calc_T(theta, P), RH)
        else:
            raise InputError('es, (P, qv), (T, RH) or (theta, P, RH) is required')
    return _pw("TD", es)


def calc_qv(P, T=None, RH=None, vapor=None, Td=None):  # :228-256
    if vapor is None:
        if not _none(RH, T):
            vapor = calc_vapor(T, RH)
        elif Td is not None:
            vapor = calc_vapor(Td)
        else:
            raise InputError('(T, RH), Td or vapor is required')
    return _pw("QV", P, vapor)


def calc_theta_v(T=None, P=None, theta=None, qv=None, es=None, RH=None, Td=None, ql=0):  # :259-309
    if theta is None:
        if _none(T, P):
            raise InputError('theta or (T, P) is required')
        theta = calc_theta(T, P)
    if qv is None:
        if not _none(es, P):
            qv = calc_qv(P=P, vapor=es)
        elif not _none(T, RH, P):
            qv = calc_qv(P=P, T=T, RH=RH)
        elif not _none(Td, P):
            qv = calc_qv(P=P, Td=Td)
        else:
            raise InputError('qv, (vapor, P), (T, RH, P) or (Td, P) is required')
    return _pw("THETA_V", theta, qv, ql)


def calc_Tv(T, P=None, RH=None, vapor=None, qv=None):  # :312-337
    if not _none(RH, P):
        return _pw("TV_VAPOR", T, calc_vapor(T, RH), P)
    if not _none(vapor, P):
        return _pw("TV_VAPOR", T, vapor, P)
    if qv is not None:
        return _pw("TV_QV", T, qv)
    raise InputError('(P, RH), (P, vapor) or qv is required')


def calc_rho(P, Tv=None, T=None, RH=None, vapor=None, qv=None):  # :340-364
    if Tv is None:
        Tv = calc_Tv(T, P=P, RH=RH, vapor=vapor, qv=qv)
    return _pw("RHO", P, Tv)


def calc_Tc(T, P, RH=None, vapor=None, qv=None, return_iterations=False):
    """Condensation temperature (:367-402): <= 10 fixed-point sweeps with the reference's global,
    NaN-sensitive stop test ``np.max(|Tc-Tc2|) < 0.1`` evaluated on the device."""
    if qv is None:
        if RH is not None:
            qv = calc_qv(P=P, T=T, RH=RH)
        elif vapor is not None:
            qv = calc_qv(P=P, vapor=vapor)
        else:
            raise InputError('qv, RH or vapor is required')
    (Td_, Pd, qd), shape, scalar_out, want_tensor = _prepare([T, P, qv])
    ops = [x if D.is_tensor(x) else D.to_device(np.full(shape, x)) for x in (Td_, Pd, qd)]
    out = D.empty(shape)
    scratch = D.empty((16 + out.numel(),))
    _lib.check(_lib.load().mt_calc_tc(D.ptr(ops[0]), D.ptr(ops[1]), D.ptr(ops[2]), D.ptr(out), out.numel(),
                                      D.ptr(scratch), D.stream()), "mt_calc_tc")
    res = out if want_tensor else D.to_host(out)
    if scalar_out and not want_tensor:
        res = float(res[0])
    if return_iterations:
        return res, int(D.to_host(scratch[:16])[15])
    return res


def calc_theta_e(T=None, theta=None, P=None, RH=None, vapor=None, qv=None, Tc=None):  # :405-456
    if theta is None:
        if _none(T, P):
            raise InputError('theta or (T, P) is required')
        theta = calc_theta(T, P)
    if qv is None:
        if not _none(T, RH, P):
            qv = calc_qv(P, T, RH=RH)
        elif not _none(vapor, P):
            qv = calc_qv(P, vapor=vapor)
        else:
            raise InputError('qv, (T, RH, P) or (vapor, P) is required')
    if Tc is None:
        if P is None:
            raise InputError('P or Tc is required')
        Tc = calc_Tc(T, P, qv=qv)
    return _pw("THETA_E", theta, qv, Tc)


def calc_dTdz(T=None, P=None, qvs=None):  # :459-489
    if T is None:
        return -g / Cp
    if qvs is None:
        if P is None:
            raise InputError('P or qvs is required')
        qvs = calc_saturated_qv(T, P)
    return _pw("MOIST_DTDZ", T, qvs)


def calc_RH(T=None, Th=None, vapor=None, qv=None, P=None, Td=None):  # :492-533
    if T is None:
        if _none(Th, P):
            raise InputError('(theta, P) or T is required')
        T = calc_T(Th, P)
    if vapor is None:
        if not _none(qv, P):
            vapor = qv_to_vapor(P, qv)
        elif Td is not None:
            vapor = calc_vapor(Td)
        else:
            raise InputError('vapor, Td or (P, qv) is required')
    return _pw("RH", vapor, T)
