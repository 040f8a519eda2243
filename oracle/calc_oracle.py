"""numpy restatement of the reference's diagnostics -- TEST INFRASTRUCTURE ONLY.

Restates, operation by operation and in the reference's order of evaluation
(so that pure-arithmetic results are bit-identical), the parts of
``meteotools/calc`` and ``meteotools/grids`` that lie on the hot path:

* finite differences          calc/differential.py:6-180
* moist thermodynamics        calc/thermaldynamic.py:4-533
* cylindrical diagnostics     grids/cylindrical_grid.py:264-508
* Cartesian diagnostics       grids/cartesian_grid.py:232-364
* terrain masks               grids/cylindrical_grid.py:139-181, cartesian_grid.py:127-136
* azimuthal statistics        grids/cylindrical_grid.py:187-235,300-314

It is validated against the UNMODIFIED reference package (imported from
/root/reference with stub ``wrf``/``netCDF4`` modules) by
``tests/golden/make_golden.py``; the committed ``tests/golden/*.npz`` files are
outputs of the reference itself, and ``tests/test_oracle_calc.py`` pins this
restatement to them bit for bit.
"""
from __future__ import annotations

import numpy as np

# calc/thermaldynamic.py:4-13
Rd, Cp, g = 287.0, 1004.0, 9.8
kappa = Rd / Cp
Lv, A, B = 2.5e6, 2.53e11, 5420.0


# --------------------------------------------------------------------------- finite differences
def _along(fn, var, delta, axis):
    """calc/check_arrays.py:52-63: float(delta), moveaxis(axis->0), compute, move back."""
    v = np.moveaxis(np.asarray(var), axis, 0)
    return np.moveaxis(fn(v, float(delta)), 0, axis)


def _nan_like(v):
    return np.zeros(v.shape) * np.nan  # differential.py:31 -> always float64


def _fd2(v, d):  # differential.py:30-34
    o = _nan_like(v)
    o[0] = (-3 * v[0] + 4 * v[1] - v[2]) / 2 / d
    o[1:-1] = (v[2:] - v[:-2]) / 2 / d
    o[-1] = (3 * v[-1] - 4 * v[-2] + v[-3]) / 2 / d
    return o


def _fd4(v, d):  # differential.py:66-72
    o = _nan_like(v)
    o[0] = (-3 * v[0] + 4 * v[1] - v[2]) / 2 / d
    o[1] = (v[2] - v[0]) / 2 / d
    o[2:-2] = (-v[4:] + 8 * v[3:-1] - 8 * v[1:-3] + v[:-4]) / 12 / d
    o[-2] = (v[-1] - v[-3]) / 2 / d
    o[-1] = (3 * v[-1] - 4 * v[-2] + v[-3]) / 2 / d
    return o


def _fd2_front(v, d):  # differential.py:92-94
    o = _nan_like(v)
    o[:-2] = (-3 * v[:-2] + 4 * v[1:-1] - v[2:]) / 2 / d
    return o


def _fd2_back(v, d):  # differential.py:113-115
    o = _nan_like(v)
    o[2:] = (3 * v[2:] - 4 * v[1:-1] + v[:-2]) / 2 / d
    return o


def _fd2_2(v, d):  # differential.py:134-138
    o = _nan_like(v)
    o[0] = (2 * v[0] - 5 * v[1] + 4 * v[2] - v[3]) / d / d
    o[1:-1] = (v[2:] - 2 * v[1:-1] + v[:-2]) / d / d
    o[-1] = (2 * v[-1] - 5 * v[-2] + 4 * v[-3] - v[-4]) / d / d
    return o


def _fd2_2_front(v, d):  # differential.py:157-159
    o = _nan_like(v)
    o[:-3] = (2 * v[:-3] - 5 * v[1:-2] + 4 * v[2:-1] - v[3:]) / d / d
    return o


def _fd2_2_back(v, d):  # differential.py:178-180
    o = _nan_like(v)
    o[3:] = (2 * v[3:] - 5 * v[2:-1] + 4 * v[1:-2] - v[:-3]) / d / d
    return o


def FD2(var, delta, axis):
    return _along(_fd2, var, delta, axis)


def FD4(var, delta, axis):
    return _along(_fd4, var, delta, axis)


def FD2_front(var, delta, axis):
    return _along(_fd2_front, var, delta, axis)


def FD2_back(var, delta, axis):
    return _along(_fd2_back, var, delta, axis)


def FD2_2(var, delta, axis):
    return _along(_fd2_2, var, delta, axis)


def FD2_2_front(var, delta, axis):
    return _along(_fd2_2_front, var, delta, axis)


def FD2_2_back(var, delta, axis):
    return _along(_fd2_2_back, var, delta, axis)


FD_KINDS = {"FD2": FD2, "FD4": FD4, "FD2_front": FD2_front, "FD2_back": FD2_back,
            "FD2_2": FD2_2, "FD2_2_front": FD2_2_front, "FD2_2_back": FD2_2_back}


# --------------------------------------------------------------------------- thermodynamics
def theta(T, P):  # thermaldynamic.py:70-84
    return T * (100000 / P) ** kappa


def T_from_theta(th, P):  # :87-101
    return th * (P / 100000) ** kappa


def sat_vapor(T):  # :104-116
    return 611.2 * np.exp(17.67 * (T - 273.15) / (T - 273.15 + 243.5))


def vapor(T, RH=1.0):  # :119-134
    return RH * sat_vapor(T)


def sat_qv(T, P):  # :137-151
    return 0.622 * sat_vapor(T) / (P - sat_vapor(T))


def qv_to_vapor(P, qv):  # :154-168
    return P * qv / (0.622 + qv)


def theta_es(T, P):  # :171-189
    return theta(T, P) * np.exp(Lv * sat_qv(T, P) / Cp / T)


def Td_from_vapor(es):  # :224-225
    lnes = np.log(es / 611.2)
    return 243.5 * lnes / (17.67 - lnes) + 273.15


def qv_from_vapor(P, vap):  # :256
    return 0.622 * vap / (P - vap)


def theta_v(th, qv, ql=0):  # :309
    return th * (1 + qv * 0.61 - ql)


def Tv_from_qv(T, qv):  # :335
    return T * (1 + (qv / 0.622)) / (1 + qv)


def rho(P, Tv):  # :364
    return P / (Rd * Tv)


def Tc(T, P, qv, return_iters=False):  # :392-402 (global, NaN-sensitive stop test)
    Tc2 = T
    out = 0
    n = 0
    for _ in range(10):
        n += 1
        out = B / (np.log(A * 0.622 / P / qv * (T / Tc2) ** (Cp / Rd)))
        if np.max(np.abs(out - Tc2)) < 0.1:
            break
        Tc2 = out
    return (out, n) if return_iters else out


def theta_e(th, qv, tc):  # :456
    return th * np.exp(Lv * qv / Cp / tc)


def moist_dTdz(T, qvs):  # :486
    return -g * ((1 + Lv * qvs / Rd / T) / (Cp + Lv ** 2 * qvs * 0.622 / Rd / T ** 2))


def RH(vap, T):  # :529-530
    return vap / sat_vapor(T)


def cyl_td_set(T, p, qv, qc=None, qr=None):
    """The thermodynamic chain of test_cyl_thermaldynamics.py:96-260 on a grid that holds
    T, p, qv (, qc, qr) -- dispatch per grids/thermaldynamic_calculation.py:7-133."""
    o = {}
    o["Th"] = theta(T, p)
    o["vapor_s"] = sat_vapor(T)
    o["qvs"] = sat_qv(T, p)
    o["vapor"] = qv_to_vapor(p, qv)            # thermaldynamic_calculation.py:36-37
    o["RH"] = RH(o["vapor"], T)
    o["Td"] = Td_from_vapor(o["vapor"])
    o["Th_es"] = theta_es(T, p)
    ql = 0
    if qc is not None and qr is not None:     # :70-72
        ql = qc + qr
    elif qc is not None:
        ql = qc
    elif qr is not None:
        ql = qr
    o["Th_v"] = theta_v(o["Th"], qv, ql)
    o["Tv"] = Tv_from_qv(T, qv)
    o["rho"] = rho(p, o["Tv"])
    o["Tc"], o["_Tc_iters"] = Tc(T, p, qv, return_iters=True)
    o["Th_e"] = theta_e(o["Th"], qv, o["Tc"])
    # calc_moist_dTdz with qvs present takes the qvs-only branch, which returns the DRY rate
    # (thermaldynamic_calculation.py:120-121 -> thermaldynamic.py:485-486); the T,P branch is
    # the moist formula.  Both are exposed.
    o["moist_dTdz"] = moist_dTdz(T, o["qvs"])
    return o


# --------------------------------------------------------------------------- cylindrical set
def density_factor(qv, hydrometeors):  # cylindrical_grid.py:423-447 / cartesian_grid.py:279-303
    a = 1 + qv / 0.622
    b = 1 + qv
    for q in hydrometeors:                     # left-to-right sum, as written
        b = b + q
    return a / b


def cyl_d_set(f, grid):
    """All cylindrical dynamics diagnostics (test_cyl_dynamics.py:148-227) from a dict ``f`` of
    (Nz,Ntheta,Nr) float64 arrays u,v,w,p,T,Th,qv,qc,qr,qi,qs,qg and the (Ntheta,Nr) Coriolis
    ``f``; ``grid`` = dict(r, theta, dr, dtheta, dz).  Returns a dict of outputs."""
    r, th = grid["r"], grid["theta"].reshape(1, -1, 1)
    dr, dth, dz = grid["dr"], grid["dtheta"], grid["dz"]
    w, p, T, fc = f["w"], f["p"], f["T"], f["f"]
    u, v = f.get("u"), f.get("v")
    o = {}
    if "vr" in f and "vt" in f:      # every method uses self.vr / self.vt as they are once they exist (:290-508)
        o["vr"], o["vt"] = f["vr"], f["vt"]
    else:
        o["vr"] = v * np.sin(th) + u * np.cos(th)                  # :266-268
        o["vt"] = v * np.cos(th) - u * np.sin(th)
    vr, vt = o["vr"], o["vt"]
    if u is not None and v is not None:
        o["wind_speed"] = (u ** 2 + v ** 2) ** 0.5                 # :356
    o["Tv"] = Tv_from_qv(T, f["qv"])                               # calc_rho -> calc_Tv
    o["rho"] = rho(p, o["Tv"])
    o["coriolis_force"] = fc * vt                                  # :341-347
    o["centrifugal_force"] = vt ** 2 / r
    o["radial_PG_force"] = -FD2(p, dr, 2) / o["rho"]
    o["agradient_force"] = o["coriolis_force"] + o["centrifugal_force"] + o["radial_PG_force"]
    o["kinetic_energy"] = (vr ** 2 + vt ** 2 + w ** 2) / 2         # :369-371
    o["zeta_r"] = FD2(w, dth, 1) / r - FD2(vt, dz, 0)              # :387-392
    o["zeta_t"] = FD2(vr, dz, 0) - FD2(w, dr, 2)
    o["zeta_z"] = FD2(vt * r, dr, 2) / r - FD2(vr, dth, 1) / r
    o["abs_zeta_z"] = o["zeta_z"] + fc                             # :400
    o["div_hori"] = FD2(r * vr, dr, 2) / r + FD2(vt, dth, 1) / r   # :408-409
    o["div"] = FD2(r * vr, dr, 2) / r + FD2(vt, dth, 1) / r + FD2(w, dz, 0)  # :417-419
    hyd = [f[k] for k in ("qc", "qr", "qi", "qs", "qg", "qh") if k in f]
    o["density_factor"] = density_factor(f["qv"], hyd)
    Th = f["Th"] if "Th" in f else theta(T, p)
    o["Th_rho"] = Th * o["density_factor"]                         # :453
    r_part = o["zeta_r"] * FD2(o["Th_rho"], dr, 2)                 # :463-467
    t_part = o["zeta_t"] * FD2(o["Th_rho"], dth, 1) / r
    z_part = o["abs_zeta_z"] * FD2(o["Th_rho"], dz, 0)
    o["PV"] = (r_part + t_part + z_part) / o["rho"]
    o["shear_deform"] = FD2(vr, dr, 2) - vr / r - FD2(vt, dth, 1) / r      # :490-493
    o["stretch_deform"] = FD2(vt, dr, 2) - vt / r + FD2(vr, dth, 1) / r
    with np.errstate(all="ignore"):
        ft = 2 * (o["shear_deform"] ** 2 + o["stretch_deform"] ** 2 - o["zeta_z"] ** 2) ** (-1 / 2)
    o["strain_region"] = o["shear_deform"] ** 2 + o["stretch_deform"] ** 2 > o["zeta_z"] ** 2
    o["filamentation_time"] = np.where(o["strain_region"], ft, np.inf)     # :501-508
    return o


def azimuthal_stats(var, axis=1):  # cylindrical_grid.py:214-235
    sym = np.nanmean(var, axis=axis)
    asy = var - np.expand_dims(sym, axis)
    return sym, asy


# --------------------------------------------------------------------------- Cartesian set
def car_d_set(f, grid):
    """Cartesian dynamics diagnostics (test_car_dynamics.py:144-202); ``grid``=dict(dx,dy,dz).
    The odd vorticity components are reproduced as written (cartesian_grid.py:250-252)."""
    dx, dy, dz = grid["dx"], grid["dy"], grid["dz"]
    u, v, w, p, T, fc = f["u"], f["v"], f["w"], f["p"], f["T"], f["f"]
    o = {}
    o["wind_speed"] = (u ** 2 + v ** 2) ** 0.5
    o["kinetic_energy"] = (u ** 2 + v ** 2 + w ** 2) / 2
    o["zeta_r"] = FD2(w, dy, 1) - FD2(v, dz, 0)
    o["zeta_t"] = FD2(v, dz, 0) - FD2(w, dx, 2)
    o["zeta_z"] = FD2(v, dx, 2) - FD2(u, dy, 1)
    o["abs_zeta_z"] = o["zeta_z"] + fc
    o["div_hori"] = FD2(u, dx, 2) + FD2(v, dy, 1)
    o["div"] = FD2(u, dx, 2) + FD2(v, dy, 1) + FD2(w, dz, 0)
    hyd = [f[k] for k in ("qc", "qr", "qi", "qs", "qg", "qh") if k in f]
    o["density_factor"] = density_factor(f["qv"], hyd)
    Th = f["Th"] if "Th" in f else theta(T, p)
    o["Th_rho"] = Th * o["density_factor"]
    o["Tv"] = Tv_from_qv(T, f["qv"])
    o["rho"] = rho(p, o["Tv"])
    r_part = o["zeta_r"] * FD2(o["Th_rho"], dx, 2)                 # cartesian_grid.py:319-322
    t_part = o["zeta_t"] * FD2(o["Th_rho"], dy, 1)
    z_part = o["abs_zeta_z"] * FD2(o["Th_rho"], dz, 0)
    o["PV"] = (r_part + t_part + z_part) / o["rho"]
    o["shear_deform"] = FD2(u, dx, 2) - FD2(v, dy, 1)              # :344-347
    o["stretch_deform"] = FD2(v, dx, 2) + FD2(u, dy, 1)
    with np.errstate(all="ignore"):
        ft = 2 * (o["shear_deform"] ** 2 + o["stretch_deform"] ** 2 - o["zeta_z"] ** 2) ** (-1 / 2)
    strain = o["shear_deform"] ** 2 + o["stretch_deform"] ** 2 > o["zeta_z"] ** 2
    o["filamentation_time"] = np.where(strain, ft, np.inf)
    return o


# --------------------------------------------------------------------------- terrain
def cyl_terrain_index(hgt, z_start, dz):  # cylindrical_grid.py:139-176
    idx = ((hgt - z_start) // dz + 1).astype("int32")
    idx = np.where(idx < 0, 0, idx)
    h = idx  # edge / corner fix-up, in the reference's (sequential, in-place) order
    h[0, 1:-1] = np.where(h[0, 1:-1] < h[1, 1:-1], h[1, 1:-1], h[0, 1:-1])
    h[-1, 1:-1] = np.where(h[-1, 1:-1] < h[-2, 1:-1], h[-2, 1:-1], h[-1, 1:-1])
    h[1:-1, 0] = np.where(h[1:-1, 0] < h[1:-1, 1], h[1:-1, 1], h[1:-1, 0])
    h[1:-1, -1] = np.where(h[1:-1, -1] < h[1:-1, -2], h[1:-1, -2], h[1:-1, -1])
    h[0, 0] = np.max([h[0, 0], h[1, 0], h[0, 1]])
    h[0, -1] = np.max([h[0, -1], h[1, -1], h[0, -2]])
    h[-1, 0] = np.max([h[-1, 0], h[-2, 0], h[-1, 1]])
    h[-1, -1] = np.max([h[-1, -1], h[-2, -1], h[-1, -2]])
    return h


def cyl_terrain_mask(var, hgt_idx):  # :171-181
    nz = var.shape[0]
    mask = hgt_idx.reshape((1,) + hgt_idx.shape) > np.arange(nz).reshape(-1, 1, 1)
    return np.where(mask, np.nan, var)


def car_terrain_mask(var, z, hgt):  # cartesian_grid.py:127-136
    return np.where(z.reshape(-1, 1, 1) <= hgt[None], np.nan, var)


# --------------------------------------------------------------------------- SURVEY 8(f) rows 1-2
# Range averages (calc/averages.py:13-114), axisymmetry (cylindrical_grid.py:187-212), integrated
# kinetic energy (:554-600) and the 1-c-1 smoothers (cartesian_grid.py:142-229), restated with the
# reference's NumPy calls so that the summation order (hence every bit) is NumPy's own.
def _range(axis, lo, hi):  # averages.py:5-10 + the selection masks
    lo = axis[0] if np.isnan(lo) else lo
    hi = axis[-1] if np.isnan(hi) else hi
    return np.logical_and(axis >= lo, axis <= hi)


def Zaverage(var, z_axis, axis, zmin=np.nan, zmax=np.nan):  # averages.py:87-114
    sel = _range(np.array(z_axis), zmin, zmax)
    v = np.moveaxis(var, axis, -1)   # check_arrays.py:47 (pre_process_for_1Daverage)
    return np.nansum(v * sel, axis=-1) / np.sum(sel)


def Raverage(var, r_axis, axis, rmin=np.nan, rmax=np.nan):  # averages.py:56-84
    r_axis = np.array(r_axis)
    sel = _range(r_axis, rmin, rmax)
    v = np.moveaxis(var, axis, -1)
    return np.nansum(v * sel * r_axis, axis=-1) / np.nansum(sel * r_axis)


def RZaverage(var, z_axis, r_axis, rmin=np.nan, rmax=np.nan, zmin=np.nan, zmax=np.nan, r_weight=True):  # :13-53
    zs, rs = _range(z_axis, zmin, zmax), _range(r_axis, rmin, rmax)
    tmp = np.nanmean(var[zs], axis=0)
    if r_weight:
        return np.nansum(tmp * rs * r_axis) / np.nansum(rs * r_axis)
    return np.nanmean(tmp[rs])


def axisymmetry(var, dtheta):  # cylindrical_grid.py:187-235 for (z,theta,r), (theta,r) or (theta,)
    ax = 1 if var.ndim == 3 else 0
    sym = np.nanmean(var, axis=ax)
    asy = var - (sym.reshape(var.shape[0], 1, var.shape[2]) if var.ndim == 3 else sym)
    ssq = np.nansum(asy ** 2, axis=ax) if var.ndim > 1 else np.nansum(asy ** 2)
    return sym, asy, sym ** 2 / (sym ** 2 + ssq * dtheta / 2 / np.pi)


def IKE(rho_, ke, r, z, dr, dtheta, dz, rmin=np.nan, rmax=np.nan, zmin=np.nan, zmax=np.nan):  # :554-575
    rmin = r[0] if np.isnan(rmin) else rmin
    rmax = r[-1] if np.isnan(rmax) else rmax
    zmin = z[0] if np.isnan(zmin) else zmin
    zmax = z[-1] if np.isnan(zmax) else zmax

    def filt(axis, lo, hi):  # create_filter_axis :548-552
        f = np.ones(axis.shape)
        f = np.where(axis >= lo, f, 0)
        return np.where(axis <= hi, f, 0)
    nz, nt, nr = rho_.shape
    volume = np.stack([np.stack([r] * nt)] * nz) * dr * dtheta * dz
    volume *= filt(z, zmin, zmax).reshape(-1, 1, 1)
    volume *= filt(r, rmin, rmax).reshape(1, 1, -1)
    return np.nansum(rho_ * ke * volume)


def IKE10(rho2d, ke10, wspd10, r, dr, dtheta, rmax=np.nan, min_wspd10=0.0):  # :577-600 (rho already 2-D / scalar)
    r2d = np.stack([r] * ke10.shape[0])
    if not np.isnan(rmax):
        r2d = np.where(r2d <= rmax, r2d, 0.)
    r2d = np.where(wspd10 > min_wspd10, r2d, 0.)
    return np.nansum(rho2d * ke10 * (r2d * dr * dtheta * 1.))


def smooth1d(var, axis=0, center_weight=2, passes=1):  # cartesian_grid.py:142-167
    v = np.moveaxis(var, axis, 0)
    shape = list(v.shape)
    shape[0] += 2
    tmp = np.zeros(shape)
    tmp[1:-1], tmp[0], tmp[-1] = v, v[0], v[-1]
    for _ in range(passes):
        tmp[1:-1] = (tmp[:-2] + tmp[2:] + tmp[1:-1] * center_weight) / (2 + center_weight)
    return np.moveaxis(tmp[1:-1], 0, axis)


def smooth2d(var, axis=(1, 2), center_weight=2, passes=1):  # :169-199
    v = np.moveaxis(var, axis[1], 0)
    v = np.moveaxis(v, axis[0] + 1, 0)
    shape = list(v.shape)
    shape[0] += 2
    shape[1] += 2
    tmp = np.zeros(shape)
    tmp[1:-1, 1:-1] = v
    tmp[0, 1:-1], tmp[-1, 1:-1], tmp[1:-1, 0], tmp[1:-1, -1] = v[0, :], v[-1, :], v[:, 0], v[:, -1]
    for _ in range(passes):
        tmp[1:-1, 1:-1] = (tmp[:-2, 1:-1] + tmp[2:, 1:-1] + tmp[1:-1, :-2] + tmp[1:-1, 2:]
                           + tmp[1:-1, 1:-1] * center_weight) / (4 + center_weight)
    o = np.moveaxis(tmp[1:-1, 1:-1], 0, axis[0] + 1)
    return np.moveaxis(o, 0, axis[1])


def smooth3d(var, center_weight=2, passes=1):  # :201-229
    shape = [n + 2 for n in var.shape]
    tmp = np.zeros(shape)
    tmp[1:-1, 1:-1, 1:-1] = var
    tmp[0, 1:-1, 1:-1], tmp[-1, 1:-1, 1:-1] = var[0], var[-1]
    tmp[1:-1, 0, 1:-1], tmp[1:-1, -1, 1:-1] = var[:, 0, :], var[:, -1, :]
    tmp[1:-1, 1:-1, 0], tmp[1:-1, 1:-1, -1] = var[:, :, 0], var[:, :, -1]
    for _ in range(passes):
        tmp[1:-1, 1:-1, 1:-1] = (tmp[:-2, 1:-1, 1:-1] + tmp[2:, 1:-1, 1:-1] + tmp[1:-1, :-2, 1:-1]
                                 + tmp[1:-1, 2:, 1:-1] + tmp[1:-1, 1:-1, 2:] + tmp[1:-1, 1:-1, :-2]
                                 + tmp[1:-1, 1:-1, 1:-1] * center_weight) / (6 + center_weight)
    return tmp[1:-1, 1:-1, 1:-1]
