"""Pins oracle/calc_oracle.py (numpy restatement of meteotools.calc / meteotools.grids) bit for
bit to outputs of the unmodified reference (tests/golden/{calc,grids}.npz, made by
tests/golden/make_golden.py) and to the reference's own KAT for FD2 (test_calc.py:81-91)."""
import numpy as np

from conftest import assert_same
from oracle import calc_oracle as co
from golden.make_golden import CAR, CAR_OUT, CYL, CYL_OUT, TD_OUT


def test_kat_fd2_linear():  # test_calc.py:81-91
    a = np.add.reduce(np.meshgrid(*[np.arange(n, dtype=float) for n in (4, 5, 6)], indexing="ij"))
    assert np.all(co.FD2(a, 2, 0) == 0.5)
    b = co.FD2_front(a, 2, 0)
    assert np.all(np.isnan(b[-2:])) and abs(np.nansum(b) - 30) < 1e-13


def test_fd_family_golden(golden_calc):
    g = golden_calc
    for name, fn in co.FD_KINDS.items():
        for ax in range(3):
            assert_same(fn(g["c_a"], 0.7, ax), g[f"o_{name}_{ax}"], f"{name} axis {ax}")


def test_thermo_golden(golden_calc):
    g = golden_calc
    T, P, qv, ql, RH = g["t_T"], g["t_P"], g["t_qv"], g["t_ql"], g["t_RH"]
    th, vap = co.theta(T, P), co.qv_to_vapor(P, qv)
    pairs = {
        "theta": th, "T": co.T_from_theta(th, P), "sat_vapor": co.sat_vapor(T),
        "vapor_RH": co.vapor(T, RH), "sat_qv": co.sat_qv(T, P), "qv_to_vapor": vap,
        "theta_es": co.theta_es(T, P), "Td": co.Td_from_vapor(vap), "qv": co.qv_from_vapor(P, vap),
        "theta_v": co.theta_v(th, qv, ql), "Tv": co.Tv_from_qv(T, qv),
        "Tv_vapor": T / (1 - (vap / P) * (1 - 0.622)), "rho": co.rho(P, co.Tv_from_qv(T, qv)),
        "Tc": co.Tc(T, P, qv), "theta_e": co.theta_e(th, qv, co.Tc(T, P, qv)),
        "dTdz": co.moist_dTdz(T, co.sat_qv(T, P)), "RH": co.RH(vap, T),
    }
    for k, v in pairs.items():
        assert_same(v, g[f"o_{k}"], k)
    Tn = T.copy()
    Tn[0, 0, 0] = np.nan
    tc, n = co.Tc(Tn, P, qv, return_iters=True)
    assert n == 10
    assert_same(tc, g["o_Tc_nan"], "Tc with NaN")


def _cyl_grid():
    r = np.arange(CYL["r_start"], CYL["Nr"] * float(CYL["dr"]), float(CYL["dr"]))
    th = np.arange(CYL["theta_start"], CYL["Ntheta"] * float(CYL["dtheta"]), float(CYL["dtheta"]))
    return dict(r=r, theta=th, dr=float(CYL["dr"]), dtheta=float(CYL["dtheta"]), dz=float(CYL["dz"]))


def test_cyl_d_golden(golden_grids):
    g = golden_grids
    f = {k[len("g_cyl_"):]: v for k, v in g.items() if k.startswith("g_cyl_")}
    with np.errstate(all="ignore"):
        o = co.cyl_d_set(f, _cyl_grid())
        o["Th"] = co.theta(f["T"], f["p"])
        o["sym_vt"], o["asy_vt"] = co.azimuthal_stats(o["vt"])
    for k in CYL_OUT:
        if k in o:
            assert_same(o[k], g[f"o_cyl_{k}"], f"cyl {k}")


def test_cyl_td_golden(golden_grids):
    g = golden_grids
    f = {k[len("g_cyl_"):]: v for k, v in g.items() if k.startswith("g_cyl_")}
    o = co.cyl_td_set(f["T"], f["p"], f["qv"], f["qc"], f["qr"])
    for k in TD_OUT:
        assert_same(o[k], g[f"o_td_{k}"], f"td {k}")


def test_car_d_golden(golden_grids):
    g = golden_grids
    f = {k[len("g_car_"):]: v for k, v in g.items() if k.startswith("g_car_")}
    with np.errstate(all="ignore"):
        o = co.car_d_set(f, dict(dx=float(CAR["dx"]), dy=float(CAR["dy"]), dz=float(CAR["dz"])))
        o["Th"] = co.theta(f["T"], f["p"])
    for k in CAR_OUT:
        assert_same(o[k], g[f"o_car_{k}"], f"car {k}")


def test_terrain_golden(golden_setdata):
    import oracle
    g = golden_setdata
    dx = float(g["s_dx"])
    lev = np.arange(10) * 600.0 + 100.0
    tmp = oracle.interplevel(g["s_u"], g["s_z3d"], lev)
    u = oracle.interp2D_fast_layers(dx, dx, g["o_sd_xTR"], g["o_sd_yTR"], tmp)
    assert_same(u, g["o_sd_cyl_u_nomask"], "set_data composition")
    hgt = oracle.interp2D_fast(dx, dx, g["o_sd_xTR"], g["o_sd_yTR"], g["s_hgt"])
    assert_same(hgt, g["o_sd_cyl_hgt"], "hgt")
    idx = co.cyl_terrain_index(hgt, 100.0, 600.0)
    assert_same(idx, g["o_sd_cyl_hgt_idx"], "hgt_idx")
    assert_same(co.cyl_terrain_mask(u, idx), g["o_sd_cyl_u"], "masked u")
