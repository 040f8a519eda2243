"""SURVEY 8(f) rows 1-2 on the CPU: the oracle's restatement of the range averages, axisymmetry,
integrated kinetic energy and smoothers reproduces the outputs of the UNMODIFIED reference
(tests/golden/reductions.npz, made by tests/golden/make_golden.py) bit for bit."""
import numpy as np

from conftest import assert_same
from golden.make_golden import RED_CAR, RED_CYL
from oracle import calc_oracle as co


def test_averages_golden(golden_reductions):
    g = golden_reductions
    a, ax = g["r_a"], g["r_ax"]
    with np.errstate(all="ignore"):
        assert_same(co.Zaverage(a, ax, 1, ax[4], ax[30]), g["o_Zavg_1"], "Zaverage axis 1")
        assert_same(co.Zaverage(a, ax, 1), g["o_Zavg_1_full"], "Zaverage full range")
        assert_same(co.Raverage(a, ax, 1, ax[2], ax[35]), g["o_Ravg_1"], "Raverage axis 1")
        assert_same(co.Zaverage(np.ascontiguousarray(np.moveaxis(a, 1, 0)), ax, 0, ax[4], ax[30]), g["o_Zavg_0"], "axis 0")
        a2 = np.ascontiguousarray(np.moveaxis(a, 1, 2))
        assert_same(co.Zaverage(a2, ax, 2, ax[4], ax[30]), g["o_Zavg_2"], "Zaverage contiguous axis")
        assert_same(co.Raverage(a2, ax, 2), g["o_Ravg_2"], "Raverage contiguous axis")
        assert_same(co.Zaverage(a[2, :, 1], ax, 0, ax[1], ax[20]), g["o_Zavg_1d"], "1-D")
        zr, zax = g["r_zr"], g["r_zax"]
        assert_same(co.RZaverage(zr, zax, ax, ax[3], ax[33], zax[1], zax[6]), g["o_RZavg"], "RZaverage")
        assert_same(co.RZaverage(zr, zax, ax, ax[3], ax[33], zax[1], zax[6], r_weight=False), g["o_RZavg_now"],
                    "RZaverage unweighted")


def _cyl_axes():
    c = RED_CYL
    r = np.arange(float(c["r_start"]), c["Nr"] * float(c["dr"]), float(c["dr"]))
    th = np.arange(float(c["theta_start"]), c["Ntheta"] * float(c["dtheta"]), float(c["dtheta"]))
    z = np.arange(0, c["Nz"] * float(c["dz"]), float(c["dz"])) + float(c["z_start"])
    return r, th, z


def test_cyl_reductions_golden(golden_reductions):
    g = golden_reductions
    r, th, z = _cyl_axes()
    dth = float(RED_CYL["dtheta"])
    f = {k[len("g_rcyl_"):]: v for k, v in g.items() if k.startswith("g_rcyl_")}
    with np.errstate(all="ignore"):
        s, c = np.sin(th).reshape(1, -1, 1), np.cos(th).reshape(1, -1, 1)
        vt = f["v"] * c - f["u"] * s
        for name, var in (("vt", vt), ("f", f["f"]), ("ring", f["f"][:, 4])):
            sym, asy, axs = co.axisymmetry(var, dth)
            assert_same(sym, g[f"o_rcyl_sym_{name}"], f"sym_{name}")
            assert_same(asy, g[f"o_rcyl_asy_{name}"], f"asy_{name}")
            assert_same(axs, g[f"o_rcyl_axisymmetry_{name}"], f"axisymmetry_{name}")
        assert_same(co.Raverage(co.Zaverage(f["T"], th, 1, 0.4, 4.0), r, 1, 1500., 9000.), g["o_rcyl_havg_T"], "havg_T")
        assert_same(co.Raverage(co.Zaverage(f["qv"], th, 1), r, 1), g["o_rcyl_havg_qv"], "havg_qv")
        assert_same(co.RZaverage(f["f"], th, r, 2000., 11000., th[1], th[9]), g["o_rcyl_havg_f"], "havg_f")
        assert_same(co.Zaverage(f["T"], z, 0, 200., 1600.), g["o_rcyl_zavg_T"], "zavg_T")
        assert_same(co.Zaverage(f["T"][:, 2, 3], z, 0, 200., 1600.), g["o_rcyl_zavg_prof"], "zavg_prof")
        vr = f["v"] * s + f["u"] * c
        ke = (vr ** 2 + vt ** 2 + f["w"] ** 2) / 2
        rho = co.rho(f["p"], co.Tv_from_qv(f["T"], f["qv"]))
        a = (r, z, float(RED_CYL["dr"]), dth, float(RED_CYL["dz"]))
        assert_same(co.IKE(rho, ke, *a), g["o_rcyl_IKE"], "IKE")
        assert_same(co.IKE(rho, ke, *a, 2000., 10000., 200., 1800.), g["o_rcyl_IKE_range"], "IKE range")
        s2, c2 = np.sin(th).reshape(-1, 1), np.cos(th).reshape(-1, 1)
        vr10, vt10 = f["v10"] * s2 + f["u10"] * c2, f["v10"] * c2 - f["u10"] * s2
        ke10 = (vr10 ** 2 + vt10 ** 2) / 2
        assert_same(ke10, g["o_rcyl_ke10"], "kinetic_energy10")
        ws10 = (f["u10"] ** 2 + f["v10"] ** 2) ** 0.5
        zidx = np.nanargmin(np.abs(z - 10.))
        b = (r, float(RED_CYL["dr"]), dth)
        assert_same(co.IKE10(rho[zidx], ke10, ws10, *b, 9000., 3.0), g["o_rcyl_IKE10"], "IKE10")
        assert_same(co.IKE10(rho[zidx] * 0 + 1.1, ke10, ws10, *b), g["o_rcyl_IKE10_rho"], "IKE10 with rho")
        assert_same(co.IKE10(np.nan, ke10, ws10, *b), g["o_rcyl_IKE10_nanrho"], "IKE10 with the NaN default")


def test_car_reductions_and_smoothers_golden(golden_reductions):
    g = golden_reductions
    c = RED_CAR
    x, y = np.arange(0., c["Nx"] * float(c["dx"]), float(c["dx"])), np.arange(0., c["Ny"] * float(c["dy"]), float(c["dy"]))
    z = np.arange(0, c["Nz"] * float(c["dz"]), float(c["dz"])) + float(c["z_start"])
    T, u, f2 = g["g_rcar_T"], g["g_rcar_u"], g["g_rcar_f"]
    with np.errstate(all="ignore"):
        assert_same(co.Zaverage(co.Zaverage(T, x, 2, 2000., 16000.), y, 1, 4000., 14000.), g["o_rcar_havg_T"], "havg_T")
        assert_same(co.Zaverage(co.Zaverage(f2, x, 1), y, 0, 2000.), g["o_rcar_havg_f"], "havg_f")
        assert_same(co.Zaverage(T, z, 0, zmax=1200.), g["o_rcar_zavg_T"], "zavg_T")
        for ax in range(3):
            assert_same(co.smooth1d(u, ax, 2, 3), g[f"o_rcar_smooth1d_u_{ax}"], f"smooth1d axis {ax}")
        assert_same(co.smooth1d(u, 1, 1.5, 1), g["o_rcar_smooth1d_u_cw"], "smooth1d cw")
        assert_same(co.smooth1d(T, 0, 2, 2), g["o_rcar_smooth1d_T"], "smooth1d NaN")
        for axes in ((1, 2), (0, 2), (0, 1)):
            assert_same(co.smooth2d(u, axes, 2, 4), g[f"o_rcar_smooth2d_u_{axes[0]}{axes[1]}"], f"smooth2d {axes}")
        assert_same(co.smooth2d(f2, (0, 1), 2, 2), g["o_rcar_smooth2d_f"], "smooth2d 2-D")
        assert_same(co.smooth3d(u, 2, 10), g["o_rcar_smooth3d_u"], "smooth3d")
        assert_same(co.smooth3d(u, 0.5, 1), g["o_rcar_smooth3d_u_cw"], "smooth3d cw")
    assert not g["o_rcar_smooth2d_side"].any() and g["o_rcar_smooth2d_side"].shape == u.shape
