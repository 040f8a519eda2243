"""GPU parity of the SURVEY 8(f) rows 1-2 (range averages, axisymmetry, integrated kinetic
energy, smoothers) against outputs of the UNMODIFIED reference (tests/golden/reductions.npz).

Sums: 1e-10 relative (the order of the additions is not NumPy's pairwise order; north-star
tolerance), identical NaN masks.  Smoothers and kinetic_energy10: pure +,*,/ -> BIT-EXACT."""
import numpy as np
import pytest

from conftest import assert_close, assert_same
from golden.make_golden import RED_CAR, RED_CYL
from oracle import calc_oracle as co

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_calc_averages(golden_reductions):
    import torch
    import meteotools_b200.calc as C
    from meteotools_b200.exceptions import DimensionError
    g = golden_reductions
    a, ax = g["r_a"], g["r_ax"]
    # sums along one axis follow NumPy's order for a C-order array (sequential along strided axes,
    # pairwise blocks along the contiguous one): bit-identical to the reference
    assert_same(C.calc_Zaverage(a, ax, 1, ax[4], ax[30]), g["o_Zavg_1"], "Zaverage axis 1")
    assert_same(C.calc_Zaverage(a, ax, 1), g["o_Zavg_1_full"], "Zaverage full range")
    assert_same(C.calc_Raverage(a, ax, 1, ax[2], ax[35]), g["o_Ravg_1"], "Raverage axis 1")
    assert_same(C.calc_Zaverage(np.ascontiguousarray(np.moveaxis(a, 1, 0)), ax, 0, ax[4], ax[30]), g["o_Zavg_0"], "axis 0")
    a2 = np.ascontiguousarray(np.moveaxis(a, 1, 2))
    assert_same(C.calc_Zaverage(a2, ax, 2, ax[4], ax[30]), g["o_Zavg_2"], "contiguous axis (warp kernel)")
    assert_same(C.calc_Raverage(torch.from_numpy(a2).cuda(), ax, 2), g["o_Ravg_2"], "CUDA tensor input")
    assert_same(C.calc_Zaverage(a[2, :, 1], ax, 0, ax[1], ax[20]), g["o_Zavg_1d"], "1-D")
    zr, zax = g["r_zr"], g["r_zax"]
    assert_same(C.calc_RZaverage(zr, zax, ax, ax[3], ax[33], zax[1], zax[6]), g["o_RZavg"], "RZaverage")
    assert_same(C.calc_RZaverage(zr, zax, ax, ax[3], ax[33], zax[1], zax[6], r_weight=False), g["o_RZavg_now"],
                 "RZaverage unweighted")
    with pytest.raises(ValueError):
        C.calc_Zaverage(a, ax[::-1], 1)
    with pytest.raises(DimensionError):
        C.calc_Zaverage(a, ax[:-1], 1)
    # a larger case against the oracle (NumPy's own summation order)
    rng = np.random.default_rng(8)
    big = rng.standard_normal((33, 257, 65))
    big[rng.random(big.shape) < 0.05] = np.nan
    for axis in range(3):
        axv = np.cumsum(0.1 + rng.random(big.shape[axis]))
        with np.errstate(all="ignore"):
            assert_same(C.calc_Zaverage(big, axv, axis, axv[3], axv[-4]), co.Zaverage(big, axv, axis, axv[3], axv[-4]),
                        f"big Zaverage axis {axis}")
            assert_same(C.calc_Raverage(big, axv, axis, axv[3], axv[-4]), co.Raverage(big, axv, axis, axv[3], axv[-4]),
                        f"big Raverage axis {axis}")


def _cyl(g, names):
    from meteotools_b200.grids import cylindrical_grid
    cyl = cylindrical_grid(dict(RED_CYL))
    for k in names:
        setattr(cyl, k, g[f"g_rcyl_{k}"].copy())
    return cyl


def test_cyl_reductions(golden_reductions):
    g = golden_reductions
    cyl = _cyl(g, ("u", "v", "w", "p", "T", "qv", "f"))
    cyl.calc_vr_vt()
    cyl.calc_axisymmetry("vt")
    cyl.calc_axisymmetry("f")
    cyl.prof = g["g_rcyl_T"][:, 2, 3].copy()
    cyl.ring = g["g_rcyl_f"][:, 4].copy()
    cyl.calc_axisymmetry("ring")
    cyl.calc_horizontal_average("T", 0.4, 4.0, 1500., 9000.)
    cyl.calc_horizontal_average("qv")
    cyl.calc_horizontal_average("f", 1, 9, 2000., 11000., unit="index")
    cyl.calc_vertical_average("T", 200., 1600.)
    cyl.calc_vertical_average("prof", 200., 1600.)
    for k in ("sym_vt", "asy_vt", "axisymmetry_vt", "sym_f", "asy_f", "axisymmetry_f", "sym_ring", "asy_ring",
              "axisymmetry_ring", "havg_T", "havg_qv", "havg_f", "zavg_T", "zavg_prof"):
        assert_same(getattr(cyl, k), g[f"o_rcyl_{k}"], k)
    cyl.calc_IKE(rmin=2000., rmax=10000., zmin=200., zmax=1800.)
    assert_close(cyl.IKE, g["o_rcyl_IKE_range"], TOL, "IKE range")
    cyl.calc_IKE()
    assert_close(cyl.IKE, g["o_rcyl_IKE"], TOL, "IKE")
    from meteotools_b200.exceptions import DimensionError, UnitError
    with pytest.raises(UnitError):
        cyl.calc_horizontal_average("T", unit="furlong")
    with pytest.raises(DimensionError):
        cyl.calc_vertical_average("f")
    with pytest.raises(AttributeError):
        cyl.calc_axisymmetry("nope")


def test_cyl_IKE10_reference_quirks(golden_reductions):
    g = golden_reductions
    c10 = _cyl(g, ("u10", "v10", "p", "T", "qv"))
    c10.calc_IKE10(rmax=9000., min_wspd10=3.0)          # rho computed inside: level nearest 10 m
    assert_same(c10.kinetic_energy10, g["o_rcyl_ke10"], "kinetic_energy10 (pure arithmetic)")
    assert_close(c10.IKE10, g["o_rcyl_IKE10"], TOL, "IKE10")
    with pytest.raises(AttributeError):                  # vr10 and vt10 exist now (cylindrical_grid.py:375-381)
        c10.calc_kinetic_energy10()
    c10b = _cyl(g, ("u10", "v10", "p", "T", "qv"))
    c10b.calc_IKE10(rho=1.1)
    assert_close(c10b.IKE10, g["o_rcyl_IKE10_rho"], TOL, "IKE10 with rho")
    c10b.calc_wind_speed10()
    del c10b.kinetic_energy10, c10b.vr10, c10b.vt10
    c10b.calc_IKE10()                                    # rho exists: the NaN default argument is used as it is
    assert c10b.IKE10 == 0.0 == float(g["o_rcyl_IKE10_nanrho"])


def test_car_averages_and_smoothers(golden_reductions):
    from meteotools_b200.grids import cartesian_grid
    g = golden_reductions
    car = cartesian_grid(dict(RED_CAR))
    for k in ("T", "u", "f"):
        setattr(car, k, g[f"g_rcar_{k}"].copy())
    car.calc_horizontal_average("T", 2000., 16000., 4000., 14000.)
    car.calc_horizontal_average("f", ymin=2000.)
    car.calc_vertical_average("T", zmax=1200.)
    for k in ("havg_T", "havg_f", "zavg_T"):
        assert_same(getattr(car, k), g[f"o_rcar_{k}"], k)
    for ax in range(3):
        car.smooth1d("u", axis=ax, center_weight=2, passes=3)
        assert_same(car.smooth1d_u, g[f"o_rcar_smooth1d_u_{ax}"], f"smooth1d axis {ax}")
    car.smooth1d("u", axis=1, center_weight=1.5, passes=1)
    assert_same(car.smooth1d_u, g["o_rcar_smooth1d_u_cw"], "smooth1d cw")
    car.smooth1d("T", axis=0, passes=2)
    assert_same(car.smooth1d_T, g["o_rcar_smooth1d_T"], "smooth1d with NaNs")
    for axes in ((1, 2), (0, 2), (0, 1)):
        car.smooth2d("u", axis=axes, center_weight=2, passes=4)
        assert_same(car.smooth2d_u, g[f"o_rcar_smooth2d_u_{axes[0]}{axes[1]}"], f"smooth2d {axes}")
    assert_same(car.smooth1d_u, g["o_rcar_smooth2d_side"], "smooth1d_u reset by smooth2d (reference side effect)")
    car.smooth2d("f", axis=(0, 1), passes=2)
    assert_same(car.smooth2d_f, g["o_rcar_smooth2d_f"], "smooth2d of a 2-D field")
    car.smooth3d("u", center_weight=2, passes=10)
    assert_same(car.smooth3d_u, g["o_rcar_smooth3d_u"], "smooth3d, 10 passes")
    car.smooth3d("u", center_weight=0.5, passes=1)
    assert_same(car.smooth3d_u, g["o_rcar_smooth3d_u_cw"], "smooth3d cw")
    car.smooth3d("u", passes=0)
    assert_same(car.smooth3d_u, g["g_rcar_u"], "zero passes")


def test_smooth3d_larger_against_oracle():
    from meteotools_b200.grids import cartesian_grid
    rng = np.random.default_rng(9)
    s = dict(RED_CAR, Nx=130, Ny=70, Nz=20)
    car = cartesian_grid(s)
    car.q = rng.standard_normal((20, 70, 130))
    car.smooth3d("q", center_weight=3, passes=5)
    assert_same(car.smooth3d_q, co.smooth3d(np.asarray(car.q), 3, 5), "smooth3d 20x70x130")
    car.smooth2d("q", passes=7)
    assert_same(car.smooth2d_q, co.smooth2d(np.asarray(car.q), (1, 2), 2, 7), "smooth2d 20x70x130")


def test_smoothers_degenerate_and_long_axes_against_oracle():
    """Shapes that exercise the launch geometry of mt_smooth_pass (csrc/reduce.cu): a 1-D array and the last axis of
    a 2-D one arrive as (outer, n, 1) and are re-read with their axes shifted by one; a middle axis longer than
    the 65535 blocks of grid.y is walked by a stride loop; fewer planes than one trip of four; 81 planes (a
    second, partial trip of the plane loop)."""
    from meteotools_b200.grids import cartesian_grid
    rng = np.random.default_rng(21)
    car = cartesian_grid(dict(RED_CAR, Nx=40, Ny=30, Nz=10))
    car.a1 = rng.standard_normal(70001)                       # 1-D, longer than 65535
    car.smooth1d("a1", axis=0, center_weight=2, passes=3)
    assert_same(car.smooth1d_a1, co.smooth1d(np.asarray(car.a1), 0, 2, 3), "smooth1d of a long 1-D array")
    car.a2 = rng.standard_normal((70001, 3))                   # smoothed along a first axis longer than 65535
    car.a2[5, 1] = np.nan
    for ax in (0, 1):
        car.smooth1d("a2", axis=ax, center_weight=1.5, passes=2)
        assert_same(car.smooth1d_a2, co.smooth1d(np.asarray(car.a2), ax, 1.5, 2), f"smooth1d (70001, 3) axis {ax}")
    car.a3 = rng.standard_normal((81, 7, 37))                  # 81 planes, ragged rows
    for ax in (0, 1, 2):
        car.smooth1d("a3", axis=ax, passes=2)
        assert_same(car.smooth1d_a3, co.smooth1d(np.asarray(car.a3), ax, 2, 2), f"smooth1d (81, 7, 37) axis {ax}")
    car.smooth2d("a3", axis=(0, 2), center_weight=2, passes=3)
    assert_same(car.smooth2d_a3, co.smooth2d(np.asarray(car.a3), (0, 2), 2, 3), "smooth2d (81, 7, 37) axes (0, 2)")
    car.smooth3d("a3", center_weight=1, passes=2)
    assert_same(car.smooth3d_a3, co.smooth3d(np.asarray(car.a3), 1, 2), "smooth3d (81, 7, 37)")
    car.a4 = rng.standard_normal((2, 5, 9))                    # fewer planes than one trip
    car.smooth3d("a4", passes=4)
    assert_same(car.smooth3d_a4, co.smooth3d(np.asarray(car.a4), 2, 4), "smooth3d (2, 5, 9)")
