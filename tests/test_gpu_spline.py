"""GPU parity of the spline vertical interpolation (``ver_interp_order = 2`` of
cylindrical_grid.set_data, ``vertical_interp_order = 2 / 3`` of cartesian_grid.set_data) against
vectors from the UNMODIFIED reference (tests/golden/spline.npz) and against scipy run here.

csrc/spline.cu performs FITPACK's operations in FITPACK's order in float64 without FMA, so the
results are expected to be bit-identical to scipy's; the tests hold them to 1e-12 relative
(north_star tolerance: 1e-10) and additionally report through ``assert_same`` on the golden
vectors, where bit equality was observed."""
import numpy as np
import pytest

import oracle
from conftest import assert_close, assert_same, load_golden
from golden.make_golden import SD

pytestmark = pytest.mark.gpu

CAR_S = dict(Nx=10, Ny=9, Nz=10, dx=3000, dy=3000, dz=600, z_start=100, vertical_coord_type="z",
             SOR_parameter=1.9, max_iteration_times=10, iteration_tolerance=0.05)


def _spline_dev(field, vert, levels, order, box=None):
    import torch
    from meteotools_b200 import _device as D, _lib
    lib = _lib.load()
    nz, ny, nx = vert.shape
    y0, y1, x0, x1 = box or (0, ny, 0, nx)
    dtype = _lib.MT_F32 if vert.dtype == np.float32 else _lib.MT_F64
    f = torch.from_numpy(np.ascontiguousarray(field)).cuda()
    v = torch.from_numpy(np.ascontiguousarray(vert)).cuda()
    lv = D.to_device(np.asarray(levels, float))
    out = torch.empty((len(levels), y1 - y0, x1 - x0), dtype=torch.float64, device="cuda")
    _lib.check(lib.mt_spline_levels(D.ptr_array([f]), 1, D.ptr(v), dtype, nz, ny, nx, y0, y1, x0, x1, D.ptr(lv),
                                    len(levels), order, D.ptr_array([out]), (y1 - y0) * (x1 - x0), x1 - x0, 1,
                                    D.stream()), "mt_spline_levels")
    status = lib.mt_spline_status(D.stream())
    return out.cpu().numpy(), status


def test_spline_levels_golden():
    g = load_golden("spline")
    sub = (slice(None), slice(3, 6), slice(4, 9))
    u, z = g["s_u"][sub], g["s_z3d"][sub]
    for order, key in ((2, "o_spline_k2"), (3, "o_spline_k3")):
        got, st = _spline_dev(u, z, g["s_lev"], order)
        assert st == 0
        assert_close(got, g[key], rtol=1e-12, what=key)
        assert_same(got, g[key], key)
    got, st = _spline_dev(u.astype(np.float32), z.astype(np.float32), g["s_lev"], 3)
    assert_same(got, g["o_spline_k3_f32"], "float32 input")
    # a column box of the full array
    got, _ = _spline_dev(g["s_u"], g["s_z3d"], g["s_lev"], 3, box=(3, 6, 4, 9))
    assert_same(got, g["o_spline_k3"], "box")


@pytest.mark.parametrize("order", [1, 2, 3])
def test_spline_levels_random_vs_scipy(order):
    """Irregular columns (60 levels, stretched), extrapolation on both sides, several hundred columns."""
    rng = np.random.default_rng(20 + order)
    nz, ny, nx = 60, 12, 17
    hgt = 1500.0 * rng.random((ny, nx))
    k = np.arange(nz).reshape(-1, 1, 1)
    vert = hgt[None] + (20000.0 - hgt[None]) * (k / (nz - 1)) ** 1.5 + 5.0 * rng.random((nz, ny, nx))
    fld = 300.0 - 6.5e-3 * vert + rng.standard_normal((nz, ny, nx))
    lev = np.arange(60) * 350.0 - 500.0
    got, st = _spline_dev(fld, vert, lev, order)
    assert st == 0
    ref = oracle.spline_levels(fld, vert, lev, order)
    assert_close(got, ref, rtol=1e-12, what=f"k={order}")


def test_spline_decreasing_coordinate_raises():
    """scipy.interpolate.splrep rejects a decreasing abscissa (ValueError: Error on input data);
    the device routine reports it through mt_spline_status and set_data raises the same error."""
    from meteotools_b200.grids import cartesian_grid
    g = load_golden("spline")
    p3d = 1e5 * np.exp(-g["s_z3d"] / 8000.0)
    _, st = _spline_dev(g["s_u"], p3d, [90000.0, 80000.0], 2)
    assert st == 1
    _, st = _spline_dev(g["s_u"], g["s_z3d"], [500.0], 2)
    assert st == 0   # the status was cleared by the previous read
    car = cartesian_grid(dict(CAR_S))
    car.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    with pytest.raises(ValueError, match="Error on input data"):
        car.set_data(float(g["s_dx"]), float(g["s_dx"]), p3d, vertical_interp_order=2, u=g["s_u"])
    with pytest.raises(ValueError):
        oracle.spline_levels(g["s_u"][:, :2, :2], p3d[:, :2, :2], [90000.0], 2)


def test_set_data_spline_golden():
    from meteotools_b200.grids import cartesian_grid, cylindrical_grid
    g = load_golden("spline")
    dx = float(g["s_dx"])
    cyl = cylindrical_grid(dict(SD))
    cyl.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    cyl.set_data(dx, dx, g["s_z3d"], ver_interp_order=2, u=g["s_u"], T=g["s_T"], hgt=g["s_hgt"])
    assert_same(cyl.hgt_idx, g["o_cyl_hgt_idx"], "hgt_idx")
    assert_close(cyl.u, g["o_cyl_u_k2"], rtol=1e-12, what="cyl u k=2")
    assert_close(cyl.T, g["o_cyl_T_k2"], rtol=1e-12, what="cyl T k=2")
    assert_same(cyl.u, g["o_cyl_u_k2"], "cyl u k=2")
    for order in (2, 3):
        car = cartesian_grid(dict(CAR_S))
        car.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
        car.set_data(dx, dx, g["s_z3d"], vertical_interp_order=order, u=g["s_u"], hgt=g["s_hgt"])
        assert_close(car.u, g[f"o_car_u_k{order}"], rtol=1e-12, what=f"car u k={order}")
    with pytest.raises(NameError):
        cyl.set_data(dx, dx, g["s_z3d"], ver_interp_order=3, u=g["s_u"])
    # float32 sources: widened, result float64 as in the reference
    cyl.set_data(dx, dx, g["s_z3d"].astype(np.float32), ver_interp_order=2, u=g["s_u"].astype(np.float32))
    ref = oracle.interp2D_fast_layers(dx, dx, cyl.xTR, cyl.yTR,
                                      oracle.spline_levels(g["s_u"].astype(np.float32),
                                                           g["s_z3d"].astype(np.float32), cyl.z, 2))
    assert_close(cyl.u, ref, rtol=1e-12, what="float32 sources")
