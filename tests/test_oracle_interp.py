"""Pins the C restatement of fastcompute.f90 (oracle/) against the reference's own known-answer
tests (test_interpolation.py, test_calc.py:58-78), against golden outputs produced by the
reference's Python wrappers (tests/golden/interp.npz) and against an independent pure-Python
reading of the Fortran (the loops of test_interp.py:10-189, with their 0-based off-by-one fixed)."""
import numpy as np
import pytest

import oracle
from conftest import assert_same


def gen(*shape):
    return np.add.reduce(np.meshgrid(*[np.arange(n, dtype=float) for n in shape], indexing="ij"))


def test_kat_interp1D_fast():  # test_interpolation.py:70-92
    a = gen(5)
    x = [0.5, 2.5, 3.5]
    assert np.nanmax(np.abs(oracle.interp1D_fast(1, x, a) - [0.5, 2.5, 3.5])) < 1e-13
    a[1] = np.nan
    b = oracle.interp1D_fast(1, x, a)
    assert np.isnan(b[0]) and np.nanmax(np.abs(b[1:] - [2.5, 3.5])) < 1e-13
    with pytest.raises(oracle.DimensionError):
        oracle.interp1D_fast(1, x, np.stack([a] * 2))
    b = oracle.interp1D_fast(1, [1.7, 2.5, np.nan], a)
    assert np.isnan(b[2])
    with pytest.raises(ValueError):
        oracle.interp1D_fast(1, [1.7, 7.5], a)


def test_kat_interp1D_nonequal():  # test_interpolation.py:95-150
    a = gen(5)
    xin = [0, 2, 4, 5, 6]
    assert np.nanmax(np.abs(oracle.interp1D_nonequal(xin, [0.5, 2.5, 5.5], a) - [0.25, 1.25, 3.5])) < 1e-13
    a[1] = 1
    b = oracle.interp1D_nonequal(xin, [1.7, 2.5, np.nan], a)
    assert np.nanmax(np.abs(b[:2] - [0.85, 1.25])) < 1e-13 and np.isnan(b[2])
    with pytest.raises(ValueError):
        oracle.interp1D_nonequal(xin, [1.7, 7.5], a)
    al = gen(2, 7, 5)
    b = oracle.interp1D_nonequal_layers(xin, [0.5, 2.5, 5.5], al)
    exp = np.array([0.25, 1.25, 3.5])[None, None, :] + gen(2, 7)[..., None]
    assert b.shape == (2, 7, 3) and np.nanmax(np.abs(b - exp)) < 1e-13
    al[1, 1, 1] = np.nan
    b = oracle.interp1D_nonequal_layers(xin, [0.5, 2.5, 5.5], al)
    assert np.isnan(b[1, 1, 0]) and np.isnan(b[1, 1, 1])


def test_kat_interp2D_fast():  # test_interpolation.py:153-181
    a = gen(5, 5)
    x, y = [1.7, 2.5, 3.5], [1.2, 2.3, 3.4]
    assert np.nanmax(np.abs(oracle.interp2D_fast(1, 1, x, y, a) - [2.9, 4.8, 6.9])) < 1e-13
    a[1, 1] = np.nan
    assert np.isnan(oracle.interp2D_fast(1, 1, x, y, a)[0])
    assert np.isnan(oracle.interp2D_fast(1, 1, [1.7, 2.5, np.nan], y, a)[2])
    with pytest.raises(oracle.DimensionError):
        oracle.interp2D_fast(1, 1, x, y, np.stack([a] * 2))
    with pytest.raises(oracle.DimensionError):
        oracle.interp2D_fast(1, 1, [1.7, 2.5], y, a)
    with pytest.raises(ValueError):
        oracle.interp2D_fast(1, 1, [1.7, 7.5, 3.5], y, a)


def test_kat_interp3D_fast():  # test_interpolation.py:184-216
    a = gen(5, 5, 5)
    x, y, z = [0, 2.5, 3.5], [0, 2.3, 3.4], [0.1, 1.8, 2.9]
    assert np.nanmax(np.abs(oracle.interp3D_fast(1, 1, 1, x, y, z, a) - [0.1, 6.6, 9.8])) < 1e-13
    a[1, 1, 1] = np.nan
    assert np.isnan(oracle.interp3D_fast(1, 1, 1, x, y, z, a)[0])   # NaN leaks in through weight 0
    b = oracle.interp3D_fast(1, 1, 1, [2.1, 2.5, np.nan], y, z, a)
    assert np.nanmax(np.abs(b[:2] - [2.2, 6.6])) < 1e-13 and np.isnan(b[2])


def test_kat_layers_shapes():  # test_interpolation.py:219-296
    x, y, z = [0.5, 2.5, 3.5], [1.9, 2.3, 3.4], [0.3, 1.8, 2.9]
    b = oracle.interp1D_fast_layers(1, x, gen(2, 7, 5))
    assert b.shape == (2, 7, 3) and np.nanmax(np.abs(b - (np.array(x) + gen(2, 7)[..., None]))) < 1e-13
    a = gen(2, 3, 7, 5)
    b = oracle.interp2D_fast_layers(1, 1, x, y, a)
    assert b.shape == (2, 3, 3)
    assert np.nanmax(np.abs(b - (np.array(x) + np.array(y) + gen(2, 3)[..., None]))) < 1e-13
    xx, yy = np.meshgrid(x, y)
    b = oracle.interp2D_fast_layers(1, 1, xx, yy, a)
    assert b.shape == (2, 3, 3, 3)
    assert np.nanmax(np.abs(b - (xx + yy + gen(2, 3)[..., None, None]))) < 1e-13
    a = gen(2, 4, 5, 7, 5)
    b = oracle.interp3D_fast_layers(1, 1, 1, x, y, z, a)
    assert b.shape == (2, 4, 3)
    xx, yy, zz = np.meshgrid(x, y, z)
    b = oracle.interp3D_fast_layers(1, 1, 1, xx, yy, zz, a)
    assert b.shape == (2, 4, 3, 3, 3)
    assert np.nanmax(np.abs(b - (xx + yy + zz + gen(2, 4)[..., None, None, None]))) < 1e-13


def test_kat_cartesian2cylindrical_exact():  # test_calc.py:58-78 (exact ==)
    r = np.arange(0, 100 + 0.0001, 1)
    theta = np.linspace(0, np.pi * 2, 361)[:-1]
    data = gen(21, 251, 251)
    rr, tt = np.meshgrid(r, theta)                 # calc/coordinates.py:136-138
    xTR, yTR = rr * np.cos(tt) + 125, rr * np.sin(tt) + 125
    out = oracle.interp2D_fast_layers(1, 1, xTR, yTR, data)
    for (k, t, i), v in {(0, 0, 0): 250., (0, 0, 10): 260., (0, 0, 100): 350., (0, 90, 100): 350.,
                         (0, 180, 100): 150., (0, 270, 100): 150., (20, 0, 0): 270.,
                         (20, 0, 10): 280., (20, 0, 100): 370., (20, 90, 100): 370.,
                         (20, 180, 100): 170., (20, 270, 100): 170.}.items():
        assert out[k, t, i] == v


def test_golden_reference_wrappers(golden_interp):
    g = golden_interp
    dx, dy, dz = float(g["i_dx"]), float(g["i_dy"]), float(g["i_dz"])
    xn, yn, zn = g["i_xn"], g["i_yn"], g["i_zn"]
    assert_same(oracle.interp1D_fast(dx, xn, g["i_a1"]), g["o_interp1D_fast"], "1D")
    assert_same(oracle.interp1D_fast_layers(dx, xn, g["i_a1l"]), g["o_interp1D_fast_layers"], "1Dl")
    assert_same(oracle.interp1D_nonequal(g["i_xin"], g["i_xo"], g["i_a1"]), g["o_interp1D_nonequal"], "ne")
    assert_same(oracle.interp1D_nonequal(g["i_xin_dec"], g["i_xo"], g["i_a1"]), g["o_interp1D_nonequal_dec"], "ne-dec")
    assert_same(oracle.interp1D_nonequal_layers(g["i_xin"], g["i_xo"], g["i_a1l"]), g["o_interp1D_nonequal_layers"], "nel")
    assert_same(oracle.interp2D_fast(dx, dy, xn, yn, g["i_a2"]), g["o_interp2D_fast"], "2D")
    assert_same(oracle.interp2D_fast_layers(dx, dy, xn, yn, g["i_a2l"]), g["o_interp2D_fast_layers"], "2Dl")
    assert_same(oracle.interp3D_fast(dx, dy, dz, xn, yn, zn, g["i_a3"]), g["o_interp3D_fast"], "3D")
    assert_same(oracle.interp3D_fast_layers(dx, dy, dz, xn, yn, zn, g["i_a3l"]), g["o_interp3D_fast_layers"], "3Dl")


# ---- independent pure-Python reading of fastcompute.f90 (small sizes only) ---------------------
def py_bilinear(dx, dy, x, y, d):  # fastcompute.f90:80-102 with 0-based indices
    xt, yt = x / dx, y / dy
    x0, y0 = int(xt), int(yt)
    x1 = x0 if xt == int(xt) else x0 + 1
    y1 = y0 if yt == int(yt) else y0 + 1
    t0 = (d[y0, x1] - d[y0, x0]) * (x / dx - x0) + d[y0, x0]
    t1 = (d[y1, x1] - d[y1, x0]) * (x / dx - x0) + d[y1, x0]
    return (t1 - t0) * (y / dy - y0) + t0


def py_trilinear(dx, dy, dz, x, y, z, d):  # fastcompute.f90:133-172
    nz, ny, nx = d.shape
    x0, y0, z0 = int(x / dx), int(y / dy), int(z / dz)
    x1 = x0 if x == (nx - 1) * dx else x0 + 1
    y1 = y0 if y == (ny - 1) * dy else y0 + 1
    z1 = z0 if z == (nz - 1) * dz else z0 + 1
    zf, yf, xf = z / dz - z0, y / dy - y0, x / dx - x0
    t00 = (d[z1, y0, x0] - d[z0, y0, x0]) * zf + d[z0, y0, x0]
    t10 = (d[z1, y1, x0] - d[z0, y1, x0]) * zf + d[z0, y1, x0]
    t01 = (d[z1, y0, x1] - d[z0, y0, x1]) * zf + d[z0, y0, x1]
    t11 = (d[z1, y1, x1] - d[z0, y1, x1]) * zf + d[z0, y1, x1]
    t0 = (t10 - t00) * yf + t00
    t1 = (t11 - t01) * yf + t01
    return (t1 - t0) * xf + t0


def py_nonequal(xin, xo, d):  # fastcompute.f90:340-354
    x0, x1 = 0, len(xin) - 1
    while x1 - x0 > 1:
        mid = (x0 + 1 + x1 + 1) // 2 - 1
        if (xo - xin[mid]) * (xo - xin[x0]) <= 0:
            x1 = mid
        else:
            x0 = mid
    return (d[x1] - d[x0]) * ((xo - xin[x0]) / (xin[x1] - xin[x0])) + d[x0]


def test_c_vs_pure_python_loops():
    rng = np.random.default_rng(11)
    dx, dy, dz = 1.0, 3.0, 5.0
    d3 = rng.standard_normal((6, 9, 14))
    n = 64
    x = rng.random(n) * 13 * dx
    y = rng.random(n) * 8 * dy
    z = rng.random(n) * 5 * dz
    x[:3], y[:3], z[:3] = [13 * dx, 2 * dx, 0.0], [8 * dy, 3 * dy, 0.0], [5 * dz, 1 * dz, 0.0]
    exp2 = np.array([[py_bilinear(dx, dy, x[i], y[i], d3[k]) for i in range(n)] for k in range(6)])
    assert_same(oracle.interp2D_fast_layers(dx, dy, x, y, d3), exp2, "2D layers vs loops")
    assert_same(oracle.interp2D_fast(dx, dy, x, y, d3[0]), exp2[0], "2D vs loops")
    exp3 = np.array([py_trilinear(dx, dy, dz, x[i], y[i], z[i], d3) for i in range(n)])
    assert_same(oracle.interp3D_fast(dx, dy, dz, x, y, z, d3), exp3, "3D vs loops")
    assert_same(oracle.interp3D_fast_layers(dx, dy, dz, x, y, z, np.stack([d3, 2 * d3]))[1],
                np.array([py_trilinear(dx, dy, dz, x[i], y[i], z[i], 2 * d3) for i in range(n)]), "3Dl")
    xin = np.cumsum(0.2 + rng.random(40))
    xo = xin[0] + rng.random(n) * (xin[-1] - xin[0])
    xo[:2] = [xin[0], xin[-1]]
    d1 = rng.standard_normal(40)
    assert_same(oracle.interp1D_nonequal(xin, xo, d1), np.array([py_nonequal(xin, v, d1) for v in xo]), "ne")
    assert_same(oracle.interp1D_nonequal(xin[::-1].copy(), xo, d1),
                np.array([py_nonequal(xin[::-1], v, d1) for v in xo]), "ne decreasing")


def test_interplevel_semantics(golden_setdata):
    """DINTERP3DZ restatement (parity unpinned): strict brackets, first match from the top,
    missing -> NaN, direction from column (0,0); analytic checks + committed vectors."""
    g = golden_setdata
    z3d, T = g["s_z3d"], g["s_T"]
    lev = np.arange(10) * 600.0 + 100.0
    out = oracle.interplevel(T, z3d, lev)
    assert_same(out, g["o_interplevel_z"], "golden z")
    assert_same(oracle.interplevel(T, g["s_p3d"], g["s_plev"]), g["o_interplevel_p"], "golden p")
    # a field linear in the coordinate is reproduced; below-ground levels are NaN
    lin = oracle.interplevel(2.0 * z3d + 1.0, z3d, lev)
    below = lev.reshape(-1, 1, 1) <= z3d[0][None]
    assert np.isnan(lin[below]).all() and not np.isnan(lin[~below]).any()
    assert np.nanmax(np.abs(lin - (2.0 * lev.reshape(-1, 1, 1) + 1.0))) < 1e-9
    # a level equal to a model level: missing when strict, the level value when inclusive
    v = np.arange(5.0).reshape(5, 1, 1) * np.ones((5, 2, 2))
    f = v * 10
    assert np.isnan(oracle.interplevel(f, v, [2.0])).all()
    assert (oracle.interplevel(f, v, [2.0], inclusive=True) == 20.0).all()
    assert oracle.interplevel(f.astype(np.float32), v, [1.5]).dtype == np.float32
