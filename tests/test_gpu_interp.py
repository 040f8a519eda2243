"""GPU parity of the interpolation rows: the 8 legacy C symbols (host pointers,
Fortran layout -- what the unmodified reference binds) and the Python API of
meteotools_b200.interpolation, against the CPU oracle, the reference's KATs and the golden
vectors.  Bit-exact (integer/index work AND the double arithmetic, thanks to -fmad=false)."""
import ctypes as ct

import numpy as np
import pytest

import oracle
from conftest import assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from meteotools_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="module")
def I():
    import meteotools_b200.interpolation as I
    return I


def _p(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_double))


def gen(*shape):
    return np.add.reduce(np.meshgrid(*[np.arange(n, dtype=float) for n in shape], indexing="ij"))


# ------------------------------------------------------------------ legacy symbols vs oracle lib
def test_legacy_symbols_match_oracle(lib):
    olib = oracle.lib()
    rng = np.random.default_rng(5)
    L, nz, ny, nx, n = 7, 5, 11, 13, 257
    dx, dy, dz = 2.0, 3.0, 0.5
    x = rng.random(n) * (nx - 1) * dx
    y = rng.random(n) * (ny - 1) * dy
    z = rng.random(n) * (nz - 1) * dz
    x[:4], y[:4], z[:4] = [(nx - 1) * dx, 4 * dx, 0, 1.7e308], [(ny - 1) * dy, 2 * dy, 0, 1.0], [(nz - 1) * dz, dz, 0, 1.0]
    d2 = np.asfortranarray(rng.standard_normal((ny, nx)))
    d2l = np.asfortranarray(rng.standard_normal((L, ny, nx)))
    d3 = np.asfortranarray(rng.standard_normal((nz, ny, nx)))
    d3l = np.asfortranarray(rng.standard_normal((L, nz, ny, nx)))
    d1 = rng.standard_normal(nx)
    d1l = np.asfortranarray(rng.standard_normal((L, nx)))
    xin = np.cumsum(0.1 + rng.random(nx))
    xo = xin[0] + rng.random(n) * (xin[-1] - xin[0])
    xo[:3] = [xin[0], xin[-1], 1.7e308]

    def both(name, *args, out_shape):
        outs = []
        for L_ in (olib, lib):
            o = np.asfortranarray(np.full(out_shape, -7.0))
            getattr(L_, name)(*args, _p(o))
            outs.append(o)
        assert_same(outs[1], outs[0], name)

    both("interp2D_fast", nx, ny, dx, dy, n, _p(x), _p(y), _p(d2), out_shape=(n,))
    both("interp2D_fast_layers", nx, ny, L, dx, dy, n, _p(x), _p(y), _p(d2l), out_shape=(L, n))
    both("interp3D_fast", nx, ny, nz, dx, dy, dz, n, _p(x), _p(y), _p(z), _p(d3), out_shape=(n,))
    both("interp3D_fast_layers", nx, ny, nz, L, dx, dy, dz, n, _p(x), _p(y), _p(z), _p(d3l), out_shape=(L, n))
    both("interp1D_fast", nx, dx, n, _p(x), _p(d1), out_shape=(n,))
    both("interp1D_fast_layers", nx, L, dx, n, _p(x), _p(d1l), out_shape=(L, n))
    both("interp1D_nonequal", nx, _p(xin), n, _p(xo), _p(d1), out_shape=(n,))
    both("interp1D_nonequal_layers", nx, _p(xin), L, n, _p(xo), _p(d1l), out_shape=(L, n))
    xd = np.ascontiguousarray(xin[::-1])
    both("interp1D_nonequal", nx, _p(xd), n, _p(xo), _p(d1), out_shape=(n,))


def test_legacy_layers_hot_kernel_shapes(lib):
    """layer counts that exercise the warp-per-point kernel (even -> 16-byte loads, odd -> scalar)."""
    olib = oracle.lib()
    rng = np.random.default_rng(6)
    for L in (16, 33, 60, 64, 97):
        ny, nx, n = 40, 37, 1000
        x, y = rng.random(n) * (nx - 1) * 2.0, rng.random(n) * (ny - 1) * 2.0
        x[:2], y[:2] = [0.0, (nx - 1) * 2.0], [0.0, (ny - 1) * 2.0]
        d = np.asfortranarray(rng.standard_normal((L, ny, nx)))
        o0, o1 = (np.asfortranarray(np.zeros((L, n))) for _ in range(2))
        olib.interp2D_fast_layers(nx, ny, L, 2.0, 2.0, n, _p(x), _p(y), _p(d), _p(o0))
        lib.interp2D_fast_layers(nx, ny, L, 2.0, 2.0, n, _p(x), _p(y), _p(d), _p(o1))
        assert_same(o1, o0, f"L={L}")


# ------------------------------------------------------------------ Python API: KATs of the reference
def test_api_kats(I):
    from meteotools_b200.exceptions import DimensionError
    a = gen(5)
    assert np.nanmax(np.abs(I.interp1D_fast(1, [0.5, 2.5, 3.5], a) - [0.5, 2.5, 3.5])) < 1e-13
    a[1] = np.nan
    b = I.interp1D_fast(1, [0.5, 2.5, 3.5], a)
    assert np.isnan(b[0]) and np.nanmax(np.abs(b[1:] - [2.5, 3.5])) < 1e-13
    assert np.isnan(I.interp1D_fast(1, [1.7, 2.5, np.nan], a)[2])
    with pytest.raises(DimensionError):
        I.interp1D_fast(1, [0.5], np.stack([a] * 2))
    with pytest.raises(ValueError):
        I.interp1D_fast(1, [1.7, 7.5], a)
    a = gen(5)
    xin = [0, 2, 4, 5, 6]
    assert np.nanmax(np.abs(I.interp1D_nonequal(xin, [0.5, 2.5, 5.5], a) - [0.25, 1.25, 3.5])) < 1e-13
    with pytest.raises(ValueError):
        I.interp1D_nonequal(xin, [1.7, 7.5], a)
    with pytest.raises(ValueError):
        I.interp1D_nonequal([0, 2, 1, 5, 6], [1.7], a)
    a = gen(5, 5)
    x, y = [1.7, 2.5, 3.5], [1.2, 2.3, 3.4]
    assert np.nanmax(np.abs(I.interp2D_fast(1, 1, x, y, a) - [2.9, 4.8, 6.9])) < 1e-13
    a[1, 1] = np.nan
    assert np.isnan(I.interp2D_fast(1, 1, x, y, a)[0])
    with pytest.raises(DimensionError):
        I.interp2D_fast(1, 1, x, y, np.stack([a] * 2))
    with pytest.raises(DimensionError):
        I.interp2D_fast(1, 1, [1.7, 2.5], y, a)
    with pytest.raises(DimensionError):
        I.interp2D_fast(1, 1, [[1.7, 2.5, 3.5]], y, a)
    with pytest.raises(ValueError):
        I.interp2D_fast(1, 1, [1.7, 7.5, 3.5], y, a)
    a = gen(5, 5, 5)
    x, y, z = [0, 2.5, 3.5], [0, 2.3, 3.4], [0.1, 1.8, 2.9]
    assert np.nanmax(np.abs(I.interp3D_fast(1, 1, 1, x, y, z, a) - [0.1, 6.6, 9.8])) < 1e-13
    a[1, 1, 1] = np.nan
    assert np.isnan(I.interp3D_fast(1, 1, 1, x, y, z, a)[0])
    # layers variants: shapes of test_interpolation.py:219-296
    x, y, z = [0.5, 2.5, 3.5], [1.9, 2.3, 3.4], [0.3, 1.8, 2.9]
    assert I.interp1D_fast_layers(1, x, gen(2, 7, 5)).shape == (2, 7, 3)
    xx, yy = np.meshgrid(x, y)
    b = I.interp2D_fast_layers(1, 1, xx, yy, gen(2, 3, 7, 5))
    assert b.shape == (2, 3, 3, 3) and np.nanmax(np.abs(b - (xx + yy + gen(2, 3)[..., None, None]))) < 1e-13
    xx, yy, zz = np.meshgrid(x, y, z)
    b = I.interp3D_fast_layers(1, 1, 1, xx, yy, zz, gen(2, 4, 5, 7, 5))
    assert b.shape == (2, 4, 3, 3, 3)
    assert np.nanmax(np.abs(b - (xx + yy + zz + gen(2, 4)[..., None, None, None]))) < 1e-13


def test_api_cartesian2cylindrical_exact():  # test_calc.py:58-78 (exact ==)
    from meteotools_b200.calc import cartesian2cylindrical
    r = np.arange(0, 100 + 0.0001, 1)
    theta = np.linspace(0, np.pi * 2, 361)[:-1]
    out = cartesian2cylindrical(1, 1, gen(21, 251, 251), [125, 125], r, theta)
    for (k, t, i), v in {(0, 0, 0): 250., (0, 0, 10): 260., (0, 0, 100): 350., (0, 90, 100): 350.,
                         (0, 180, 100): 150., (0, 270, 100): 150., (20, 0, 0): 270.,
                         (20, 0, 10): 280., (20, 0, 100): 370., (20, 90, 100): 370.,
                         (20, 180, 100): 170., (20, 270, 100): 170.}.items():
        assert out[k, t, i] == v


def test_api_golden(I, golden_interp):
    g = golden_interp
    dx, dy, dz = float(g["i_dx"]), float(g["i_dy"]), float(g["i_dz"])
    xn, yn, zn = g["i_xn"], g["i_yn"], g["i_zn"]
    assert_same(I.interp1D_fast(dx, xn, g["i_a1"]), g["o_interp1D_fast"], "1D")
    assert_same(I.interp1D_fast_layers(dx, xn, g["i_a1l"]), g["o_interp1D_fast_layers"], "1Dl")
    assert_same(I.interp1D_nonequal(g["i_xin"], g["i_xo"], g["i_a1"]), g["o_interp1D_nonequal"], "ne")
    assert_same(I.interp1D_nonequal(g["i_xin_dec"], g["i_xo"], g["i_a1"]), g["o_interp1D_nonequal_dec"], "ne-dec")
    assert_same(I.interp1D_nonequal_layers(g["i_xin"], g["i_xo"], g["i_a1l"]), g["o_interp1D_nonequal_layers"], "nel")
    assert_same(I.interp2D_fast(dx, dy, xn, yn, g["i_a2"]), g["o_interp2D_fast"], "2D")
    assert_same(I.interp2D_fast_layers(dx, dy, xn, yn, g["i_a2l"]), g["o_interp2D_fast_layers"], "2Dl")
    assert_same(I.interp3D_fast(dx, dy, dz, xn, yn, zn, g["i_a3"]), g["o_interp3D_fast"], "3D")
    assert_same(I.interp3D_fast_layers(dx, dy, dz, xn, yn, zn, g["i_a3l"]), g["o_interp3D_fast_layers"], "3Dl")
    from meteotools_b200.calc import cartesian2cylindrical
    r = np.arange(0, 100 + 0.0001, 1)
    theta = np.linspace(0, np.pi * 2, 361)[:-1]
    data = np.add.outer(np.arange(3.0), np.add.outer(np.arange(251.0), np.arange(251.0)))
    assert_same(cartesian2cylindrical(1, 1, data, [125, 125], r, theta), g["o_c2c"], "c2c")


def test_api_float32_and_device_tensor(I):
    """float32 data is widened exactly; CUDA tensors stay on the device."""
    import torch
    rng = np.random.default_rng(8)
    d = rng.standard_normal((3, 9, 12)).astype(np.float32)
    x, y = rng.random(50) * 11 * 2.0, rng.random(50) * 8 * 2.0
    ref = oracle.interp2D_fast_layers(2.0, 2.0, x, y, d)
    assert_same(I.interp2D_fast_layers(2.0, 2.0, x, y, d), ref, "f32")
    t = I.interp2D_fast_layers(2.0, 2.0, x, y, torch.from_numpy(d.astype(np.float64)).cuda())
    assert t.is_cuda
    assert_same(t.cpu().numpy(), ref, "tensor")


def test_plan_indices_and_weights_bit_exact(lib):
    """SURVEY 8d parity gate: indices/weights exported from the plan equal the reference arithmetic."""
    import torch
    from meteotools_b200 import _device as D
    rng = np.random.default_rng(9)
    nx, ny, dx, dy = 400, 400, 2000.0, 2000.0
    r = np.arange(0, 200 * 2000., 2000.)
    th = np.arange(0, 360 * (2 * np.pi / 360), 2 * np.pi / 360)
    rr, tt = np.meshgrid(r, th)
    x = (rr * np.cos(tt) + 399000.0).reshape(-1)
    y = (rr * np.sin(tt) + 399000.0).reshape(-1)
    x[5], y[7] = np.nan, 1.7e308
    xd, yd = D.to_device(x), D.to_device(y)
    ix = torch.empty((x.size, 4), dtype=torch.int32, device="cuda")
    w = torch.empty((x.size, 2), dtype=torch.float64, device="cuda")
    from meteotools_b200 import _lib
    _lib.check(lib.mt_interp2d_plan(D.ptr(xd), D.ptr(yd), x.size, dx, dy, nx, ny, D.ptr(ix), D.ptr(w), D.stream()))
    ix, w = ix.cpu().numpy(), w.cpu().numpy()
    ok = np.ones(x.size, bool)
    ok[[5, 7]] = False
    with np.errstate(invalid="ignore"):
        xt, yt = x / dx, y / dy
        x0, y0 = np.trunc(xt), np.trunc(yt)
    assert ix[5, 0] == -2 and ix[7, 0] == -1
    assert np.array_equal(ix[ok, 0], x0[ok].astype(np.int32))
    assert np.array_equal(ix[ok, 1], np.where(xt == x0, x0, x0 + 1)[ok].astype(np.int32))
    assert np.array_equal(ix[ok, 2], y0[ok].astype(np.int32))
    assert np.array_equal(ix[ok, 3], np.where(yt == y0, y0, y0 + 1)[ok].astype(np.int32))
    assert np.array_equal(w[ok, 0], (xt - x0)[ok]) and np.array_equal(w[ok, 1], (yt - y0)[ok])
    assert np.count_nonzero(xt[ok] == x0[ok]) > 100   # exact-node rule is really exercised
