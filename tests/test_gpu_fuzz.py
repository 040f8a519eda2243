"""Randomised API parity of meteotools_b200.interpolation against the oracle's restatement of the
reference wrappers (same signatures): random extents, spacings, target shapes (1-D / 2-D / 3-D),
leading "layer" dimensions, float32 / integer data, NaN targets, targets exactly on nodes and on the
last node, NaN holes in the data.  Bit-exact (assert_same); seeds are fixed, so a failure names its
case."""
import numpy as np
import pytest

import oracle
from conftest import assert_same

pytestmark = pytest.mark.gpu


def _targets(rng, n_axis, d, shape):
    """Targets inside [0, d*(n-1)] with some exact nodes, the last node and NaNs mixed in."""
    hi = d * (n_axis - 1)
    t = rng.random(shape) * hi
    flat = t.reshape(-1)
    k = flat.size
    if k >= 4:
        flat[rng.integers(0, k)] = d * rng.integers(0, n_axis)      # exact node
        flat[rng.integers(0, k)] = hi                               # last node
        flat[rng.integers(0, k)] = 0.0
        if rng.random() < 0.5:
            flat[rng.integers(0, k)] = np.nan
    return t


def _data(rng, shape, kind):
    a = rng.standard_normal(shape) * 10
    if kind == "f32":
        a = a.astype(np.float32)
    elif kind == "int":
        a = np.rint(a).astype(np.int64)
    elif kind == "holes":
        a[rng.random(shape) < 0.05] = np.nan
    return a


KINDS = ["f64", "f32", "int", "holes"]


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_interp1d(seed):
    import meteotools_b200.interpolation as I
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(3, 200))
    d = float(rng.choice([1.0, 0.5, 2000.0, 3.0]))
    kind = KINDS[seed % 4]
    tshape = tuple(rng.integers(1, 7, size=int(rng.integers(1, 3))))
    x = _targets(rng, n, d, tshape)
    a = _data(rng, (n,), kind)
    assert_same(I.interp1D_fast(d, x, a), oracle.interp1D_fast(d, x, a), f"interp1D_fast seed {seed}")
    lead = tuple(rng.integers(1, 5, size=int(rng.integers(1, 3))))
    al = _data(rng, lead + (n,), kind)
    assert_same(I.interp1D_fast_layers(d, x, al), oracle.interp1D_fast_layers(d, x, al), f"1D layers seed {seed}")
    # unequal axis, increasing or decreasing
    xin = np.cumsum(0.1 + rng.random(n))
    if seed % 2:
        xin = xin[::-1].copy()
    lo, hi = xin.min(), xin.max()
    xo = lo + rng.random(tshape) * (hi - lo)
    xo.reshape(-1)[0] = xin[int(rng.integers(0, n))]
    assert_same(I.interp1D_nonequal(xin, xo, a), oracle.interp1D_nonequal(xin, xo, a), f"nonequal seed {seed}")
    assert_same(I.interp1D_nonequal_layers(xin, xo, al), oracle.interp1D_nonequal_layers(xin, xo, al),
                f"nonequal layers seed {seed}")


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_interp2d(seed):
    import meteotools_b200.interpolation as I
    rng = np.random.default_rng(200 + seed)
    ny, nx = int(rng.integers(2, 90)), int(rng.integers(2, 90))
    dx, dy = float(rng.choice([1.0, 2000.0, 0.25])), float(rng.choice([1.0, 3.0, 1500.0]))
    kind = KINDS[seed % 4]
    tshape = tuple(rng.integers(1, 9, size=int(rng.integers(1, 4))))
    x, y = _targets(rng, nx, dx, tshape), _targets(rng, ny, dy, tshape)
    a = _data(rng, (ny, nx), kind)
    assert_same(I.interp2D_fast(dx, dy, x, y, a), oracle.interp2D_fast(dx, dy, x, y, a), f"interp2D_fast seed {seed}")
    lead = tuple(rng.integers(1, 6, size=int(rng.integers(1, 3))))
    al = _data(rng, lead + (ny, nx), kind)
    assert_same(I.interp2D_fast_layers(dx, dy, x, y, al), oracle.interp2D_fast_layers(dx, dy, x, y, al),
                f"2D layers seed {seed}")


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_interp3d(seed):
    import meteotools_b200.interpolation as I
    rng = np.random.default_rng(300 + seed)
    nz, ny, nx = (int(v) for v in rng.integers(2, 30, size=3))
    dx, dy, dz = float(rng.choice([1.0, 2000.0])), float(rng.choice([1.0, 3.0])), float(rng.choice([1.0, 250.0]))
    kind = KINDS[seed % 4]
    tshape = tuple(rng.integers(1, 7, size=int(rng.integers(1, 4))))
    x, y, z = _targets(rng, nx, dx, tshape), _targets(rng, ny, dy, tshape), _targets(rng, nz, dz, tshape)
    a = _data(rng, (nz, ny, nx), kind)
    assert_same(I.interp3D_fast(dx, dy, dz, x, y, z, a), oracle.interp3D_fast(dx, dy, dz, x, y, z, a),
                f"interp3D_fast seed {seed}")
    lead = tuple(rng.integers(1, 4, size=int(rng.integers(1, 3))))
    al = _data(rng, lead + (nz, ny, nx), kind)
    assert_same(I.interp3D_fast_layers(dx, dy, dz, x, y, z, al), oracle.interp3D_fast_layers(dx, dy, dz, x, y, z, al),
                f"3D layers seed {seed}")


FD_NAMES = ("FD2", "FD4", "FD2_front", "FD2_back", "FD2_2", "FD2_2_front", "FD2_2_back")


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_fd_family(seed):
    """Every member of the FD family (calc/differential.py) on random ranks, extents, axes and
    spacings, NaN holes included, against the numpy restatement: bit-exact."""
    import meteotools_b200.calc as C
    from oracle import calc_oracle as co
    rng = np.random.default_rng(400 + seed)
    rank = int(rng.integers(1, 5))
    shape = tuple(int(v) for v in rng.integers(5, 40 if rank < 3 else 14, size=rank))
    axis = int(rng.integers(-rank, rank))
    delta = float(rng.choice([1.0, 0.7, 250.0, 2000.0, 2 * np.pi / 360]))
    v = rng.standard_normal(shape) * 10
    if seed % 3 == 0:
        v[rng.random(shape) < 0.03] = np.nan
    for name in FD_NAMES:
        assert_same(getattr(C, name)(v, delta, axis), getattr(co, name)(v, delta, axis),
                    f"{name} {shape} axis {axis} seed {seed}")


@pytest.mark.parametrize("seed", range(14))
def test_fuzz_cyl_set_data(seed):
    """cylindrical_grid.set_data on random source / target geometries (source extents, model level
    count, even and odd target level counts -> TMA kernel or two-kernel path, radii, centres off
    the cell corners, 1..5 variables, terrain on / off, float32 sources, host arrays or resident
    tensors) against the oracle pipeline interplevel -> interp2D_fast_layers -> terrain mask."""
    import torch
    from meteotools_b200.grids import cylindrical_grid
    from oracle import calc_oracle as co
    rng = np.random.default_rng(500 + seed)
    nz = int(rng.integers(3, 41))
    ny, nx = int(rng.integers(40, 130)), int(rng.integers(40, 130))
    dx = float(rng.choice([2000.0, 1000.0, 3000.0]))
    Nz = int(rng.integers(2, 41))
    Nr, Nt = int(rng.integers(4, 40)), int(rng.integers(4, 60))
    ztop = 12000.0
    rmax_ok = (min(ny, nx) - 1) * dx / 2 - 2 * dx
    dr = float(min(rng.choice([500.0, 1000.0, 1500.0]), rmax_ok / Nr))
    settings = dict(Nr=Nr, Ntheta=Nt, Nz=Nz, dr=dr, dtheta=2 * np.pi / Nt, dz=float(ztop * 0.9 / Nz),
                    z_start=float(rng.choice([0.0, 50.0, 200.0])), r_start=0, theta_start=0,
                    vertical_coord_type="z", dt=300)
    hgt = 800.0 * rng.random((ny, nx)) * (rng.random() < 0.8)
    k = np.arange(nz).reshape(-1, 1, 1)
    z3d = hgt[None] + (ztop - hgt[None]) * (k / (nz - 1)) ** 1.4
    nvar = int(rng.integers(1, 6))
    f32 = seed % 5 == 3
    dt_ = np.float32 if f32 else np.float64
    fields = {f"v{i}": (rng.standard_normal((nz, ny, nx)) * 5 + 0.001 * z3d).astype(dt_) for i in range(nvar)}
    vert = z3d.astype(dt_)
    cyl = cylindrical_grid(settings)
    half = Nr * dr + dx
    cx = float(half + rng.random() * ((nx - 1) * dx - 2 * half))
    cy = float(half + rng.random() * ((ny - 1) * dx - 2 * half))
    cyl.set_horizontal_location(cx, cy)
    with_hgt = seed % 3 != 2
    resident = seed % 4 == 1
    kw = {k_: (torch.from_numpy(v).cuda() if resident else v) for k_, v in fields.items()}
    vv = torch.from_numpy(vert).cuda() if resident else vert
    if with_hgt:
        cyl.set_data(dx, dx, vv, hgt=(torch.from_numpy(hgt).cuda() if resident else hgt), **kw)
    else:
        cyl.set_data(dx, dx, vv, **kw)
    hidx = None
    if with_hgt:
        hidx = co.cyl_terrain_index(oracle.interp2D_fast(dx, dx, cyl.xTR, cyl.yTR, hgt), cyl.z_start, cyl.dz)
    for name, a in fields.items():
        ref = oracle.interp2D_fast_layers(dx, dx, cyl.xTR, cyl.yTR, oracle.interplevel(a, vert, cyl.z))
        if hidx is not None:
            ref = co.cyl_terrain_mask(ref, hidx)
        assert_same(getattr(cyl, name), ref, f"seed {seed} {name} nz={nz} Nz={Nz} src=({ny},{nx}) f32={f32}")


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_car_set_data(seed):
    """cartesian_grid.set_data on random geometries: the TMA kernel's rows go through ONE batched
    transpose that also applies the terrain mask (mt_rows_to_levels) -- against the oracle pipeline
    interplevel -> interp2D_fast_layers -> z <= hgt mask (cartesian_grid.py:127-136).  Target counts
    that are not multiples of the 32-target tile, odd level counts, 1..5 variables, float32 sources,
    resident tensors, and a variable set by an EARLIER call that the later call's terrain must mask too."""
    import torch
    from meteotools_b200.grids import cartesian_grid
    rng = np.random.default_rng(900 + seed)
    nz = int(rng.integers(3, 41))
    ny, nx = int(rng.integers(40, 120)), int(rng.integers(40, 120))
    dx = float(rng.choice([2000.0, 1000.0, 3000.0]))
    Nz = int(rng.integers(2, 41))
    Ny, Nx = int(rng.integers(3, 45)), int(rng.integers(3, 45))
    ztop = 12000.0
    dxt = float(min(rng.choice([700.0, 1500.0, 2500.0]), (nx - 5) * dx / Nx, (ny - 5) * dx / Ny))
    car = cartesian_grid(dict(Nx=Nx, Ny=Ny, Nz=Nz, dx=dxt, dy=dxt, dz=float(ztop * 0.9 / Nz),
                              z_start=float(rng.choice([0.0, 50.0, 200.0])), vertical_coord_type="z",
                              SOR_parameter=1.9, max_iteration_times=10, iteration_tolerance=0.05))
    hx, hy = Nx * dxt / 2 + dx, Ny * dxt / 2 + dx
    cx = float(hx + rng.random() * ((nx - 1) * dx - 2 * hx))
    cy = float(hy + rng.random() * ((ny - 1) * dx - 2 * hy))
    car.set_horizontal_location(cx, cy)
    hgt = 900.0 * rng.random((ny, nx))
    k = np.arange(nz).reshape(-1, 1, 1)
    z3d = hgt[None] + (ztop - hgt[None]) * (k / (nz - 1)) ** 1.4
    nvar = int(rng.integers(1, 6))
    f32 = seed % 4 == 3
    dt_ = np.float32 if f32 else np.float64
    fields = {f"v{i}": (rng.standard_normal((nz, ny, nx)) * 5 + 0.001 * z3d).astype(dt_) for i in range(nvar)}
    vert = z3d.astype(dt_)
    resident = seed % 3 == 1
    dev = (lambda a: torch.from_numpy(a).cuda()) if resident else (lambda a: a)
    with_hgt = seed % 5 != 4
    two_calls = seed % 2 == 0 and nvar >= 2
    kw = {k_: dev(v) for k_, v in fields.items()}
    if two_calls:      # v0 first, without terrain; the second call brings hgt and must mask v0 as well
        first = {"v0": kw.pop("v0")}
        car.set_data(dx, dx, dev(vert), **first)
    if with_hgt:
        car.set_data(dx, dx, dev(vert), hgt=dev(hgt), **kw)
    else:
        car.set_data(dx, dx, dev(vert), **kw)
    hgt_c = oracle.interp2D_fast(dx, dx, car.xTR, car.yTR, hgt) if with_hgt else None
    for name, a in fields.items():
        ref = oracle.interp2D_fast_layers(dx, dx, car.xTR, car.yTR, oracle.interplevel(a, vert, car.z))
        if with_hgt:
            ref = np.where(car.z.reshape(-1, 1, 1) <= hgt_c[None], np.nan, ref)
        got = getattr(car, name)
        assert got.shape == (Nz, Ny, Nx)
        assert_same(got, ref, f"seed {seed} {name} nz={nz} grid=({Nz},{Ny},{Nx}) f32={f32} two_calls={two_calls}")


DYN_EXACT_CYL = ["vr", "vt", "wind_speed", "Tv", "rho", "coriolis_force", "centrifugal_force", "radial_PG_force",
                 "agradient_force", "kinetic_energy", "zeta_r", "zeta_t", "zeta_z", "abs_zeta_z", "div_hori", "div",
                 "density_factor", "shear_deform", "stretch_deform", "strain_region"]
DYN_EXACT_CAR = ["wind_speed", "kinetic_energy", "zeta_r", "zeta_t", "zeta_z", "abs_zeta_z", "div_hori", "div",
                 "density_factor", "Tv", "rho", "shear_deform", "stretch_deform"]


def _dyn_fields(rng, shape, dz, z_start):
    zz = (np.arange(shape[0]) * dz + z_start).reshape(-1, 1, 1)
    f = {k: 10 * rng.standard_normal(shape) for k in ("u", "v", "w")}
    f["p"] = 1e5 * np.exp(-zz / 8000.0) + rng.standard_normal(shape)
    f["T"] = 300 - 6.5e-3 * zz + rng.standard_normal(shape)
    f["qv"] = 1e-3 + 1e-3 * np.abs(rng.standard_normal(shape))
    for k in ("qc", "qr", "qi", "qs", "qg"):
        f[k] = 1e-4 * np.abs(rng.standard_normal(shape))
    f["f"] = 5e-5 + 1e-6 * rng.standard_normal(shape[1:])
    return f


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_dynamics_sets(seed):
    """The fused cyl_d / car_d sets (stage A + marching stage B with its segments, compact shell
    launch and straight-line instantiation) on random small grids, minimum extents included."""
    from conftest import assert_close
    from meteotools_b200.grids import cartesian_grid, cylindrical_grid
    from oracle import calc_oracle as co
    rng = np.random.default_rng(600 + seed)
    nz, n1, n2 = (int(v) for v in rng.integers(3, 34, size=3))
    if seed == 0:
        nz, n1, n2 = 3, 3, 3                      # the smallest grid FD2 accepts
    dz = float(rng.choice([250.0, 300.0, 500.0]))
    f = _dyn_fields(rng, (nz, n1, n2), dz, 100.0)
    if seed % 2 == 0:
        dr = float(rng.choice([1000.0, 1500.0]))
        cyl = cylindrical_grid(dict(Nr=n2, Ntheta=n1, Nz=nz, dr=dr, dtheta=2 * np.pi / n1, dz=dz, z_start=100,
                                    r_start=0, theta_start=0, vertical_coord_type="z", dt=300))
        for k, v in f.items():
            setattr(cyl, k, v.copy())
        cyl.calc_diagnostics()
        with np.errstate(all="ignore"):
            ref = co.cyl_d_set(f, dict(r=cyl.r, theta=cyl.theta, dr=cyl.dr, dtheta=cyl.dtheta, dz=cyl.dz))
            ref2 = co.cyl_d_set(dict(f, Th=np.asarray(cyl.Th)), dict(r=cyl.r, theta=cyl.theta, dr=cyl.dr,
                                                                   dtheta=cyl.dtheta, dz=cyl.dz))
        for k in DYN_EXACT_CYL:
            assert_same(getattr(cyl, k), ref[k], f"cyl seed {seed} {(nz, n1, n2)} {k}")
        assert_close(cyl.Th_rho, ref["Th_rho"], 1e-10, f"cyl seed {seed} Th_rho")
        # with the GPU's own Th as input the whole chain is reproduced bit for bit
        assert_same(cyl.Th_rho, ref2["Th_rho"], f"cyl seed {seed} Th_rho from the GPU's Th")
        assert_same(cyl.PV, ref2["PV"], f"cyl seed {seed} PV from the GPU's Th")
    else:
        dx = float(rng.choice([2000.0, 3000.0]))
        car = cartesian_grid(dict(Nx=n2, Ny=n1, Nz=nz, dx=dx, dy=dx, dz=dz, z_start=100, vertical_coord_type="z",
                                  SOR_parameter=1.9, max_iteration_times=10, iteration_tolerance=0.05))
        for k, v in f.items():
            setattr(car, k, v.copy())
        car.calc_diagnostics()
        with np.errstate(all="ignore"):
            ref = co.car_d_set(f, dict(dx=dx, dy=dx, dz=dz))
            ref2 = co.car_d_set(dict(f, Th=np.asarray(car.Th)), dict(dx=dx, dy=dx, dz=dz))
        for k in DYN_EXACT_CAR:
            assert_same(getattr(car, k), ref[k], f"car seed {seed} {(nz, n1, n2)} {k}")
        assert_same(car.Th_rho, ref2["Th_rho"], f"car seed {seed} Th_rho from the GPU's Th")
        assert_same(car.PV, ref2["PV"], f"car seed {seed} PV from the GPU's Th")


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_cyl_td_set(seed):
    """The fused cyl_td pass on random grid shapes (odd point counts: the scalar tails of the vector
    loops), wide physical ranges, with and without NaN (terrain) -- the two branches of calc_Tc's
    stop test and the speculation in k_td_main -- and with / without qc, qr (the generic kernel)."""
    from conftest import assert_close
    from meteotools_b200.grids import cylindrical_grid
    from oracle import calc_oracle as co
    rng = np.random.default_rng(700 + seed)
    nz, nt, nr = (int(v) for v in rng.integers(3, 22, size=3))
    shape = (nz, nt, nr)
    T = 200.0 + 120.0 * rng.random(shape)
    p = 2000.0 + 1.0e5 * rng.random(shape)
    qv = 10.0 ** (-7 + 5.5 * rng.random(shape))
    qc, qr = 1e-4 * rng.random(shape), 1e-4 * rng.random(shape)
    if seed % 2 == 0:                                   # masked points: the ten sweeps are certain
        hole = rng.random(shape) < 0.1
        for a in (T, p, qv, qc, qr):
            a[hole] = np.nan
    td = cylindrical_grid(dict(Nr=nr, Ntheta=nt, Nz=nz, dr=1000.0, dtheta=2 * np.pi / nt, dz=250.0, z_start=100,
                               r_start=0, theta_start=0, vertical_coord_type="z", dt=300))
    td.T, td.p, td.qv = T.copy(), p.copy(), qv.copy()
    with_ql = seed % 4 < 2
    if with_ql:
        td.qc, td.qr = qc.copy(), qr.copy()
    td.calc_all_thermaldynamic()
    with np.errstate(all="ignore"):
        ref = co.cyl_td_set(T, p, qv, qc if with_ql else None, qr if with_ql else None)
    assert int(td._tc_scratch[15].item()) == ref["_Tc_iters"], "calc_Tc sweep count"
    for k in ("vapor", "Tv", "rho"):
        assert_same(getattr(td, k), ref[k], f"seed {seed} {shape} {k}")
    for k in ("Th", "vapor_s", "qvs", "RH", "Td", "Th_es", "Th_v", "Tc", "Th_e", "moist_dTdz"):
        assert_close(getattr(td, k), ref[k], 1e-10, f"seed {seed} {shape} {k}")


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_range_averages(seed):
    """calc_Zaverage / calc_Raverage / calc_RZaverage on random ranks, axes, ranges and NaN patterns:
    the sums follow NumPy's order for C-order arrays (sequential along strided axes, pairwise blocks
    -- < 8, <= 128 and > 128 terms -- along the contiguous one), so they equal the numpy oracle bit
    for bit."""
    import meteotools_b200.calc as C
    from oracle import calc_oracle as co
    rng = np.random.default_rng(800 + seed)
    rank = int(rng.integers(1, 4))
    shape = tuple(int(v) for v in rng.choice([3, 7, 8, 9, 37, 128, 129, 300, 517], size=rank))
    if rank == 3:
        shape = tuple(min(v, 129) for v in shape)
    a = rng.standard_normal(shape) * 100
    a[rng.random(shape) < 0.05] = np.nan
    for axis in range(rank):
        ax = np.cumsum(0.1 + rng.random(shape[axis]))
        lo, hi = ax[int(rng.integers(0, max(1, shape[axis] // 3)))], ax[-1 - int(rng.integers(0, max(1, shape[axis] // 3)))]
        with np.errstate(all="ignore"):
            assert_same(C.calc_Zaverage(a, ax, axis, lo, hi), co.Zaverage(a, ax, axis, lo, hi), f"Z seed {seed} {shape} axis {axis}")
            assert_same(C.calc_Raverage(a, ax, axis, lo, hi), co.Raverage(a, ax, axis, lo, hi), f"R seed {seed} {shape} axis {axis}")
    if rank == 2:
        zax, rax = np.cumsum(0.1 + rng.random(shape[0])), np.cumsum(0.1 + rng.random(shape[1]))
        for rw in (True, False):
            with np.errstate(all="ignore"):
                got = C.calc_RZaverage(a, zax, rax, rax[1], rax[-2], zax[0], zax[-1], r_weight=rw)
                ref = co.RZaverage(a, zax, rax, rax[1], rax[-2], zax[0], zax[-1], r_weight=rw)
            assert_same(np.float64(got), np.float64(ref), f"RZ seed {seed} {shape} r_weight={rw}")
