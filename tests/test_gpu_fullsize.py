"""Parity at BASELINE.json's full sizes: against the oracle where it finishes in
seconds (S1, S2 on two variables, S3/S4 one step), and through size-independent properties plus an
oracle-checked sub-slab for S5 (80 x 2000 x 2000: the numpy oracle would need ~40 GB of temporaries)."""
import numpy as np
import pytest

import oracle
from conftest import assert_close, assert_same, parity_record
from oracle import calc_oracle as co

pytestmark = pytest.mark.gpu


def test_S1_cart2cyl_2d():
    import meteotools_b200.interpolation as I
    rng = np.random.default_rng(0)
    Lx = 400 * 2000.0
    yy, xx = np.meshgrid(np.arange(400) * 2000.0, np.arange(400) * 2000.0, indexing="ij")
    f = 10 * np.sin(2 * np.pi * xx / Lx) * np.cos(2 * np.pi * yy / Lx) + 0.1 * rng.standard_normal((400, 400))
    r = np.arange(0, 200 * 2000., 2000.)
    th = np.arange(0, 360 * (2 * np.pi / 360), 2 * np.pi / 360)
    rr, tt = np.meshgrid(r, th)
    x, y = rr * np.cos(tt) + 399000.0, rr * np.sin(tt) + 399000.0
    assert_same(I.interp2D_fast(2000.0, 2000.0, x, y, f), oracle.interp2D_fast(2000.0, 2000.0, x, y, f), "S1")


def _s3_fields(seed=0, shape=(60, 360, 300), dz=250.0, z_start=100.0):
    rng = np.random.default_rng(seed)
    zz = (np.arange(shape[0]) * dz + z_start).reshape(-1, 1, 1)
    f = {k: 10 * rng.standard_normal(shape) for k in ("u", "v", "w")}
    f["p"] = 1e5 * np.exp(-zz / 8000.0) + rng.standard_normal(shape)
    f["T"] = 300 - 6.5e-3 * zz + rng.standard_normal(shape)
    f["qv"] = 1e-3 + 1e-3 * np.abs(rng.standard_normal(shape))
    for k in ("qc", "qr", "qi", "qs", "qg"):
        f[k] = 1e-4 * np.abs(rng.standard_normal(shape))
    f["f"] = np.full(shape[1:], 5e-5)
    return f


S3 = dict(Nr=300, Ntheta=360, Nz=60, dr=1000, dtheta=2 * np.pi / 360, dz=250, z_start=100, r_start=0,
          theta_start=0, vertical_coord_type="z", dt=300)
EXACT = ["vr", "vt", "wind_speed", "Tv", "rho", "coriolis_force", "centrifugal_force", "radial_PG_force",
         "agradient_force", "kinetic_energy", "zeta_r", "zeta_t", "zeta_z", "abs_zeta_z", "div_hori", "div",
         "density_factor", "shear_deform", "stretch_deform", "strain_region"]


def test_S3_cyl_d_full_size():
    from meteotools_b200.grids import cylindrical_grid
    f = _s3_fields()
    cyl = cylindrical_grid(dict(S3))
    for k, v in f.items():
        setattr(cyl, k, v)
    cyl.calc_diagnostics()
    with np.errstate(all="ignore"):
        ref = co.cyl_d_set(f, dict(r=cyl.r, theta=cyl.theta, dr=cyl.dr, dtheta=cyl.dtheta, dz=cyl.dz))
    for k in EXACT + ["Th_rho", "filamentation_time"]:
        parity_record("S3 cyl_d (60,360,300) vs numpy oracle", k, getattr(cyl, k), ref[k])
    parity_record("S3 cyl_d (60,360,300) vs numpy oracle", "PV", cyl.PV, ref["PV"], floor=1e-3)
    for k in EXACT:
        assert_same(getattr(cyl, k), ref[k], f"S3 {k}")
    assert_close(cyl.Th_rho, ref["Th_rho"], 1e-10, "S3 Th_rho")
    assert_close(cyl.filamentation_time, ref["filamentation_time"], 1e-10, "S3 filamentation_time")
    assert_close(cyl.PV, ref["PV"], 1e-10, "S3 PV", nonfinite="class", floor=1e-3)
    # the same chain with the GPU's own Th as input is reproduced BIT FOR BIT by the oracle: the
    # only source of difference above is the <= 1-ulp disagreement of pow between CUDA and NumPy
    with np.errstate(all="ignore"):
        ref2 = co.cyl_d_set(dict(f, Th=np.asarray(cyl.Th)), dict(r=cyl.r, theta=cyl.theta, dr=cyl.dr,
                                                               dtheta=cyl.dtheta, dz=cyl.dz))
    for k in ("Th_rho", "PV"):
        parity_record("S3 cyl_d, oracle fed with the GPU's own Th (isolates pow's ulps)", k, getattr(cyl, k), ref2[k])
    assert_same(cyl.Th_rho, ref2["Th_rho"], "S3 Th_rho from the GPU's Th")
    assert_same(cyl.PV, ref2["PV"], "S3 PV from the GPU's Th")


def test_S4_cyl_td_one_step_full_size():
    from meteotools_b200.grids import cylindrical_grid
    f = _s3_fields(seed=4)
    cyl = cylindrical_grid(dict(S3))
    for k in ("T", "p", "qv", "qc", "qr"):
        setattr(cyl, k, f[k])
    cyl.calc_all_thermaldynamic()
    ref = co.cyl_td_set(f["T"], f["p"], f["qv"], f["qc"], f["qr"])
    assert int(cyl._tc_scratch[15].item()) == ref["_Tc_iters"]
    for k in ("vapor", "Tv", "rho", "Th", "vapor_s", "qvs", "RH", "Td", "Th_es", "Th_v", "Tc", "Th_e", "moist_dTdz"):
        parity_record("S4 cyl_td, one step (60,360,300) vs numpy oracle", k, getattr(cyl, k), ref[k])
    for k in ("vapor", "Tv", "rho"):
        assert_same(getattr(cyl, k), ref[k], f"S4 {k}")
    for k in ("Th", "vapor_s", "qvs", "RH", "Td", "Th_es", "Th_v", "Tc", "Th_e", "moist_dTdz"):
        assert_close(getattr(cyl, k), ref[k], 1e-10, f"S4 {k}")


def test_S2_set_data_full_size():
    """1000 x 1000 x 60 source -> (60, 360, 300): two variables against the CPU pipeline in the
    reference's order (full-domain interplevel, bilinear, terrain mask)."""
    import bench
    from meteotools_b200.grids import cylindrical_grid
    step = bench.make_step(7, pinned=False)
    cyl = bench.make_grid()
    names = ["T", "qv"]
    cyl.set_data(bench.DX, bench.DX, step["z3d"], hgt=step["hgt"], f=step["f"],
                 **{k: step["vars3d"][k] for k in names})
    hgt_c = oracle.interp2D_fast(bench.DX, bench.DX, cyl.xTR, cyl.yTR, step["hgt"])
    assert_same(cyl.hgt, hgt_c, "S2 hgt")
    idx = co.cyl_terrain_index(hgt_c, 100.0, 300.0)
    assert_same(cyl.hgt_idx, idx, "S2 hgt_idx")
    for k in names:
        lev = oracle.interplevel(step["vars3d"][k], step["z3d"], cyl.z)
        ref = co.cyl_terrain_mask(oracle.interp2D_fast_layers(bench.DX, bench.DX, cyl.xTR, cyl.yTR, lev), idx)
        parity_record("S2 set_data 1000x1000x60 -> (60,360,300) vs CPU pipeline", k, getattr(cyl, k), ref)
        assert_same(getattr(cyl, k), ref, f"S2 {k}")
        assert 0.005 < np.isnan(ref).mean() < 0.2


def test_S2_cartesian_set_data_and_full_domain_interplevel():
    """The same 1000 x 1000 x 60 source onto a Cartesian (60, 400, 400) grid (TMA kernel + the batched
    transpose with the terrain mask fused), and interpolation.interplevel over ALL 10^6 columns in the
    level-major layout wrf.interplevel returns (tile kernel, walking brackets) -- one variable each,
    bit for bit against the CPU pipeline."""
    import bench
    from meteotools_b200.grids import cartesian_grid
    from meteotools_b200.interpolation import interplevel
    step = bench.make_step(11, pinned=False)
    car = cartesian_grid(dict(Nx=400, Ny=400, Nz=60, dx=1500, dy=1500, dz=300, z_start=100, vertical_coord_type="z",
                              SOR_parameter=1.9, max_iteration_times=10, iteration_tolerance=0.05))
    c = ((bench.NX + 1) / 2 - 1) * bench.DX
    car.set_horizontal_location(c, c)
    car.set_data(bench.DX, bench.DX, step["z3d"], hgt=step["hgt"], T=step["vars3d"]["T"])
    lev = oracle.interplevel(step["vars3d"]["T"], step["z3d"], car.z)
    got_lev = interplevel(step["vars3d"]["T"], step["z3d"], car.z)
    assert got_lev.shape == lev.shape == (60, bench.NY, bench.NX)
    assert_same(np.asarray(got_lev), lev, "full-domain interplevel")
    hgt_c = oracle.interp2D_fast(bench.DX, bench.DX, car.xTR, car.yTR, step["hgt"])
    ref = oracle.interp2D_fast_layers(bench.DX, bench.DX, car.xTR, car.yTR, lev)
    ref = np.where(car.z.reshape(-1, 1, 1) <= hgt_c[None], np.nan, ref)
    assert_same(car.T, ref, "S2 Cartesian T")
    assert 0.001 < np.isnan(ref).mean() < 0.3


def test_S5_car_d_properties_and_slab():
    """80 x 2000 x 2000 (2.56 GB per field, 74 GB in flight): linear fields give exactly constant
    derivatives, and a y-slab is checked bit for bit against the numpy oracle."""
    import torch
    from meteotools_b200 import _lib
    from meteotools_b200.pipeline import LevelShardDyn
    nz, ny, nx = 80, 2000, 2000
    dz, dy, dx = 250.0, 2000.0, 2000.0
    gen = torch.Generator(device="cuda").manual_seed(5)
    shape = (nz, ny, nx)
    zz = (torch.arange(nz, device="cuda", dtype=torch.float64) * dz + 100.0).reshape(-1, 1, 1)
    local = {k: 10 * torch.randn(shape, generator=gen, device="cuda", dtype=torch.float64) for k in ("u", "v", "w")}
    local["p"] = 1e5 * torch.exp(-zz / 8000.0) + torch.randn(shape, generator=gen, device="cuda", dtype=torch.float64)
    local["T"] = 300 - 6.5e-3 * zz + torch.randn(shape, generator=gen, device="cuda", dtype=torch.float64)
    local["qv"] = 1e-3 + 1e-3 * torch.randn(shape, generator=gen, device="cuda", dtype=torch.float64).abs()
    local["qc"] = 1e-4 * torch.randn(shape, generator=gen, device="cuda", dtype=torch.float64).abs()
    f2d = torch.full((ny, nx), 5e-5, device="cuda", dtype=torch.float64)
    outs = ["zeta_r", "zeta_t", "zeta_z", "div", "PV", "shear_deform", "wind_speed", "Th_rho", "rho"]
    job = LevelShardDyn("car", local, f2d, (dz, dy, dx), (0, nz, 0, 0), nz, outs)
    job.stage_a()
    job.stage_b()
    res = job.result()
    ys = slice(700, 764)
    sub = {k: v[:, 699:765].cpu().numpy() for k, v in local.items()}
    sub["f"] = f2d[699:765].cpu().numpy()
    with np.errstate(all="ignore"):
        ref = co.car_d_set(sub, dict(dx=dx, dy=dy, dz=dz))
    for k in ("zeta_r", "zeta_t", "zeta_z", "div", "shear_deform", "wind_speed", "rho", "Th_rho"):
        parity_record("S5 car_d (80,2000,2000), y-slab of 64 rows vs numpy oracle", k, res[k][:, ys].cpu().numpy(),
                      ref[k][:, 1:-1])
    parity_record("S5 car_d (80,2000,2000), y-slab of 64 rows vs numpy oracle", "PV", res["PV"][:, ys].cpu().numpy(),
                  ref["PV"][:, 1:-1], floor=1e-3)
    for k in ("zeta_r", "zeta_t", "zeta_z", "div", "shear_deform", "wind_speed", "rho"):
        assert_same(res[k][:, ys].cpu().numpy(), ref[k][:, 1:-1], f"S5 slab {k}")
    assert_close(res["PV"][:, ys].cpu().numpy(), ref["PV"][:, 1:-1], 1e-10, "S5 slab PV", floor=1e-3)
    del job, res
    # linearity: u = a*x, v = b*y, w = c*z  ->  div = a+b+c, zeta = 0 (exact for these values)
    x = torch.arange(nx, device="cuda", dtype=torch.float64) * dx
    y = torch.arange(ny, device="cuda", dtype=torch.float64) * dy
    z = torch.arange(nz, device="cuda", dtype=torch.float64) * dz
    local["u"] = (0.5 * x).reshape(1, 1, -1).expand(shape).contiguous()
    local["v"] = (0.25 * y).reshape(1, -1, 1).expand(shape).contiguous()
    local["w"] = (2.0 * z).reshape(-1, 1, 1).expand(shape).contiguous()
    job = LevelShardDyn("car", local, f2d, (dz, dy, dx), (0, nz, 0, 0), nz, ["div", "zeta_r", "zeta_t", "zeta_z"])
    job.stage_a()
    job.stage_b()
    res = job.result()
    assert bool((res["div"] == 2.75).all())
    for k in ("zeta_r", "zeta_t", "zeta_z"):
        assert bool((res[k] == 0.0).all()), k
