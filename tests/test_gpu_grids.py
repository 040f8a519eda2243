"""GPU parity of meteotools_b200.grids against golden outputs
of the UNMODIFIED reference grid classes (tests/golden/{grids,setdata}.npz) and the oracle.

Pure-arithmetic outputs (everything made of +,-,*,/,sqrt in the reference's order) must be
BIT-EXACT; outputs that pass through exp/log/pow are held to 1e-10 relative with identical NaN /
inf masks (tolerance rule: conftest.assert_close)."""
import numpy as np
import pytest

import oracle
from conftest import assert_close, assert_same
from golden.make_golden import CAR, CAR_OUT, CYL, SD, TD_OUT
from oracle import calc_oracle as co

pytestmark = pytest.mark.gpu

# cyl_d outputs with no transcendental in their dependency chain (Th_rho/PV depend on theta=pow)
CYL_EXACT = ["vr", "vt", "wind_speed", "Tv", "rho", "coriolis_force", "centrifugal_force",
             "radial_PG_force", "agradient_force", "kinetic_energy", "zeta_r", "zeta_t", "zeta_z",
             "abs_zeta_z", "div_hori", "div", "density_factor", "shear_deform", "stretch_deform",
             "strain_region"]
CYL_CLOSE = ["Th", "Th_rho", "PV", "filamentation_time"]
CAR_EXACT = ["wind_speed", "kinetic_energy", "zeta_r", "zeta_t", "zeta_z", "abs_zeta_z", "div_hori",
             "div", "density_factor", "Tv", "rho", "shear_deform", "stretch_deform"]


def _cyl(golden, with_Th=False):
    from meteotools_b200.grids import cylindrical_grid
    cyl = cylindrical_grid(dict(CYL))
    for k, v in golden.items():
        if k.startswith("g_cyl_"):
            setattr(cyl, k[len("g_cyl_"):], v.copy())
    return cyl


def test_cyl_d_fused(golden_grids):
    g = golden_grids
    cyl = _cyl(g).calc_diagnostics()
    for k in CYL_EXACT:
        assert_same(getattr(cyl, k), g[f"o_cyl_{k}"], f"cyl {k}")
    for k in CYL_CLOSE:
        assert_close(getattr(cyl, k), g[f"o_cyl_{k}"], 1e-10, f"cyl {k}")
    # r[0] = 0 column: NaN / inf exactly where the reference has them
    assert np.isnan(cyl.zeta_z[:, :, 0]).any() or np.isinf(cyl.zeta_z[:, :, 0]).any()


def test_cyl_d_methods_match_reference_sequence(golden_grids):
    """The per-method API of the reference, in the order of test_cyl_dynamics.py:148-227."""
    g = golden_grids
    cyl = _cyl(g)
    for m in ("calc_vr_vt", "calc_wind_speed", "find_rmw_and_speed", "calc_agradient_force",
              "calc_kinetic_energy", "calc_vorticity_3D", "calc_abs_vorticity_3D", "calc_vorticity_z",
              "calc_abs_vorticity_z", "calc_divergence_hori", "calc_divergence",
              "calc_density_potential_temperature", "calc_PV", "calc_deformation", "calc_filamentation"):
        getattr(cyl, m)()
    for k in CYL_EXACT:
        assert_same(getattr(cyl, k), g[f"o_cyl_{k}"], f"cyl {k}")
    for k in CYL_CLOSE:
        assert_close(getattr(cyl, k), g[f"o_cyl_{k}"], 1e-10, f"cyl {k}")
    # the azimuthal mean runs over a strided axis, where NumPy adds sequentially (its pairwise
    # summation only applies along the contiguous axis): the kernel's in-order sum is bit-identical
    assert_same(cyl.sym_vt, g["o_cyl_sym_vt"], "sym_vt")
    assert_same(cyl.asy_vt, g["o_cyl_asy_vt"], "asy_vt")
    assert_same(cyl.RMW, g["o_cyl_RMW"], "RMW")
    assert_same(cyl.vt_max_RMW, g["o_cyl_vt_max_RMW"], "vt_max_RMW")
    cyl.calc_aam()
    r = cyl.r.reshape(1, 1, -1)
    assert_same(cyl.aam, g["o_cyl_vt"] * r + 0.5 * g["g_cyl_f"] * r ** 2, "aam")


def test_cyl_attribute_contract(golden_grids):
    """AttributeError dispatch of the reference (test_cyl_dynamics.py:245-377) and the field store."""
    cyl = _cyl(golden_grids)
    del cyl.w
    for m in ("calc_kinetic_energy", "calc_vorticity_3D", "calc_divergence"):
        with pytest.raises(AttributeError):
            getattr(cyl, m)()
    cyl.calc_vorticity_z()          # does not need w
    del cyl.f
    with pytest.raises(AttributeError):
        cyl.calc_abs_vorticity_z()
    del cyl.u
    del cyl.vr
    with pytest.raises(AttributeError):
        cyl.calc_vr_vt()
    assert 'u' not in dir(cyl) and 'zeta_z' in dir(cyl)
    with pytest.raises(ValueError):
        cyl.zeta_z[0, 0, 0] = 1.0      # read-only view: modify by assignment
    cyl.zeta_z = np.asarray(cyl.zeta_z) * 2
    assert cyl.zeta_z.flags.writeable is False


def test_cyl_td_fused_and_methods(golden_grids):
    g = golden_grids
    from meteotools_b200.grids import cylindrical_grid
    exact = {"vapor", "Tv", "rho"}      # no exp/log/pow
    for mode in ("fused", "methods"):
        td = cylindrical_grid(dict(CYL))
        for k in ("T", "p", "qv", "qc", "qr"):
            setattr(td, k, g[f"g_cyl_{k}"].copy())
        if mode == "fused":
            td.calc_all_thermaldynamic()
        else:
            for m in ("calc_theta", "calc_saturated_vapor", "calc_saturated_qv", "calc_vapor", "calc_RH",
                      "calc_Td", "calc_theta_es", "calc_theta_v", "calc_Tv", "calc_rho", "calc_Tc",
                      "calc_theta_e"):
                getattr(td, m)()
            td.calc_moist_dTdz()    # qvs present -> the dry rate (reference quirk)
            assert td.moist_dTdz == -9.8 / 1004.
            del td.qvs
            td.calc_moist_dTdz()
            td.calc_saturated_qv()
        for k in TD_OUT:
            if k in exact:
                assert_same(getattr(td, k), g[f"o_td_{k}"], f"{mode} td {k}")
            else:
                assert_close(getattr(td, k), g[f"o_td_{k}"], 1e-10, f"{mode} td {k}")
    # dependency contract of test_cyl_thermaldynamics.py:96-260
    td = cylindrical_grid(dict(CYL))
    td.T, td.p = g["g_cyl_T"].copy(), g["g_cyl_p"].copy()
    td.calc_theta()
    del td.T
    with pytest.raises(AttributeError):
        td.calc_theta()
    td.calc_T()
    assert_close(td.T, g["g_cyl_T"], 1e-10, "T from theta")
    del td.p
    with pytest.raises(AttributeError):
        td.calc_saturated_qv()


def test_car_d(golden_grids):
    g = golden_grids
    from meteotools_b200.grids import cartesian_grid
    for mode in ("fused", "methods"):
        car = cartesian_grid(dict(CAR))
        for k, v in g.items():
            if k.startswith("g_car_"):
                setattr(car, k[len("g_car_"):], v.copy())
        if mode == "fused":
            car.calc_diagnostics()
        else:
            for m in ("calc_wind_speed", "calc_kinetic_energy", "calc_vorticity_3D", "calc_abs_vorticity_3D",
                      "calc_divergence_hori", "calc_divergence", "calc_density_potential_temperature",
                      "calc_PV", "calc_deformation", "calc_filamentation"):
                getattr(car, m)()
        for k in CAR_OUT:
            if k in CAR_EXACT:
                assert_same(getattr(car, k), g[f"o_car_{k}"], f"{mode} car {k}")
            else:
                assert_close(getattr(car, k), g[f"o_car_{k}"], 1e-10, f"{mode} car {k}")


def test_set_data_cyl_golden(golden_setdata):
    """Fused set_data (box upload, vertical + bilinear + terrain mask) == the reference pipeline."""
    g = golden_setdata
    from meteotools_b200.grids import cylindrical_grid
    dx = float(g["s_dx"])
    cyl = cylindrical_grid(dict(SD))
    cyl.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    assert_same(cyl.xTR, g["o_sd_xTR"], "xTR")
    cyl.set_data(dx, dx, g["s_z3d"], u=g["s_u"], T=g["s_T"], f=g["s_f"], hgt=g["s_hgt"])
    assert_same(cyl.f, g["o_sd_cyl_f"], "f")
    assert_same(cyl.hgt, g["o_sd_cyl_hgt"], "hgt")
    assert_same(cyl.hgt_idx, g["o_sd_cyl_hgt_idx"], "hgt_idx")
    assert_same(cyl.u, g["o_sd_cyl_u"], "u")
    assert_same(cyl.T, g["o_sd_cyl_T"], "T")
    assert np.isnan(cyl.u).any() and not np.isnan(cyl.u).all()
    cyl2 = cylindrical_grid(dict(SD))
    cyl2.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    cyl2.set_data(dx, dx, g["s_z3d"], u=g["s_u"])
    assert_same(cyl2.u, g["o_sd_cyl_u_nomask"], "u without terrain")
    # float32 sources: wrf-python rounds the level-interpolated field to float32
    cyl3 = cylindrical_grid(dict(SD))
    cyl3.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    u32, z32 = g["s_u"].astype(np.float32), g["s_z3d"].astype(np.float32)
    cyl3.set_data(dx, dx, z32, u=u32)
    ref = oracle.interp2D_fast_layers(dx, dx, g["o_sd_xTR"], g["o_sd_yTR"], oracle.interplevel(u32, z32, cyl3.z))
    assert_same(cyl3.u, ref, "float32 pipeline")
    with pytest.raises(ValueError):
        cyl3.set_horizontal_location(0.0, 0.0)
        cyl3.set_data(dx, dx, g["s_z3d"], u=g["s_u"])


def test_set_data_kernel_variants_agree(golden_setdata, monkeypatch):
    """The three CUDA implementations of set_data -- the default warp-specialised TMA kernel
    (MT_SETDATA=tma, mt_setdata_tma), the two-kernel path (split, mt_setdata) and the first fused
    kernel (fused, mt_setdata_fused) -- give the same bits as the reference pipeline, terrain mask
    and float32 rounding included."""
    g = golden_setdata
    from meteotools_b200 import _lib
    from meteotools_b200.grids import cylindrical_grid
    dx = float(g["s_dx"])
    lib = _lib.load()
    for mode, tile in (("tma", ""), ("split", ""), ("fused", "16x4"), ("fused", "16x2")):
        monkeypatch.setenv("MT_SETDATA", mode)
        monkeypatch.setenv("MT_SETDATA_TILE", tile)
        cyl = cylindrical_grid(dict(SD))
        cyl.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
        before = dict(lib_profile_names(lib, lambda: cyl.set_data(dx, dx, g["s_z3d"], u=g["s_u"], T=g["s_T"],
                                                                   f=g["s_f"], hgt=g["s_hgt"])))
        want = {"tma": "k_setdata_tma", "split": "k_gather2d_lf", "fused": "k_setdata_fused"}[mode]
        assert want in before, (mode, sorted(before))
        assert_same(cyl.u, g["o_sd_cyl_u"], f"{mode} {tile} u")
        assert_same(cyl.T, g["o_sd_cyl_T"], f"{mode} {tile} T")
        u32, z32 = g["s_u"].astype(np.float32), g["s_z3d"].astype(np.float32)
        cyl.set_data(dx, dx, z32, u=u32)
        ref = oracle.interp2D_fast_layers(dx, dx, g["o_sd_xTR"], g["o_sd_yTR"], oracle.interplevel(u32, z32, cyl.z))
        assert_same(cyl.u, ref, f"{mode} {tile} float32")


def lib_profile_names(lib, fn):
    """Kernel names (-> launch counts) recorded by the library's per-launch profiler while fn runs."""
    import ctypes as ct
    import torch
    lib.mt_profile_enable(1)
    try:
        fn()
        torch.cuda.synchronize()
        buf = ct.create_string_buffer(1 << 16)
        n = lib.mt_profile_fetch(buf, len(buf))
        text = buf.value.decode()
    finally:
        lib.mt_profile_enable(0)
    out = {}
    for line in text.splitlines():
        parts = line.split()
        if parts:
            out[parts[0]] = out.get(parts[0], 0) + 1
    return out


@pytest.mark.parametrize("case", ["pressure", "nonmonotone_inclusive", "resident_odd", "many_vars", "levels72",
                                  "levels100", "levels_dense", "levels_descending", "levels_odd", "levels_one",
                                  "resident_oddnx"])
def test_set_data_tma_against_oracle(case, monkeypatch):
    """mt_setdata_tma on shapes the golden file does not hold: decreasing (pressure) coordinate,
    non-monotone columns + the inclusive bracket test, resident device sources with the cylinder
    at an odd offset (tiles hanging over the array edge: TMA out-of-bounds fill), 16 variables,
    72 target levels (the kernel instantiation with 8 level pairs per thread), 100 target levels
    (more than the TMA kernel takes: the two-kernel path must be chosen), closely spaced levels
    (several target levels between two model levels: the bracket walk does not move), descending
    levels on an increasing coordinate (no walk: every bracket takes the binary search)."""
    import torch
    from meteotools_b200.grids import cylindrical_grid
    monkeypatch.setenv("MT_SETDATA", "tma")
    rng = np.random.default_rng({"pressure": 1, "nonmonotone_inclusive": 2, "resident_odd": 3, "many_vars": 4,
                                 "levels72": 5, "levels100": 6, "levels_dense": 7, "levels_descending": 8,
                                 "levels_odd": 9, "levels_one": 10, "resident_oddnx": 11}[case])
    nz, ny, nx, dx = 14, 70, 84, 2000.0
    if case == "resident_oddnx":   # rows of a resident float64 source not 16-byte aligned: the box is staged
        nx = 85
    settings = dict(Nr=30, Ntheta=40, Nz=12, dr=1100, dtheta=2 * np.pi / 40, dz=450, z_start=150,
                    r_start=0, theta_start=0, vertical_coord_type="z", dt=300)
    hgt = 900.0 * rng.random((ny, nx))
    k = np.arange(nz).reshape(-1, 1, 1)
    z3d = hgt[None] + (7000.0 - hgt[None]) * (k / (nz - 1)) ** 1.3
    nvar = 16 if case == "many_vars" else 2
    fields = {f"v{i}": rng.standard_normal((nz, ny, nx)) for i in range(nvar)}
    vert, inclusive = z3d, False
    if case == "pressure":
        settings.update(vertical_coord_type="p", dz=-5000, z_start=95000)   # 95000, 90000, ... Pa
        vert = 1e5 * np.exp(-z3d / 8000.0)
    if case == "levels72":
        settings.update(Nz=72, dz=95, z_start=40)
    if case == "levels100":
        settings.update(Nz=100, dz=70, z_start=40)
    if case == "levels_odd":      # odd level count: 8-byte stores, the last level is its own partner
        settings.update(Nz=61, dz=100, z_start=40)
    if case == "levels_one":
        settings.update(Nz=1, dz=100, z_start=2500)
    if case == "levels_dense":
        settings.update(Nz=40, dz=11, z_start=3000)
    if case == "levels_descending":   # height coordinate, levels from the top down: every bracket is searched
        settings.update(Nz=12, dz=-450, z_start=5500)
    if case == "nonmonotone_inclusive":
        vert = z3d.copy()
        vert[5, ::3, ::4] -= 900.0
        vert[2, 10:20, 10:30] = 150.0 + 450.0 * 3      # exact hits of a target level
        inclusive = True
    cyl = cylindrical_grid(settings)
    cx, cy = (33 * dx + 517.0, 34 * dx + 1311.0) if case != "resident_odd" else (nx * dx - 35e3, 34.5e3)
    cyl.set_horizontal_location(cx, cy)
    kw = dict(fields)
    if case in ("resident_odd", "resident_oddnx"):
        from meteotools_b200 import _lib as _l
        _l.load().mt_profile_enable(1)
        cyl.set_data(dx, dx, torch.from_numpy(vert).cuda(), inclusive=inclusive,
                     **{k_: torch.from_numpy(v).cuda() for k_, v in kw.items()})
        torch.cuda.synchronize()
        import ctypes as _ct
        _buf = _ct.create_string_buffer(1 << 16)
        _l.load().mt_profile_fetch(_buf, len(_buf))
        _l.load().mt_profile_enable(0)
        assert "k_setdata_tma" in _buf.value.decode(), _buf.value.decode()
    elif case in ("pressure", "levels_descending"):
        cyl.set_data(dx, dx, vert, inclusive=inclusive, **kw)
    else:
        cyl.set_data(dx, dx, vert, inclusive=inclusive, hgt=hgt, **kw)
    hidx = None
    if case not in ("resident_odd", "resident_oddnx", "pressure", "levels_descending"):
        hgt_c = oracle.interp2D_fast(dx, dx, cyl.xTR, cyl.yTR, hgt)
        hidx = co.cyl_terrain_index(hgt_c, cyl.z_start, cyl.dz)
    if case in ("levels72", "levels100", "levels_odd", "levels_one"):
        import ctypes as ct
        from meteotools_b200 import _lib
        lib = _lib.load()
        lib.mt_profile_enable(1)
        cyl.set_data(dx, dx, vert, inclusive=inclusive, hgt=hgt, **kw)
        torch.cuda.synchronize()
        buf = ct.create_string_buffer(1 << 16)
        lib.mt_profile_fetch(buf, len(buf))
        lib.mt_profile_enable(0)
        names = buf.value.decode()
        assert ("k_setdata_tma" in names) == (case != "levels100"), names
    for name, a in fields.items():
        levf = oracle.interplevel(a, vert, cyl.z, inclusive=inclusive)
        ref = oracle.interp2D_fast_layers(dx, dx, cyl.xTR, cyl.yTR, levf)
        if hidx is not None:
            ref = co.cyl_terrain_mask(ref, hidx)
        assert_same(getattr(cyl, name), ref, f"{case} {name}")
        assert np.isfinite(ref).any()
def test_set_data_car_golden(golden_setdata):
    g = golden_setdata
    from meteotools_b200.grids import cartesian_grid
    dx = float(g["s_dx"])
    car = cartesian_grid(dict(Nx=10, Ny=9, Nz=10, dx=3000, dy=3000, dz=600, z_start=100,
                              vertical_coord_type="z", SOR_parameter=1.9, max_iteration_times=10,
                              iteration_tolerance=0.05))
    car.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    car.set_data(dx, dx, g["s_z3d"], u=g["s_u"], hgt=g["s_hgt"])
    assert_same(car.hgt, g["o_sd_car_hgt"], "car hgt")
    assert_same(car.u, g["o_sd_car_u"], "car u")


def test_set_data_car_device_tensors(golden_setdata):
    """Cartesian set_data fed with CUDA tensors (2-D field first: the tensor test must not depend on
    torch having been imported through the package)."""
    import torch
    g = golden_setdata
    from meteotools_b200.grids import cartesian_grid
    dx = float(g["s_dx"])
    car = cartesian_grid(dict(Nx=10, Ny=9, Nz=10, dx=3000, dy=3000, dz=600, z_start=100,
                              vertical_coord_type="z", SOR_parameter=1.9, max_iteration_times=10,
                              iteration_tolerance=0.05))
    car.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    car.set_data(dx, dx, torch.from_numpy(g["s_z3d"]).cuda(), hgt=torch.from_numpy(g["s_hgt"]).cuda(),
                 u=torch.from_numpy(g["s_u"]).cuda())
    assert_same(car.hgt, g["o_sd_car_hgt"], "car hgt")
    assert_same(car.u, g["o_sd_car_u"], "car u")


def test_interplevel_device(golden_setdata):
    """mt_interplevel alone: increasing (height) and decreasing (pressure) coordinates, descending
    level list, inclusive switch, non-monotone columns (literal scan) -- against the oracle."""
    import torch
    from meteotools_b200 import _device as D, _lib
    lib = _lib.load()
    g = golden_setdata

    def run(field, vert, levels, order, flags=0, box=None):
        nz, ny, nx = vert.shape
        y0, y1, x0, x1 = box or (0, ny, 0, nx)
        f, v, lv = D.to_device(field), D.to_device(vert), D.to_device(np.asarray(levels, float))
        out = torch.empty((len(levels), y1 - y0, x1 - x0), dtype=torch.float64, device="cuda")
        _lib.check(lib.mt_interplevel(D.ptr_array([f]), 1, D.ptr(v), _lib.MT_F64, nz, ny, nx, y0, y1, x0, x1,
                                      D.ptr(lv), len(levels), order, -1, D.ptr_array([out]),
                                      (y1 - y0) * (x1 - x0), x1 - x0, 1, flags, D.stream()))
        return out.cpu().numpy()

    z3d, T, p3d, plev = g["s_z3d"], g["s_T"], g["s_p3d"], g["s_plev"]
    zlev = np.arange(10) * 600.0 + 100.0
    assert_same(run(T, z3d, zlev, 1), g["o_interplevel_z"], "height, ascending levels")
    assert_same(run(T, p3d, plev, -1), g["o_interplevel_p"], "pressure, descending levels")
    assert_same(run(T, p3d, plev[::-1].copy(), 1), g["o_interplevel_p"][::-1], "pressure, ascending levels")
    assert_same(run(T, z3d, zlev[[3, 0, 7, 2]], 0), g["o_interplevel_z"][[3, 0, 7, 2]], "unsorted levels")
    assert_same(run(T, z3d, zlev, 1, box=(3, 20, 5, 31)), g["o_interplevel_z"][:, 3:20, 5:31], "box")
    # inclusive switch at exact model levels
    v = np.arange(5.0).reshape(5, 1, 1) * np.ones((5, 3, 4))
    assert np.isnan(run(v * 10, v, [2.0], 1)).all()
    assert (run(v * 10, v, [2.0], 1, _lib.MT_IL_INCLUSIVE) == 20.0).all()
    # non-monotone columns take the literal first-match-from-the-top scan
    rng = np.random.default_rng(12)
    vert = np.cumsum(rng.random((9, 6, 7)), axis=0)
    vert[4, ::2, ::3] -= 1.2
    fld = rng.standard_normal((9, 6, 7))
    lev = np.linspace(0.5, 4.0, 8)
    assert_same(run(fld, vert, lev, 1), oracle.interplevel(fld, vert, lev), "non-monotone")
    assert_same(run(fld, vert, lev, 1, _lib.MT_IL_INCLUSIVE), oracle.interplevel(fld, vert, lev, inclusive=True),
                "non-monotone inclusive")


def test_interplevel_python_wrapper(golden_setdata):
    """meteotools_b200.interpolation.interplevel (the wrf.interplevel call of the reference's grids)."""
    import torch
    from meteotools_b200.interpolation import interplevel
    g = golden_setdata
    z3d, T, p3d, plev = g["s_z3d"], g["s_T"], g["s_p3d"], g["s_plev"]
    zlev = np.arange(10) * 600.0 + 100.0
    assert_same(interplevel(T, z3d, zlev), g["o_interplevel_z"], "height")
    assert_same(interplevel(T, p3d, plev), g["o_interplevel_p"], "pressure")
    assert_same(interplevel(T, p3d, 85000.0), g["o_interplevel_p"][1:2], "scalar level")
    got32 = interplevel(T.astype(np.float32), z3d.astype(np.float32), zlev)
    assert got32.dtype == np.float32
    assert_same(got32, oracle.interplevel(T.astype(np.float32), z3d.astype(np.float32), zlev), "float32")
    dev = interplevel(torch.from_numpy(T).cuda(), torch.from_numpy(z3d).cuda(), zlev)
    assert dev.is_cuda and dev.dtype == torch.float64
    assert_same(dev.cpu().numpy(), g["o_interplevel_z"], "device tensors")
    from meteotools_b200.exceptions import DimensionError
    with pytest.raises(DimensionError):
        interplevel(T, z3d[:, :5], zlev)


def test_time_step_stream_matches_grid_api(golden_setdata):
    """pipeline.TimeStepStream (two overlapped slots, host in -> host out) == the grid methods."""
    from meteotools_b200.grids import cylindrical_grid
    from meteotools_b200.pipeline import TimeStepStream
    g = golden_setdata
    dx = float(g["s_dx"])
    z3d, hgt = g["s_z3d"], g["s_hgt"]
    p3d = 1e5 * np.exp(-z3d / 8000.0)
    rng = np.random.default_rng(21)

    def make(k):
        T = g["s_T"] + 280.0 + k
        return dict(z3d=z3d, hgt=hgt, f=g["s_f"],
                    vars3d=dict(T=T, p=p3d + k, qv=1e-3 + 1e-4 * np.abs(g["s_u"]) + 1e-5 * k,
                                qc=1e-4 * np.abs(g["s_u"]), qr=1e-4 * np.abs(g["s_T"])))
    steps = [make(k) for k in range(5)]
    cyl = cylindrical_grid(dict(SD))
    cyl.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    outs = ["Th", "Tc", "Th_e", "rho", "vapor"]
    ts = TimeStepStream(cyl, z3d.shape, dx, dx, ["T", "p", "qv", "qc", "qr"], names2d=("f",), hgt=True, outputs=outs)
    seen = []
    for k, res in ts.process(steps):
        ref = cylindrical_grid(dict(SD))
        ref.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
        ref.set_data(dx, dx, z3d, hgt=hgt, f=g["s_f"], **steps[k]["vars3d"])
        ref.calc_all_thermaldynamic()
        for name in outs:
            assert_same(res[name], getattr(ref, name), f"step {k} {name}")
        seen.append(k)
    assert seen == list(range(5))


def test_time_step_stream_float32_results(golden_setdata):
    """out_dtype=float32: the same results, rounded to float32 on the device (what astype does)."""
    from meteotools_b200.grids import cylindrical_grid
    from meteotools_b200.pipeline import TimeStepStream
    g = golden_setdata
    dx = float(g["s_dx"])
    src = dict(z3d=g["s_z3d"], f=g["s_f"], hgt=g["s_hgt"],
               vars3d=dict(T=g["s_T"] + 280.0, p=1e5 * np.exp(-g["s_z3d"] / 8000.0),
                           qv=1e-3 + 1e-4 * np.abs(g["s_u"])))
    outs = ["Th", "Tc", "rho"]
    res = {}
    for od in (np.float64, np.float32):
        cyl = cylindrical_grid(dict(SD))
        cyl.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
        ts = TimeStepStream(cyl, g["s_z3d"].shape, dx, dx, ["T", "p", "qv"], names2d=("f",), hgt=True, outputs=outs,
                            out_dtype=od)
        for _, r in ts.process([src, src]):
            res[od] = {k: v.copy() for k, v in r.items()}
    for k in outs:
        assert res[np.float32][k].dtype == np.float32
        with np.errstate(all="ignore"):
            assert_same(res[np.float32][k], res[np.float64][k].astype(np.float32), f"float32 results {k}")


def test_time_step_stream_float32_sources(golden_setdata):
    """float32 host sources are staged as float32 (half the PCIe bytes), widened in the
    kernel, and give exactly what the grid API gives for the same float32 arrays (wrf-python's
    float32 rounding of the level-interpolated field included)."""
    from meteotools_b200.grids import cylindrical_grid
    from meteotools_b200.pipeline import TimeStepStream
    g = golden_setdata
    dx = float(g["s_dx"])
    z32, hgt = g["s_z3d"].astype(np.float32), g["s_hgt"]
    v32 = dict(T=(g["s_T"] + 280.0).astype(np.float32), p=(1e5 * np.exp(-g["s_z3d"] / 8000.0)).astype(np.float32),
               qv=(1e-3 + 1e-4 * np.abs(g["s_u"])).astype(np.float32))
    cyl = cylindrical_grid(dict(SD))
    cyl.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    outs = ["Th", "Tc", "rho"]
    ts = TimeStepStream(cyl, z32.shape, dx, dx, ["T", "p", "qv"], names2d=("f",), hgt=True, outputs=outs + ["T"],
                        dtype=np.float32)
    assert ts.slots[0]["pipe"].h2d_bytes() < 0.6 * (4 * z32.size * 8)
    steps = [dict(z3d=z32, hgt=hgt, f=g["s_f"], vars3d=v32)] * 3
    ref = cylindrical_grid(dict(SD))
    ref.set_horizontal_location(float(g["s_cx"]), float(g["s_cy"]))
    ref.set_data(dx, dx, z32, hgt=hgt, f=g["s_f"], **v32)
    ref.calc_all_thermaldynamic()
    lev = oracle.interplevel(v32["T"], z32, ref.z)
    assert lev.dtype == np.float32
    want_T = co.cyl_terrain_mask(oracle.interp2D_fast_layers(dx, dx, g["o_sd_xTR"], g["o_sd_yTR"], lev), ref.hgt_idx)
    n = 0
    for k, res in ts.process(steps):
        assert_same(res["T"], want_T, f"step {k} T vs the CPU pipeline on float32 input")
        for name in outs:
            assert_same(res[name], getattr(ref, name), f"step {k} {name}")
        n += 1
    assert n == 3


@pytest.mark.parametrize("n,nlev,nvar,mask", [(1, 1, 1, False), (31, 7, 2, True), (32, 60, 3, True), (1000, 96, 16, False),
                                              (4097, 61, 5, True)])
def test_rows_to_levels_abi(n, nlev, nvar, mask):
    """mt_rows_to_levels through the C ABI: out[v][k, i] = rows[v][i, k], NaN where z[k] <= hgt[i]
    (cartesian_grid.py:127-136); target counts off the 32-target tile, 1 and 96 levels, 16 variables,
    NaN already present in the rows."""
    import torch
    from meteotools_b200 import _device as D, _lib
    lib = _lib.load()
    rng = np.random.default_rng(n * 131 + nlev)
    rows = [rng.standard_normal((n, nlev)) for _ in range(nvar)]
    rows[0][rng.random((n, nlev)) < 0.05] = np.nan
    z = np.sort(rng.random(nlev)) * 3000.0
    hgt = rng.random(n) * 2000.0
    if n > 2:
        hgt[1] = z[nlev // 2]          # equality masks (<=)
    d_rows = [torch.from_numpy(r).cuda() for r in rows]
    d_out = [torch.full((nlev, n), -7.0, dtype=torch.float64, device="cuda") for _ in range(nvar)]
    dz, dh = torch.from_numpy(z).cuda(), torch.from_numpy(hgt).cuda()
    _lib.check(lib.mt_rows_to_levels(D.ptr_array(d_rows), D.ptr_array(d_out), nvar, n, nlev,
                                     D.ptr(dz) if mask else None, D.ptr(dh) if mask else None, D.stream()),
               "mt_rows_to_levels")
    for r, o in zip(rows, d_out):
        ref = r.T.copy()
        if mask:
            ref = np.where(z[:, None] <= hgt[None, :], np.nan, ref)
        assert_same(o.cpu().numpy(), ref, f"rows_to_levels n={n} nlev={nlev}")
    # contract errors: too many levels, z without hgt
    assert lib.mt_rows_to_levels(D.ptr_array(d_rows), D.ptr_array(d_out), nvar, n, 97, None, None, D.stream()) != 0
    assert lib.mt_rows_to_levels(D.ptr_array(d_rows), D.ptr_array(d_out), nvar, n, nlev, D.ptr(dz), None,
                                 D.stream()) != 0
    assert b"mt_rows_to_levels" in lib.mt_last_error()


@pytest.mark.parametrize("nt,nr,with_hgt", [(2, 2, True), (5, 3, True), (48, 40, True), (360, 300, True), (37, 11, False)])
def test_cyl_prelude_abi(nt, nr, with_hgt):
    """mt_cyl_prelude (one launch) == mt_interp2d_layers per field + mt_terrain_index_cyl (index and the
    edge / corner fix-up by the last block), bit for bit, launched repeatedly (the ticket is left zero);
    NaN targets get the flagged fill values."""
    import torch
    from meteotools_b200 import _device as D, _lib
    lib = _lib.load()
    rng = np.random.default_rng(nt * 1000 + nr)
    ny, nx, dx = 57, 64, 2000.0
    n = nt * nr
    xs = rng.random(n) * (nx - 1) * dx
    ys = rng.random(n) * (ny - 1) * dx
    xs[0], ys[0] = 3 * dx, 5 * dx               # exactly on a node
    if n > 6:
        xs[5] = np.nan
        ys[5] = np.nan
    xd, yd = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda()
    fields = [rng.standard_normal((ny, nx)), 1500.0 * rng.random((ny, nx)), rng.standard_normal((ny, nx))]
    src = [torch.from_numpy(f).cuda() for f in fields]
    ix = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    w = torch.empty((n, 2), dtype=torch.float64, device="cuda")
    _lib.check(lib.mt_interp2d_plan(D.ptr(xd), D.ptr(yd), n, dx, dx, nx, ny, D.ptr(ix), D.ptr(w), D.stream()), "plan")
    z_start, dz = 100.0, 300.0
    ref = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in fields]
    for s_, o in zip(src, ref):
        _lib.check(lib.mt_interp2d_layers(D.ptr(s_), 1, ny, nx, 0, nx, 1, dx, dx, D.ptr(xd), D.ptr(yd), n, D.ptr(o), 0, 1,
                                          D.stream()), "mt_interp2d_layers")
    ref_idx = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(lib.mt_terrain_index_cyl(D.ptr(ref[1]), nt, nr, z_start, dz, D.ptr(ref_idx), D.stream()), "terrain")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    for rep in range(3):
        out = [torch.full((n,), -1.0, dtype=torch.float64, device="cuda") for _ in fields]
        idx = torch.full((n,), -5, dtype=torch.int32, device="cuda")
        _lib.check(lib.mt_cyl_prelude(D.ptr_array(src), D.ptr_array(out), len(src), ny, nx, D.ptr(ix), D.ptr(w), nt, nr,
                                      1 if with_hgt else -1, z_start, dz, D.ptr(idx) if with_hgt else None,
                                      D.ptr(ticket) if with_hgt else None, D.stream()), "mt_cyl_prelude")
        for o, r in zip(out, ref):
            assert_same(o.cpu().numpy(), r.cpu().numpy(), f"prelude field rep {rep}")
        if with_hgt:
            assert np.array_equal(idx.cpu().numpy(), ref_idx.cpu().numpy())
            assert int(ticket.item()) == 0
    assert lib.mt_cyl_prelude(D.ptr_array(src), D.ptr_array(out), 5, ny, nx, D.ptr(ix), D.ptr(w), nt, nr, -1, z_start, dz,
                              None, None, D.stream()) != 0


@pytest.mark.parametrize("nz,ny,nx,nlev,dt_", [(12, 9, 14, 70, np.float64), (12, 9, 13, 70, np.float64),
                                               (30, 33, 40, 64, np.float64), (30, 33, 41, 5, np.float64),
                                               (7, 20, 36, 33, np.float32), (40, 64, 64, 96, np.float64)])
def test_interplevel_level_major_modes(nz, ny, nx, nlev, dt_):
    """interpolation.interplevel (level-major result) through every mode of the tile kernel: register
    brackets (<= 64 levels) and shared-memory tables (more), 16-byte tile copies (even nx) and 8-byte
    ones (odd nx), float32 sources, pressure (decreasing) coordinate with descending levels -- bit for
    bit against the oracle."""
    from meteotools_b200.interpolation import interplevel
    rng = np.random.default_rng(nz * 100 + nx)
    hgt = 600.0 * rng.random((ny, nx))
    k = np.arange(nz).reshape(-1, 1, 1)
    z3d = hgt[None] + (15000.0 - hgt[None]) * (k / (nz - 1)) ** 1.3
    fld = (rng.standard_normal((nz, ny, nx)) * 3 + 0.002 * z3d).astype(dt_)
    levels = np.linspace(50.0, 16000.0, nlev)
    got = interplevel(fld, z3d.astype(dt_), levels)
    ref = oracle.interplevel(fld, z3d.astype(dt_), levels)
    assert got.shape == (nlev, ny, nx) and got.dtype == ref.dtype
    assert_same(np.asarray(got), ref, f"interplevel z {(nz, ny, nx, nlev)}")
    assert np.isnan(ref).any() and not np.isnan(ref).all()
    p3d = (1e5 * np.exp(-z3d / 8000.0)).astype(dt_)
    plev = np.linspace(99000.0, 12000.0, nlev)
    assert_same(np.asarray(interplevel(fld, p3d, plev)), oracle.interplevel(fld, p3d, plev),
                f"interplevel p {(nz, ny, nx, nlev)}")


# ---------------------------------------------------------------------------------------------------------
# The reference's methods use self.vr / self.vt AS THEY ARE once they exist (cylindrical_grid.py:290-508): a
# grid that holds only vr, vt (user-assigned, storm-relative ...) needs no u, v; a modified vt must be honoured.
def _vrvt_case(seed=5, nz=8, nt=12, nr=10):
    rng = np.random.default_rng(seed)
    st = dict(Nr=nr, Ntheta=nt, Nz=nz, dr=1000, dtheta=2 * np.pi / nt, dz=250, z_start=100, r_start=0,
              theta_start=0, vertical_coord_type="z", dt=300)
    shp = (nz, nt, nr)
    zz = (np.arange(nz) * 250.0 + 100.0).reshape(-1, 1, 1)
    f = {k: 10 * rng.standard_normal(shp) for k in ("vr", "vt", "w")}
    f["p"] = 1e5 * np.exp(-zz / 8000.0) + rng.standard_normal(shp)
    f["T"] = 300 - 6.5e-3 * zz + rng.standard_normal(shp)
    f["qv"] = 1e-3 + 1e-3 * np.abs(rng.standard_normal(shp))
    f["qc"] = 1e-4 * np.abs(rng.standard_normal(shp))
    f["qr"] = 1e-4 * np.abs(rng.standard_normal(shp))
    return st, f


VRVT_METHODS = ("calc_aam", "calc_agradient_force", "calc_kinetic_energy", "calc_vorticity_3D",
                "calc_abs_vorticity_3D", "calc_divergence_hori", "calc_divergence",
                "calc_density_potential_temperature", "calc_PV", "calc_deformation", "calc_filamentation")
VRVT_EXACT = ["coriolis_force", "centrifugal_force", "radial_PG_force", "agradient_force", "kinetic_energy",
              "zeta_r", "zeta_t", "zeta_z", "abs_zeta_z", "div_hori", "div", "shear_deform", "stretch_deform",
              "strain_region"]


@pytest.mark.parametrize("f_kind", ["field", "scalar"])
def test_cyl_methods_use_the_grids_vr_vt_without_u_v(f_kind):
    from meteotools_b200.grids import cylindrical_grid
    st, f = _vrvt_case()
    cyl = cylindrical_grid(dict(st))
    for k, v in f.items():
        setattr(cyl, k, v.copy())
    if f_kind == "scalar":      # a scalar Coriolis parameter broadcasts, as in the reference's self.f*self.vt
        cyl.f = 5e-5
        fc = 5e-5
    else:
        fc = 5e-5 + 1e-7 * np.arange(st["Ntheta"] * st["Nr"], dtype=np.float64).reshape(st["Ntheta"], st["Nr"])
        cyl.f = fc
    for m in VRVT_METHODS:
        getattr(cyl, m)()
    assert not cyl._has("u") and not cyl._has("v")
    with np.errstate(all="ignore"):
        ref = co.cyl_d_set(dict(f, f=fc), dict(r=cyl.r, theta=cyl.theta, dr=cyl.dr, dtheta=cyl.dtheta, dz=cyl.dz))
    for k in VRVT_EXACT:
        assert_same(getattr(cyl, k), ref[k], f"{k} from the grid's vr / vt ({f_kind} f)")
    r = cyl.r.reshape(1, 1, -1)
    assert_same(cyl.aam, f["vt"] * r + 0.5 * fc * r ** 2, "aam")
    assert_same(cyl.vt, f["vt"], "vt untouched")
    assert_close(cyl.PV, ref["PV"], 1e-10, "PV", nonfinite="class", floor=1e-3)
    with pytest.raises(AttributeError):
        cyl.calc_wind_speed()       # needs u, v (cylindrical_grid.py:354-358)


def test_cyl_modified_vt_is_honoured():
    """u, v present AND a user-modified vt: the reference differentiates self.vt, not a recomputed one."""
    from meteotools_b200.grids import cylindrical_grid
    st, f = _vrvt_case(seed=6)
    rng = np.random.default_rng(7)
    cyl = cylindrical_grid(dict(st))
    u, v = 10 * rng.standard_normal(f["w"].shape), 10 * rng.standard_normal(f["w"].shape)
    cyl.u, cyl.v, cyl.w, cyl.f = u, v, f["w"].copy(), 5e-5
    cyl.calc_vr_vt()
    vt_rel = np.array(cyl.vt) - 5.0          # storm-relative tangential wind
    cyl.vt = vt_rel
    cyl.calc_vorticity_3D()
    cyl.calc_aam()
    th = cyl.theta.reshape(1, -1, 1)
    vr = v * np.sin(th) + u * np.cos(th)
    with np.errstate(all="ignore"):
        zeta_r = co.FD2(f["w"], cyl.dtheta, 1) / cyl.r - co.FD2(vt_rel, cyl.dz, 0)
        zeta_z = co.FD2(vt_rel * cyl.r, cyl.dr, 2) / cyl.r - co.FD2(vr, cyl.dtheta, 1) / cyl.r
    assert_same(cyl.zeta_r, zeta_r, "zeta_r from the modified vt")
    assert_same(cyl.zeta_z, zeta_z, "zeta_z from the modified vt")
    assert_same(cyl.vt, vt_rel, "vt untouched by calc_aam")
    r = cyl.r.reshape(1, 1, -1)
    assert_same(cyl.aam, vt_rel * r + 0.5 * 5e-5 * r ** 2, "aam")


def test_cyl_stale_plan_is_dropped_when_targets_are_assigned():
    """cyl.xTR / yTR may be assigned directly (the reference permits it): the cached stencil plan follows."""
    from meteotools_b200.grids import cylindrical_grid
    rng = np.random.default_rng(3)
    st = dict(Nr=12, Ntheta=16, Nz=6, dr=2000, dtheta=2 * np.pi / 16, dz=500, z_start=100, r_start=0,
              theta_start=0, vertical_coord_type="z", dt=300)
    nz, ny, nx, dx = 8, 40, 40, 2000.0
    z3d = np.broadcast_to((np.arange(nz) * 600.0).reshape(-1, 1, 1), (nz, ny, nx)).copy()
    fld = rng.standard_normal((nz, ny, nx))
    cyl = cylindrical_grid(dict(st))
    cyl.set_horizontal_location(39000.0, 39000.0)
    cyl.set_data.__wrapped__(cyl, dx, dx, z3d, u=fld)
    x2, y2 = np.array(cyl.xTR) + 3000.0, np.array(cyl.yTR) - 1000.0
    cyl.xTR, cyl.yTR = x2, y2                # direct assignment
    cyl.set_data.__wrapped__(cyl, dx, dx, z3d, u=fld)
    lev = oracle.interplevel(fld, z3d, cyl.z)
    assert_same(cyl.u, oracle.interp2D_fast_layers(dx, dx, x2, y2, lev), "u after re-targeting")


def test_cyl_rotate_and_rmw10_follow_the_reference():
    from meteotools_b200.grids import cylindrical_grid
    st, f = _vrvt_case(seed=8)
    cyl = cylindrical_grid(dict(st))
    cyl.vt = f["vt"].copy()
    cyl.rotate_to_shear_direction("vt", 95.0, unit="deg")
    idx = int(np.deg2rad(95.0) / cyl.dtheta)
    assert_same(cyl.shear_relative_vt, np.roll(f["vt"], idx, axis=1), "shear_relative_vt")
    cyl.u10 = f["vt"][0].copy()
    cyl.v10 = f["vr"][0].copy()
    with pytest.raises(np.exceptions.AxisError):     # the reference's nanargmax(axis=1) on a 1-D array
        cyl.find_rmw_and_speed10()


@pytest.mark.parametrize("src_dtype,fac_dtype", [(np.float32, np.float32), (np.float64, np.float32),
                                                  (np.float32, np.float64), (np.float64, np.float64)])
def test_set_data_map_factors_match_correct_map_scale(src_dtype, fac_dtype):
    """wrfout_grid.correct_map_scale (grids/wrfout_grid.py:48-55: ua *= MAPFAC_MX, u10 *= MAPFAC_MX)
    folded into the staging of set_data; reference = the NumPy in-place products, then the CPU pipeline."""
    from meteotools_b200.grids import cylindrical_grid
    rng = np.random.default_rng(9)
    st = dict(Nr=14, Ntheta=16, Nz=6, dr=2000, dtheta=2 * np.pi / 16, dz=500, z_start=100, r_start=0,
              theta_start=0, vertical_coord_type="z", dt=300)
    nz, ny, nx, dx = 8, 44, 46, 2000.0
    z3d = np.broadcast_to((np.arange(nz) * 600.0).reshape(-1, 1, 1), (nz, ny, nx)).astype(src_dtype)
    ua = (10 * rng.standard_normal((nz, ny, nx))).astype(src_dtype)
    va = (10 * rng.standard_normal((nz, ny, nx))).astype(src_dtype)
    u10 = (10 * rng.standard_normal((ny, nx))).astype(src_dtype)
    mx = (1.0 + 0.05 * rng.random((ny, nx))).astype(fac_dtype)
    my = (1.0 + 0.05 * rng.random((ny, nx))).astype(fac_dtype)
    cyl = cylindrical_grid(dict(st))
    cyl.set_horizontal_location(45000.0, 43000.0)
    cyl.set_data.__wrapped__(cyl, dx, dx, z3d, u=ua, v=va, u10=u10, map_factors=dict(u=mx, v=my, u10=mx))
    ua_s, va_s, u10_s = ua.copy(), va.copy(), u10.copy()
    ua_s *= mx          # exactly what correct_map_scale does
    va_s *= my
    u10_s *= mx
    for name, src in (("u", ua_s), ("v", va_s)):
        lev = oracle.interplevel(src, z3d, cyl.z)
        assert_same(getattr(cyl, name), oracle.interp2D_fast_layers(dx, dx, cyl.xTR, cyl.yTR, lev), f"scaled {name}")
    assert_same(cyl.u10, oracle.interp2D_fast(dx, dx, cyl.xTR, cyl.yTR, u10_s), "scaled u10")
    assert not np.array_equal(ua_s, ua)
