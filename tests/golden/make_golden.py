"""Generate tests/golden/*.npz by running the UNMODIFIED reference package.

Runs only in the build container (needs /root/reference).  The reference's Python is
imported from a temporary copy (the reference tree is read-only and its import needs
``lib/fastcompute.so``); that library is the oracle's C restatement of fastcompute.f90
(gfortran is unavailable; when ``oracle/_ref/fastcompute.so`` exists it is used instead).
``wrf`` and ``netCDF4`` are absent from the image and are stubbed: ``wrf.interplevel`` is
the oracle's DINTERP3DZ restatement (parity unpinned -- see oracle/fastcompute_oracle.c).

Everything else -- ``meteotools.interpolation`` wrappers, ``meteotools.calc`` and the
``meteotools.grids`` diagnostics -- is the reference's own code, so these vectors pin
the oracle's numpy restatement (oracle/calc_oracle.py, oracle/__init__.py) and the CUDA
path to the reference itself.

    python tests/golden/make_golden.py
"""
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference/meteotools"


def import_reference():
    tmp = tempfile.mkdtemp(prefix="mt_ref_")
    dst = os.path.join(tmp, "meteotools")
    shutil.copytree(REF, dst, ignore=shutil.ignore_patterns("colors", "old_code", "*.png", "*.html"))
    os.makedirs(os.path.join(dst, "lib"), exist_ok=True)
    so = oracle._REF_SO if os.path.exists(oracle._REF_SO) else oracle.build()
    shutil.copy(so, os.path.join(dst, "lib", "fastcompute.so"))
    wrf = types.ModuleType("wrf")
    wrf.omp_set_num_threads = lambda n: None
    wrf.interplevel = lambda var, vert, lev: oracle.interplevel(var, vert, lev)
    nc = types.ModuleType("netCDF4")
    nc.Dataset = object
    sys.modules["wrf"], sys.modules["netCDF4"] = wrf, nc
    sys.path.insert(0, tmp)
    import meteotools  # noqa: F401
    import meteotools.calc
    import meteotools.grids
    import meteotools.interpolation
    return sys.modules["meteotools"]


def synth_cyl_fields(rng, nz, nt, nr, z):
    """SURVEY 8(d) S3 distributions."""
    shp = (nz, nt, nr)
    zz = z.reshape(-1, 1, 1)
    f = {}
    for k in ("u", "v", "w"):
        f[k] = 10 * rng.standard_normal(shp)
    f["p"] = 1e5 * np.exp(-zz / 8000.0) + rng.standard_normal(shp)
    f["T"] = 300 - 6.5e-3 * zz + rng.standard_normal(shp)
    f["qv"] = 1e-3 + 1e-3 * np.abs(rng.standard_normal(shp))
    for k in ("qc", "qr", "qi", "qs", "qg"):
        f[k] = 1e-4 * np.abs(rng.standard_normal(shp))
    return f


def gen_interp(mt, out):
    I = mt.interpolation
    rng = np.random.default_rng(1)
    dx, dy, dz = 1.0, 3.0, 5.0
    x = np.arange(0, 30, 1.0)
    y = np.arange(0, 12, 1.0)
    z = np.arange(0, 7, 1.0)
    t = np.arange(0, 3, 1.0)
    xn = rng.random((6, 5)) * 29.0 * dx
    yn = rng.random((6, 5)) * 11.0 * dy
    zn = rng.random((6, 5)) * 6.0 * dz
    # exact nodes, last nodes and NaN targets
    xn[0, 0], yn[0, 0], zn[0, 0] = 29.0 * dx, 11.0 * dy, 6.0 * dz
    xn[0, 1], yn[0, 1], zn[0, 1] = 4.0 * dx, 3.0 * dy, 2.0 * dz
    xn[0, 2] = np.nan
    yn[0, 3] = np.nan
    a1 = rng.standard_normal(30)
    a1l = rng.standard_normal((3, 2, 30))
    a2 = rng.standard_normal((12, 30))
    a2l = rng.standard_normal((3, 2, 12, 30))
    a2l[1, 1, 5, 7] = np.nan
    a3 = rng.standard_normal((7, 12, 30))
    a3l = rng.standard_normal((3, 7, 12, 30))
    xin = np.sort(rng.random(30) * 100.0)
    xin_dec = xin[::-1].copy()
    xo = xin[0] + rng.random((4, 5)) * (xin[-1] - xin[0])
    xo[0, 0], xo[0, 1], xo[0, 2] = xin[0], xin[-1], np.nan
    out.update(
        i_dx=dx, i_dy=dy, i_dz=dz, i_xn=xn, i_yn=yn, i_zn=zn, i_a1=a1, i_a1l=a1l, i_a2=a2,
        i_a2l=a2l, i_a3=a3, i_a3l=a3l, i_xin=xin, i_xin_dec=xin_dec, i_xo=xo,
        o_interp1D_fast=I.interp1D_fast(dx, xn, a1),
        o_interp1D_fast_layers=I.interp1D_fast_layers(dx, xn, a1l),
        o_interp1D_nonequal=I.interp1D_nonequal(xin, xo, a1),
        o_interp1D_nonequal_dec=I.interp1D_nonequal(xin_dec, xo, a1),
        o_interp1D_nonequal_layers=I.interp1D_nonequal_layers(xin, xo, a1l),
        o_interp2D_fast=I.interp2D_fast(dx, dy, xn, yn, a2),
        o_interp2D_fast_layers=I.interp2D_fast_layers(dx, dy, xn, yn, a2l),
        o_interp3D_fast=I.interp3D_fast(dx, dy, dz, xn, yn, zn, a3),
        o_interp3D_fast_layers=I.interp3D_fast_layers(dx, dy, dz, xn, yn, zn, a3l),
    )
    # test_calc.py:58-78 cartesian2cylindrical KAT, reduced in z
    r = np.arange(0, 100 + 0.0001, 1)
    theta = np.linspace(0, np.pi * 2, 361)[:-1]
    data = np.add.outer(np.arange(3.0), np.add.outer(np.arange(251.0), np.arange(251.0)))
    out["o_c2c"] = mt.calc.cartesian2cylindrical(1, 1, data, [125, 125], r, theta)


def gen_calc(mt, out):
    C = mt.calc
    rng = np.random.default_rng(2)
    a = rng.standard_normal((5, 6, 7))
    out["c_a"] = a
    for name in ("FD2", "FD2_front", "FD2_back", "FD2_2", "FD2_2_front", "FD2_2_back"):
        for ax in range(3):
            out[f"o_{name}_{ax}"] = getattr(C, name)(a, 0.7, ax)
    from meteotools.calc.differential import FD4
    for ax in range(3):
        out[f"o_FD4_{ax}"] = FD4(a, 0.7, ax)
    shp = (4, 5, 6)
    zz = np.linspace(100, 12000, 4).reshape(-1, 1, 1)
    T = 300 - 6.5e-3 * zz + rng.standard_normal(shp)
    P = 1e5 * np.exp(-zz / 8000.0) + rng.standard_normal(shp)
    qv = 1e-3 + 1e-3 * np.abs(rng.standard_normal(shp))
    ql = 1e-4 * np.abs(rng.standard_normal(shp))
    RH = 0.2 + 0.7 * rng.random(shp)
    out.update(t_T=T, t_P=P, t_qv=qv, t_ql=ql, t_RH=RH)
    th = C.calc_theta(T, P)
    vap = C.qv_to_vapor(P, qv)
    out.update(
        o_theta=th, o_T=C.calc_T(th, P), o_sat_vapor=C.calc_saturated_vapor(T),
        o_vapor_RH=C.calc_vapor(T, RH), o_sat_qv=C.calc_saturated_qv(T, P), o_qv_to_vapor=vap,
        o_theta_es=C.calc_theta_es(T, P), o_Td=C.calc_Td(es=vap), o_qv=C.calc_qv(P, vapor=vap),
        o_theta_v=C.calc_theta_v(theta=th, qv=qv, ql=ql), o_Tv=C.calc_Tv(T, qv=qv),
        o_Tv_vapor=C.calc_Tv(T, P=P, vapor=vap), o_rho=C.calc_rho(P, Tv=C.calc_Tv(T, qv=qv)),
        o_Tc=C.calc_Tc(T, P, qv=qv), o_theta_e=C.calc_theta_e(theta=th, qv=qv, Tc=C.calc_Tc(T, P, qv=qv)),
        o_dTdz=C.calc_dTdz(T=T, P=P), o_RH=C.calc_RH(T=T, vapor=vap),
    )
    Tn = T.copy()
    Tn[0, 0, 0] = np.nan  # NaN-sensitive global stop test: all 10 iterations run
    out["o_Tc_nan"] = C.calc_Tc(Tn, P, qv=qv)


CYL = dict(Nr=9, Ntheta=8, Nz=6, dr=1000, dtheta=2 * np.pi / 8, dz=250, z_start=100, r_start=0,
           theta_start=0, vertical_coord_type="z", dt=300)
CAR = dict(Nx=9, Ny=8, Nz=6, dx=2000, dy=2000, dz=250, z_start=100, vertical_coord_type="z",
           SOR_parameter=1.9, max_iteration_times=10, iteration_tolerance=0.05)
CYL_OUT = ["vr", "vt", "wind_speed", "Tv", "rho", "coriolis_force", "centrifugal_force",
           "radial_PG_force", "agradient_force", "kinetic_energy", "zeta_r", "zeta_t", "zeta_z",
           "abs_zeta_z", "div_hori", "div", "density_factor", "Th", "Th_rho", "PV", "shear_deform",
           "stretch_deform", "filamentation_time", "strain_region", "sym_vt", "asy_vt", "RMW",
           "wspd_max_RMW", "vt_max_RMW", "axisymmetry_vt"]
CAR_OUT = ["wind_speed", "kinetic_energy", "zeta_r", "zeta_t", "zeta_z", "abs_zeta_z", "div_hori",
           "div", "density_factor", "Th", "Th_rho", "Tv", "rho", "PV", "shear_deform",
           "stretch_deform", "filamentation_time"]
TD_OUT = ["Th", "vapor_s", "qvs", "vapor", "RH", "Td", "Th_es", "Th_v", "Tv", "rho", "Tc", "Th_e",
          "moist_dTdz"]


def gen_grids(mt, out):
    from meteotools.grids import cartesian_grid, cylindrical_grid
    rng = np.random.default_rng(3)
    with np.errstate(all="ignore"):
        cyl = cylindrical_grid(dict(CYL))
        f = synth_cyl_fields(rng, CYL["Nz"], CYL["Ntheta"], CYL["Nr"], cyl.z)
        f["f"] = 5e-5 + 1e-6 * rng.standard_normal((CYL["Ntheta"], CYL["Nr"]))
        for k, v in f.items():
            setattr(cyl, k, v.copy())
            out[f"g_cyl_{k}"] = v
        for m in ("calc_vr_vt", "calc_wind_speed", "find_rmw_and_speed", "calc_agradient_force",
                  "calc_kinetic_energy", "calc_vorticity_3D", "calc_abs_vorticity_3D",
                  "calc_divergence_hori", "calc_divergence", "calc_density_potential_temperature",
                  "calc_PV", "calc_deformation", "calc_filamentation"):
            getattr(cyl, m)()
        cyl.calc_axisymmetry("vt")
        for k in CYL_OUT:
            out[f"o_cyl_{k}"] = np.asarray(getattr(cyl, k))
        # thermodynamic chain on a fresh grid holding T, p, qv, qc, qr
        td = cylindrical_grid(dict(CYL))
        for k in ("T", "p", "qv", "qc", "qr"):
            setattr(td, k, f[k].copy())
        for m in ("calc_theta", "calc_saturated_vapor", "calc_saturated_qv", "calc_vapor", "calc_RH",
                  "calc_Td", "calc_theta_es", "calc_theta_v", "calc_Tv", "calc_rho", "calc_Tc",
                  "calc_theta_e"):
            getattr(td, m)()
        td.moist_dTdz = mt.calc.calc_dTdz(T=td.T, P=td.p)
        for k in TD_OUT:
            out[f"o_td_{k}"] = np.asarray(getattr(td, k))
        # Cartesian
        car = cartesian_grid(dict(CAR))
        fc = synth_cyl_fields(rng, CAR["Nz"], CAR["Ny"], CAR["Nx"], car.z)
        fc["f"] = 5e-5 + 1e-6 * rng.standard_normal((CAR["Ny"], CAR["Nx"]))
        for k, v in fc.items():
            setattr(car, k, v.copy())
            out[f"g_car_{k}"] = v
        for m in ("calc_wind_speed", "calc_kinetic_energy", "calc_vorticity_3D",
                  "calc_abs_vorticity_3D", "calc_divergence_hori", "calc_divergence",
                  "calc_density_potential_temperature", "calc_PV", "calc_deformation",
                  "calc_filamentation"):
            getattr(car, m)()
        for k in CAR_OUT:
            out[f"o_car_{k}"] = np.asarray(getattr(car, k))


SD = dict(Nr=12, Ntheta=16, Nz=10, dr=1500, dtheta=2 * np.pi / 16, dz=600, z_start=100, r_start=0,
          theta_start=0, vertical_coord_type="z", dt=300)


def synth_wrf(rng, nz, ny, nx, dx):
    """SURVEY 8(d) S2-style source: terrain-following heights, smooth + noise fields."""
    yy, xx = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
    cx, cy = (nx - 1) * dx / 2, (ny - 1) * dx / 2
    hgt = 1500.0 * np.exp(-((xx - 0.7 * cx) ** 2 + (yy - 1.2 * cy) ** 2) / (0.3 * cx) ** 2)
    k = np.arange(nz).reshape(-1, 1, 1)
    z3d = hgt[None] + (20000.0 - hgt[None]) * (k / (nz - 1)) ** 1.5
    fields = {}
    for name in ("u", "T"):
        fields[name] = (10 * np.sin(2 * np.pi * xx / (nx * dx)) * np.cos(2 * np.pi * yy / (ny * dx))[None]
                        + 0.01 * z3d + 0.1 * rng.standard_normal((nz, ny, nx)))
    return hgt, z3d, fields


def gen_setdata(mt, out):
    from meteotools.grids import cartesian_grid, cylindrical_grid
    rng = np.random.default_rng(4)
    nz, ny, nx, dx = 14, 40, 44, 2000.0
    hgt, z3d, fields = synth_wrf(rng, nz, ny, nx, dx)
    cxl, cyl_ = ((nx + 1) / 2 - 1) * dx, ((ny + 1) / 2 - 1) * dx   # wrfout_grid.py:69-70
    f2d = 5e-5 + 1e-7 * rng.standard_normal((ny, nx))
    out.update(s_hgt=hgt, s_z3d=z3d, s_u=fields["u"], s_T=fields["T"], s_f=f2d, s_dx=dx,
               s_cx=cxl, s_cy=cyl_)
    cyl = cylindrical_grid(dict(SD))
    cyl.set_horizontal_location(cxl, cyl_)
    cyl.set_data(dx, dx, z3d, u=fields["u"], T=fields["T"], f=f2d, hgt=hgt)
    out.update(o_sd_cyl_u=cyl.u, o_sd_cyl_T=cyl.T, o_sd_cyl_f=cyl.f, o_sd_cyl_hgt=cyl.hgt,
               o_sd_cyl_hgt_idx=cyl.hgt_idx, o_sd_xTR=cyl.xTR, o_sd_yTR=cyl.yTR)
    cyl2 = cylindrical_grid(dict(SD))                       # without terrain masking
    cyl2.set_horizontal_location(cxl, cyl_)
    cyl2.set_data(dx, dx, z3d, u=fields["u"])
    out["o_sd_cyl_u_nomask"] = cyl2.u
    car = cartesian_grid(dict(Nx=10, Ny=9, Nz=10, dx=3000, dy=3000, dz=600, z_start=100,
                              vertical_coord_type="z", SOR_parameter=1.9, max_iteration_times=10,
                              iteration_tolerance=0.05))
    car.set_horizontal_location(cxl, cyl_)
    car.set_data(dx, dx, z3d, u=fields["u"], hgt=hgt)
    out.update(o_sd_car_u=car.u, o_sd_car_hgt=car.hgt)
    # pressure-type (decreasing) coordinate through the interplevel stub
    p3d = 1e5 * np.exp(-z3d / 8000.0)
    plev = np.array([95000.0, 85000.0, 70000.0, 50000.0, 30000.0, 20000.0])
    out.update(s_p3d=p3d, s_plev=plev, o_interplevel_p=oracle.interplevel(fields["T"], p3d, plev),
               o_interplevel_z=oracle.interplevel(fields["T"], z3d, cyl.z))


RED_CYL = dict(Nr=14, Ntheta=12, Nz=9, dr=1000, dtheta=2 * np.pi / 12, dz=250, z_start=5, r_start=0,
               theta_start=0, vertical_coord_type="z", dt=300)
RED_CAR = dict(Nx=11, Ny=10, Nz=7, dx=2000, dy=2000, dz=250, z_start=100, vertical_coord_type="z",
               SOR_parameter=1.9, max_iteration_times=10, iteration_tolerance=0.05)


def gen_reductions(mt, out):
    """range averages, axisymmetry, integrated kinetic energy, smoothers --
    outputs of the unmodified reference functions / grid methods, on fields with terrain-like NaNs."""
    from meteotools.grids import cartesian_grid, cylindrical_grid
    C = mt.calc
    rng = np.random.default_rng(5)
    with np.errstate(all="ignore"):
        # calc.averages on plain arrays
        a = rng.standard_normal((6, 40, 5))
        a[1, 3:9, 2] = np.nan
        a[:, 0, 0] = np.nan
        ax = np.cumsum(0.5 + rng.random(40))
        out.update(r_a=a, r_ax=ax)
        out["o_Zavg_1"] = C.calc_Zaverage(a, ax, 1, ax[4], ax[30])
        out["o_Zavg_1_full"] = C.calc_Zaverage(a, ax, 1)
        out["o_Ravg_1"] = C.calc_Raverage(a, ax, 1, ax[2], ax[35])
        a0 = np.ascontiguousarray(np.moveaxis(a, 1, 0))          # (40, 6, 5): axis 0
        out["o_Zavg_0"] = C.calc_Zaverage(a0, ax, 0, ax[4], ax[30])
        a2 = np.ascontiguousarray(np.moveaxis(a, 1, 2))          # (6, 5, 40): contiguous axis
        out["o_Zavg_2"] = C.calc_Zaverage(a2, ax, 2, ax[4], ax[30])
        out["o_Ravg_2"] = C.calc_Raverage(a2, ax, 2)
        out["o_Zavg_1d"] = C.calc_Zaverage(a[2, :, 1], ax, 0, ax[1], ax[20])
        zr = rng.standard_normal((9, 40))
        zr[2, 5:8] = np.nan
        zr[:, 11] = np.nan
        zax = np.arange(9.0) * 250 + 5
        out.update(r_zr=zr, r_zax=zax)
        out["o_RZavg"] = C.calc_RZaverage(zr, zax, ax, ax[3], ax[33], zax[1], zax[6])
        out["o_RZavg_now"] = C.calc_RZaverage(zr, zax, ax, ax[3], ax[33], zax[1], zax[6], r_weight=False)
        # cylindrical grid methods
        cyl = cylindrical_grid(dict(RED_CYL))
        f = synth_cyl_fields(rng, RED_CYL["Nz"], RED_CYL["Ntheta"], RED_CYL["Nr"], cyl.z)
        for k in f:                                               # terrain-like NaNs
            f[k][:2, 3:7, 8:] = np.nan
        f["f"] = 5e-5 + 1e-6 * rng.standard_normal((RED_CYL["Ntheta"], RED_CYL["Nr"]))
        f["u10"] = 8 * rng.standard_normal((RED_CYL["Ntheta"], RED_CYL["Nr"]))
        f["v10"] = 8 * rng.standard_normal((RED_CYL["Ntheta"], RED_CYL["Nr"]))
        for k, v in f.items():
            setattr(cyl, k, v.copy())
            out[f"g_rcyl_{k}"] = v
        cyl.calc_vr_vt()
        cyl.calc_axisymmetry("vt")
        cyl.calc_axisymmetry("f")
        cyl.prof = f["T"][:, 2, 3].copy()
        cyl.ring = f["f"][:, 4].copy()
        cyl.calc_axisymmetry("ring")
        cyl.calc_horizontal_average("T", 0.4, 4.0, 1500., 9000.)
        cyl.calc_horizontal_average("qv")
        cyl.calc_horizontal_average("f", 1, 9, 2000., 11000., unit="index")
        cyl.calc_vertical_average("T", 200., 1600.)
        cyl.calc_vertical_average("prof", 200., 1600.)
        cyl.calc_IKE(rmin=2000., rmax=10000., zmin=200., zmax=1800.)
        ike_a = cyl.IKE
        cyl.calc_IKE()
        out["o_rcyl_IKE_range"], out["o_rcyl_IKE"] = np.float64(ike_a), np.float64(cyl.IKE)
        for k in ("sym_vt", "asy_vt", "axisymmetry_vt", "sym_f", "asy_f", "axisymmetry_f", "sym_ring",
                  "asy_ring", "axisymmetry_ring", "havg_T", "havg_qv", "havg_f", "zavg_T", "zavg_prof"):
            out[f"o_rcyl_{k}"] = np.asarray(getattr(cyl, k))
        # calc_IKE10: rho computed inside (level nearest 10 m), default and explicit rho
        c10 = cylindrical_grid(dict(RED_CYL))
        for k in ("u10", "v10", "p", "T", "qv"):
            setattr(c10, k, f[k].copy())
        c10.calc_IKE10(rmax=9000., min_wspd10=3.0)
        out["o_rcyl_IKE10"] = np.float64(c10.IKE10)
        out["o_rcyl_ke10"] = np.asarray(c10.kinetic_energy10)
        c10b = cylindrical_grid(dict(RED_CYL))
        for k in ("u10", "v10", "p", "T", "qv"):
            setattr(c10b, k, f[k].copy())
        c10b.calc_IKE10(rho=1.1)
        out["o_rcyl_IKE10_rho"] = np.float64(c10b.IKE10)
        c10b.calc_wind_speed10()                                   # rho now exists: the argument is used as is
        del c10b.kinetic_energy10, c10b.vr10, c10b.vt10
        c10b.calc_IKE10()
        out["o_rcyl_IKE10_nanrho"] = np.float64(c10b.IKE10)
        # Cartesian grid: averages and smoothers
        car = cartesian_grid(dict(RED_CAR))
        fc = synth_cyl_fields(rng, RED_CAR["Nz"], RED_CAR["Ny"], RED_CAR["Nx"], car.z)
        fc["T"][:1, 2:5, 6:] = np.nan
        fc["f"] = 5e-5 + 1e-6 * rng.standard_normal((RED_CAR["Ny"], RED_CAR["Nx"]))
        for k in ("T", "u", "f"):
            setattr(car, k, fc[k].copy())
            out[f"g_rcar_{k}"] = fc[k]
        car.calc_horizontal_average("T", 2000., 16000., 4000., 14000.)
        car.calc_horizontal_average("f", ymin=2000.)
        car.calc_vertical_average("T", zmax=1200.)
        for k in ("havg_T", "havg_f", "zavg_T"):
            out[f"o_rcar_{k}"] = np.asarray(getattr(car, k))
        for ax_ in range(3):
            car.smooth1d("u", axis=ax_, center_weight=2, passes=3)
            out[f"o_rcar_smooth1d_u_{ax_}"] = np.asarray(car.smooth1d_u)
        car.smooth1d("u", axis=1, center_weight=1.5, passes=1)
        out["o_rcar_smooth1d_u_cw"] = np.asarray(car.smooth1d_u)
        car.smooth1d("T", axis=0, passes=2)                       # NaNs spread by one cell per pass
        out["o_rcar_smooth1d_T"] = np.asarray(car.smooth1d_T)
        for axes in ((1, 2), (0, 2), (0, 1)):
            car.smooth2d("u", axis=axes, center_weight=2, passes=4)
            out[f"o_rcar_smooth2d_u_{axes[0]}{axes[1]}"] = np.asarray(car.smooth2d_u)
        out["o_rcar_smooth2d_side"] = np.asarray(car.smooth1d_u)     # zeros of the moved shape (:183)
        car.smooth2d("f", axis=(0, 1), passes=2)
        out["o_rcar_smooth2d_f"] = np.asarray(car.smooth2d_f)
        car.smooth3d("u", center_weight=2, passes=10)
        out["o_rcar_smooth3d_u"] = np.asarray(car.smooth3d_u)
        car.smooth3d("u", center_weight=0.5, passes=1)
        out["o_rcar_smooth3d_u_cw"] = np.asarray(car.smooth3d_u)


def gen_spline(mt, out):
    """``ver_interp_order = 2`` (cylindrical) / ``vertical_interp_order = 2, 3``
    (Cartesian) through the reference's own set_data (multiprocessing pool + scipy splrep/splev)."""
    from meteotools.grids import cartesian_grid, cylindrical_grid
    rng = np.random.default_rng(7)
    nz, ny, nx, dx = 14, 40, 44, 2000.0
    hgt, z3d, fields = synth_wrf(rng, nz, ny, nx, dx)
    cxl, cyl_ = ((nx + 1) / 2 - 1) * dx, ((ny + 1) / 2 - 1) * dx
    out.update(s_hgt=hgt, s_z3d=z3d, s_u=fields["u"], s_T=fields["T"], s_dx=dx, s_cx=cxl, s_cy=cyl_)
    cyl = cylindrical_grid(dict(SD))
    cyl.set_horizontal_location(cxl, cyl_)
    cyl.set_data(dx, dx, z3d, ver_interp_order=2, u=fields["u"], T=fields["T"], hgt=hgt)
    out.update(o_cyl_u_k2=cyl.u, o_cyl_T_k2=cyl.T, o_cyl_hgt_idx=cyl.hgt_idx)
    for order in (2, 3):
        car = cartesian_grid(dict(Nx=10, Ny=9, Nz=10, dx=3000, dy=3000, dz=600, z_start=100,
                                  vertical_coord_type="z", SOR_parameter=1.9, max_iteration_times=10,
                                  iteration_tolerance=0.05))
        car.set_horizontal_location(cxl, cyl_)
        car.set_data(dx, dx, z3d, vertical_interp_order=order, u=fields["u"], hgt=hgt)
        out[f"o_car_u_k{order}"] = car.u
    # the method itself on a few columns, float32 input included
    sub = (slice(None), slice(3, 6), slice(4, 9))
    lev = np.arange(-200.0, 21000.0, 700.0)      # extrapolates below and above the column
    out["s_lev"] = lev
    out["o_spline_k2"] = cylindrical_grid.spline(None, fields["u"][sub], z3d[sub], lev)
    out["o_spline_k3"] = cartesian_grid.spline(None, fields["u"][sub], z3d[sub], lev, 3)
    out["o_spline_k3_f32"] = cartesian_grid.spline(None, fields["u"][sub].astype(np.float32),
                                                   z3d[sub].astype(np.float32), lev, 3)


def main():
    mt = import_reference()
    only = set(sys.argv[1:])
    for name, fn in (("interp", gen_interp), ("calc", gen_calc), ("grids", gen_grids),
                     ("setdata", gen_setdata), ("reductions", gen_reductions), ("spline", gen_spline)):
        if only and name not in only:
            continue
        out = {}
        fn(mt, out)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
