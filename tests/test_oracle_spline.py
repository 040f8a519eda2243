"""SURVEY 8(f)-3 on the CPU: the oracle's restatement of the reference's ``spline`` method (scipy
splrep + splev per column) and of the set_data pipeline around it, against vectors produced by
the UNMODIFIED reference grid classes (tests/golden/spline.npz, generator: make_golden.gen_spline)."""
import numpy as np

import oracle
from conftest import assert_same, load_golden
from golden.make_golden import SD
from oracle import calc_oracle as co


def test_spline_method_matches_reference():
    g = load_golden("spline")
    sub = (slice(None), slice(3, 6), slice(4, 9))
    u, z = g["s_u"][sub], g["s_z3d"][sub]
    assert_same(oracle.spline_levels(u, z, g["s_lev"], 2), g["o_spline_k2"], "k=2")
    assert_same(oracle.spline_levels(u, z, g["s_lev"], 3), g["o_spline_k3"], "k=3")
    assert_same(oracle.spline_levels(u.astype(np.float32), z.astype(np.float32), g["s_lev"], 3),
                g["o_spline_k3_f32"], "k=3 float32")
    assert g["o_spline_k3"].dtype == np.float64 and g["o_spline_k3_f32"].dtype == np.float64


def test_spline_set_data_pipeline_matches_reference():
    """spline -> interp2D_fast_layers -> terrain mask == reference cylindrical_grid.set_data(ver_interp_order=2)."""
    g = load_golden("spline")
    from meteotools_b200.calc import Make_cyclinder_coord
    dx = float(g["s_dx"])
    r = np.arange(SD["r_start"], SD["Nr"] * SD["dr"], SD["dr"], dtype=float)
    theta = np.arange(SD["theta_start"], SD["Ntheta"] * SD["dtheta"], SD["dtheta"])
    z = np.arange(0, SD["Nz"] * SD["dz"], SD["dz"], dtype=float) + SD["z_start"]
    xTR, yTR = Make_cyclinder_coord([float(g["s_cy"]), float(g["s_cx"])], r, theta)
    lev = oracle.spline_levels(g["s_u"], g["s_z3d"], z, 2)
    ref = oracle.interp2D_fast_layers(dx, dx, xTR, yTR, lev)
    ref = co.cyl_terrain_mask(ref, g["o_cyl_hgt_idx"])
    assert_same(ref, g["o_cyl_u_k2"], "cyl u, k=2")


def test_set_data_order_contract_without_gpu():
    """Argument contract of the spline branch, checked on the host: an order the reference does not
    handle ends in NameError (its `tmp` stays undefined, cylindrical_grid.py:110-127); a supported
    order without a CUDA device fails loudly (no CPU fallback)."""
    import pytest
    from meteotools_b200.exceptions import BackendError
    from meteotools_b200.grids import cartesian_grid, cylindrical_grid
    import torch
    cyl = cylindrical_grid(dict(SD))
    cyl.set_horizontal_location(40000.0, 40000.0)
    z = np.zeros((5, 40, 40))
    with pytest.raises(NameError):
        cyl.set_data(2000.0, 2000.0, z, ver_interp_order=3, u=z)
    car = cartesian_grid(dict(Nx=10, Ny=9, Nz=10, dx=3000, dy=3000, dz=600, z_start=100, vertical_coord_type="z",
                              SOR_parameter=1.9, max_iteration_times=10, iteration_tolerance=0.05))
    car.set_horizontal_location(40000.0, 40000.0)
    with pytest.raises(NameError):
        car.set_data(2000.0, 2000.0, z, vertical_interp_order=4, u=z)
    if not torch.cuda.is_available():
        with pytest.raises(BackendError):
            cyl.set_data(2000.0, 2000.0, z, ver_interp_order=2, u=z)
