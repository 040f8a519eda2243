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
