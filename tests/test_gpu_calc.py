"""GPU parity of meteotools_b200.calc against golden outputs of the
reference and the oracle.  Finite differences: bit-exact.  Thermodynamics: 1e-10 relative
(CUDA's exp/log/pow differ from NumPy's by <= 1-2 ulp), identical NaN masks."""
import numpy as np
import pytest

from conftest import assert_close, assert_same
from oracle import calc_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    import meteotools_b200.calc as C
    return C


def test_fd_golden_bit_exact(C, golden_calc):
    g = golden_calc
    for name in ("FD2", "FD4", "FD2_front", "FD2_back", "FD2_2", "FD2_2_front", "FD2_2_back"):
        for ax in range(3):
            assert_same(getattr(C, name)(g["c_a"], 0.7, ax), g[f"o_{name}_{ax}"], f"{name} axis {ax}")
    assert_same(C.FD2(g["c_a"], 0.7, -1), g["o_FD2_2"], "negative axis")


def test_fd_kat_and_shapes(C):  # test_calc.py:81-91
    a = np.add.reduce(np.meshgrid(*[np.arange(n, dtype=float) for n in (4, 5, 6)], indexing="ij"))
    assert np.all(C.FD2(a, 2, 0) == 0.5)
    b = C.FD2_front(a, 2, 0)
    assert np.all(np.isnan(b[-2:])) and abs(np.nansum(b) - 30) < 1e-13
    rng = np.random.default_rng(3)
    for shape, ax in (((17,), 0), ((9, 33), 1), ((4, 3, 5, 6), 2), ((3, 1000, 7), 1)):
        v = rng.standard_normal(shape)
        v.flat[3] = np.nan
        assert_same(C.FD2(v, 1.3, ax), co.FD2(v, 1.3, ax), f"FD2 {shape} axis {ax}")
    with pytest.raises(IndexError):
        C.FD2(np.zeros((2, 5)), 1.0, 0)
    with pytest.raises(TypeError):
        C.FD2(np.zeros((5, 5)), 1.0, 0, cyclic=True)   # unreachable in the reference too


def test_fd_contiguous_axis_flat_and_tiled_orders(C):
    """The contiguous axis runs in flat order (thread <-> element of the whole array, csrc/fd.cu k_fd_row_flat)
    when its rows are not multiples of 32 doubles and in row tiles otherwise: every kind, row lengths on both
    sides of the switch, row counts that leave ragged last blocks."""
    rng = np.random.default_rng(11)
    for rows, n in ((1, 5), (7, 31), (3, 33), (129, 100), (5, 1000), (4, 64), (9, 1024)):
        v = rng.standard_normal((rows, n))
        v[rows // 2, n // 2] = np.nan
        for name in ("FD2", "FD4", "FD2_front", "FD2_back", "FD2_2", "FD2_2_front", "FD2_2_back"):
            assert_same(getattr(C, name)(v, 0.37, 1), getattr(co, name)(v, 0.37, 1), f"{name} ({rows}, {n})")


def test_thermo_golden(C, golden_calc):
    g = golden_calc
    T, P, qv, ql, RH = g["t_T"], g["t_P"], g["t_qv"], g["t_ql"], g["t_RH"]
    th = C.calc_theta(T, P)
    vap = C.qv_to_vapor(P, qv)
    assert_same(vap, g["o_qv_to_vapor"], "qv_to_vapor (pure arithmetic)")
    assert_same(C.calc_Tv(T, qv=qv), g["o_Tv"], "Tv (pure arithmetic)")
    assert_same(C.calc_rho(P, Tv=g["o_Tv"]), g["o_rho"], "rho (pure arithmetic)")
    assert_same(C.calc_qv(P, vapor=g["o_qv_to_vapor"]), g["o_qv"], "qv (pure arithmetic)")
    pairs = {
        "theta": th, "T": C.calc_T(g["o_theta"], P), "sat_vapor": C.calc_saturated_vapor(T),
        "vapor_RH": C.calc_vapor(T, RH), "sat_qv": C.calc_saturated_qv(T, P),
        "theta_es": C.calc_theta_es(T, P), "Td": C.calc_Td(es=g["o_qv_to_vapor"]),
        "theta_v": C.calc_theta_v(theta=g["o_theta"], qv=qv, ql=ql),
        "Tv_vapor": C.calc_Tv(T, P=P, vapor=g["o_qv_to_vapor"]),
        "Tc": C.calc_Tc(T, P, qv=qv), "theta_e": C.calc_theta_e(theta=g["o_theta"], qv=qv, Tc=g["o_Tc"]),
        "dTdz": C.calc_dTdz(T=T, P=P), "RH": C.calc_RH(T=T, vapor=g["o_qv_to_vapor"]),
    }
    for k, v in pairs.items():
        assert_close(v, g[f"o_{k}"], 1e-10, k)
    assert_same(C.calc_theta_v(theta=g["o_theta"], qv=qv, ql=ql), g["o_theta_v"], "theta_v (pure arithmetic)")
    # dispatch through optional arguments, scalars and errors
    from meteotools_b200.exceptions import InputError
    assert_close(C.calc_Td(P=P, qv=qv), g["o_Td"], 1e-10, "Td from (P,qv)")
    assert_close(C.calc_theta_e(T=T, P=P, qv=qv), g["o_theta_e"], 1e-10, "theta_e from (T,P,qv)")
    assert abs(C.calc_theta(300.0, 90000.0) - co.theta(300.0, 90000.0)) < 1e-10
    assert C.calc_dTdz() == -9.8 / 1004.
    with pytest.raises(InputError):
        C.calc_Td()
    with pytest.raises(InputError):
        C.calc_Tv(T)


def test_tc_iteration_semantics(C, golden_calc):
    g = golden_calc
    T, P, qv = g["t_T"], g["t_P"], g["t_qv"]
    ref, n_ref = co.Tc(T, P, qv, return_iters=True)
    got, n = C.calc_Tc(T, P, qv=qv, return_iterations=True)
    assert n == n_ref and 1 < n < 10
    assert_close(got, ref, 1e-10, "Tc")
    Tn = T.copy()
    Tn[0, 0, 0] = np.nan       # NaN anywhere: np.max is NaN, all 10 sweeps run
    got, n = C.calc_Tc(Tn, P, qv=qv, return_iterations=True)
    assert n == 10
    assert_close(got, g["o_Tc_nan"], 1e-10, "Tc with NaN")


def test_tc_nan_path_wide_range(C):
    """NaN present => ten certain sweeps, done as: fixed point L* by Newton (two float32 steps, one
    double step) + the float32 recursion of the deviation (csrc/thermo.cu k_tc_finish).  Checked
    against the literal ten sweeps of calc/thermaldynamic.py:392-402 over a far wider range than
    any atmosphere, and on points that must take the literal fallback."""
    rng = np.random.default_rng(11)
    n = 400_000
    T = rng.uniform(180, 330, n)
    P = np.exp(rng.uniform(np.log(1e3), np.log(1.08e5), n))
    qv = np.exp(rng.uniform(np.log(1e-8), np.log(0.05), n))
    T[5] = np.nan
    # pathological points: almost no contraction / log of a non-positive number in some sweep
    T[10:14] = [300.0, 250.0, 1e-3, 5e4]
    P[10:14] = [1e5, 1e2, 1e5, 1e5]
    qv[10:14] = [5.0, 1e3, 1e-3, 1e-30]
    with np.errstate(all="ignore"):
        ref, n_ref = co.Tc(T, P, qv, return_iters=True)
    got, n_it = C.calc_Tc(T, P, qv=qv, return_iterations=True)
    assert n_it == n_ref == 10
    assert_close(got, ref, 1e-10, "Tc wide range")
    fin = np.isfinite(ref)
    rel = np.abs(np.asarray(got)[fin] - ref[fin]) / np.abs(ref[fin])
    assert rel.max() < 5e-12, rel.max()
