import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


@pytest.fixture(scope="session")
def golden_interp():
    return load_golden("interp")


@pytest.fixture(scope="session")
def golden_calc():
    return load_golden("calc")


@pytest.fixture(scope="session")
def golden_grids():
    return load_golden("grids")


@pytest.fixture(scope="session")
def golden_setdata():
    return load_golden("setdata")


@pytest.fixture(scope="session")
def golden_reductions():
    return load_golden("reductions")


def assert_same(a, b, what=""):
    """Bit-exact comparison treating NaN==NaN and checking inf signs."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype == bool or b.dtype == bool:
        assert np.array_equal(a, b), what
        return
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    if not ok.all():
        idx = np.argwhere(~ok)[0]
        raise AssertionError(f"{what}: {np.count_nonzero(~ok)} mismatches, first at {tuple(idx)}: "
                             f"{a[tuple(idx)]!r} vs {b[tuple(idx)]!r}")


def assert_close(a, b, rtol=1e-10, what="", nonfinite="exact", floor=1e-6):
    """north_star tolerance: |a-b| <= rtol*max(|b|, scale) at finite points, same finite mask.
    ``scale`` = the field's own magnitude (max finite |b|) times ``floor`` (default 1e-6), i.e.
    values more than six decades below the field's peak are compared absolutely: FD2 of a field
    that carries a 1-ulp difference cannot be relatively accurate where the derivative crosses
    zero.  PV (a sum of three products of derivatives that cancel to 1e-4 of their size) is
    compared with floor=1e-3 AND, separately, bit for bit against the oracle fed with the GPU's
    own Th (test_gpu_fullsize.py): the stencil arithmetic is exact, only pow's ulps differ.

    nonfinite="exact": NaN and +-inf must sit at identical points with identical signs.
    nonfinite="class": only finite-vs-non-finite must agree.  Used for exp/log/pow-derived fields
    in the r = 0 column of the cylindrical grid, where the reference divides rounding noise such
    as 3v-4v+v (identical v) by r[0] = 0: the sign of that noise, hence inf vs -inf vs NaN, flips
    with a 1-ulp change of the input (DESIGN.md "singular column")."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), f"{what}: finite masks differ"
    if nonfinite == "exact":
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN masks differ"
        inf = np.isinf(b)
        assert np.array_equal(a[inf], b[inf]), f"{what}: inf signs differ"
    if not fin.any():
        return
    scale = floor * np.max(np.abs(b[fin]))
    err = np.abs(a[fin] - b[fin])
    tol = rtol * np.maximum(np.abs(b[fin]), scale)
    if not (err <= tol).all():
        i = np.argmax(err / tol)
        raise AssertionError(f"{what}: max err/tol = {(err / tol)[i]:.3g} (a={a[fin][i]!r}, b={b[fin][i]!r})")
