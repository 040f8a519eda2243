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


# ---------------------------------------------------------------------------------------------------------
# Parity report: what the tolerance classes above actually measured, per output field, at
# the full sizes -- WITHOUT the floor.  The full-size GPU tests call parity_record(); at the end of the session
# the records are written to gpurun_out/parity_r02.json (copied to profiles/ after a gpurun call).
_PARITY = {}


def parity_record(case, name, a, b, floor=1e-6):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype == bool or b.dtype == bool:
        _PARITY.setdefault(case, {})[name] = dict(points=int(a.size), bit_mismatches=int(np.count_nonzero(a != b)))
        return
    a, b = a.astype(np.float64), b.astype(np.float64)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    fin = np.isfinite(b) & np.isfinite(a)
    rec = dict(points=int(a.size), bit_mismatches=int(np.count_nonzero(~same)),
               nan_masks_equal=bool(np.array_equal(np.isnan(a), np.isnan(b))),
               inf_masks_and_signs_equal=bool(np.array_equal(np.isinf(a), np.isinf(b)) and
                                              np.array_equal(a[np.isinf(b) & np.isinf(a)], b[np.isinf(b) & np.isinf(a)])),
               finite_masks_equal=bool(np.array_equal(np.isfinite(a), np.isfinite(b))),
               nan_points=int(np.isnan(b).sum()), inf_points=int(np.isinf(b).sum()))
    if fin.any():
        bb, err = np.abs(b[fin]), np.abs(a[fin] - b[fin])
        nz = bb > 0
        scale = floor * bb.max()
        rel = np.zeros_like(err)
        rel[nz] = err[nz] / bb[nz]
        above = bb >= scale
        rec.update(max_abs_err=float(err.max()), max_abs_value=float(bb.max()),
                   max_rel_err_no_floor=float(rel.max()) if nz.any() else 0.0,
                   max_rel_err_above_floor=float(rel[above].max()) if above.any() else 0.0,
                   floor=floor, points_below_floor=int(np.count_nonzero(~above)),
                   max_abs_err_below_floor_over_scale=float((err[~above] / scale).max()) if (~above).any() else 0.0,
                   points_rel_err_gt_1e10=int(np.count_nonzero(rel > 1e-10)))
    _PARITY.setdefault(case, {})[name] = rec


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_r02.json"), "w") as fh:
            json.dump(dict(note="max_rel_err_no_floor: max |gpu - oracle| / |oracle| over finite non-zero oracle values, "
                                "NO floor; bit_mismatches counts points that are not bit-identical (NaN == NaN)",
                           cases=_PARITY), fh, indent=1)
    except OSError:
        pass
