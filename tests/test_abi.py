"""CPU checks: the C-ABI library builds, loads and exports every symbol that
include/meteotools_b200.h declares; host-side argument validation raises the reference's
exceptions before any device work; the product never imports the oracle."""
import ctypes as ct
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from meteotools_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "meteotools_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^(?:void|int|int64_t|const char \*)\s*\*?\s*(\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_symbols_exported(built):
    names = declared_symbols()
    assert len(names) >= 35 and "interp2D_fast_layers" in names and "mt_setdata" in names
    lib = ct.CDLL(built)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"not exported: {missing}"
    from meteotools_b200 import _lib
    assert set(_lib.EXPORTED_SYMBOLS) == set(names), set(_lib.EXPORTED_SYMBOLS) ^ set(names)
    L = _lib.load()
    assert L.mt_abi_version() == 2 and L.mt_launch_count() == 0


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_validation_precedes_device_work(built):
    """ValueError / DimensionError contract of interpolation/check_arrays.py, reachable without a GPU."""
    import meteotools_b200.interpolation as I
    from meteotools_b200.exceptions import DimensionError
    a = np.arange(25.0).reshape(5, 5)
    with pytest.raises(ValueError):
        I.interp2D_fast(1, 1, [1.7, 7.5, 3.5], [1.2, 2.3, 3.4], a)
    with pytest.raises(ValueError):
        I.interp2D_fast(1, 1, [1.7, -0.5, 3.5], [1.2, 2.3, 3.4], a)
    with pytest.raises(DimensionError):
        I.interp2D_fast(1, 1, [1.7, 2.5], [1.2, 2.3, 3.4], a)
    with pytest.raises(DimensionError):
        I.interp2D_fast(1, 1, [1.7], [1.2], np.stack([a] * 2))
    with pytest.raises(ValueError):
        I.interp1D_nonequal([0, 2, 1, 5], [1.0], np.arange(4.0))
    with pytest.raises(ValueError):
        I.interp1D_nonequal([0, 2, 4, 5], [7.0], np.arange(4.0))


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import meteotools_b200.calc as C
    from meteotools_b200.exceptions import BackendError
    with pytest.raises(BackendError):
        C.FD2(np.zeros((4, 4)), 1.0, 0)
    with pytest.raises(BackendError):
        C.calc_theta(np.full(3, 300.0), np.full(3, 9e4))


def test_grid_field_store_protocol(built):
    from meteotools_b200.grids import cartesian_grid, cylindrical_grid
    from golden.make_golden import CAR, CYL
    cyl = cylindrical_grid(dict(CYL))
    assert cyl.r.shape == (9,) and cyl.theta.shape == (8,) and cyl.z[0] == 100.0
    cyl.u = np.ones((6, 8, 9))
    cyl.f = np.ones((8, 9))
    assert 'u' in dir(cyl) and 'f' in dir(cyl) and not cyl.u.flags.writeable
    del cyl.u
    assert 'u' not in dir(cyl)
    with pytest.raises(AttributeError):
        cyl.u
    with pytest.raises(AttributeError):
        cyl.calc_wind_speed()
    with pytest.raises(AttributeError):
        cyl.calc_theta()
    cyl.set_horizontal_location(1e5, 1e5)
    assert cyl.xTR.shape == (8, 9)
    car = cartesian_grid(dict(CAR))
    assert car.x.shape == (9,) and car.center_xloc == 8000.0
    with pytest.raises(AttributeError):
        car.calc_vorticity_3D()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "meteotools_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert not re.search(r"#include\s*[<\"][^>\"]*oracle", text), fn
                assert "liboracle" not in text and "_build/" not in text, fn
