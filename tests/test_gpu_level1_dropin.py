"""Level 1 of the drop-in: the UNMODIFIED reference Python running on the
GPU because meteotools_b200/lib/fastcompute.so is installed as its ``meteotools/lib/fastcompute.so``.

The reference copy is baseline/_ref/meteotools (made by __graft_entry__.build() in the build container; it
travels to the GPU box).  The run happens in a subprocess: this process may already hold the reference
package bound to the CPU oracle library.  Every wrapper of meteotools.interpolation and
calc.cartesian2cylindrical is compared BIT FOR BIT with tests/golden/interp.npz, which the same reference
code produced on the CPU oracle library (tests/golden/make_golden.py)."""
import json
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "meteotools")
GPU_SO = os.path.join(ROOT, "meteotools_b200", "lib", "fastcompute.so")

SCRIPT = textwrap.dedent("""
    import json, sys, types
    import numpy as np
    tmp, golden = sys.argv[1], sys.argv[2]
    for name in ("wrf", "netCDF4"):          # absent third-party modules the grids import (not used here)
        m = types.ModuleType(name)
        m.omp_set_num_threads = lambda n: None
        m.interplevel = None
        m.Dataset = object
        sys.modules[name] = m
    sys.path.insert(0, tmp)
    import meteotools.interpolation as I
    import meteotools.calc as C
    g = np.load(golden)
    dx, dy, dz = float(g["i_dx"]), float(g["i_dy"]), float(g["i_dz"])
    xn, yn, zn, xo = g["i_xn"], g["i_yn"], g["i_zn"], g["i_xo"]
    got = {
        "interp1D_fast": I.interp1D_fast(dx, xn, g["i_a1"]),
        "interp1D_fast_layers": I.interp1D_fast_layers(dx, xn, g["i_a1l"]),
        "interp1D_nonequal": I.interp1D_nonequal(g["i_xin"], xo, g["i_a1"]),
        "interp1D_nonequal_dec": I.interp1D_nonequal(g["i_xin_dec"], xo, g["i_a1"]),
        "interp1D_nonequal_layers": I.interp1D_nonequal_layers(g["i_xin"], xo, g["i_a1l"]),
        "interp2D_fast": I.interp2D_fast(dx, dy, xn, yn, g["i_a2"]),
        "interp2D_fast_layers": I.interp2D_fast_layers(dx, dy, xn, yn, g["i_a2l"]),
        "interp3D_fast": I.interp3D_fast(dx, dy, dz, xn, yn, zn, g["i_a3"]),
        "interp3D_fast_layers": I.interp3D_fast_layers(dx, dy, dz, xn, yn, zn, g["i_a3l"]),
    }
    r = np.arange(0, 100 + 0.0001, 1)
    theta = np.linspace(0, np.pi * 2, 361)[:-1]
    data = np.add.outer(np.arange(3.0), np.add.outer(np.arange(251.0), np.arange(251.0)))
    got["c2c"] = C.cartesian2cylindrical(1, 1, data, [125, 125], r, theta)
    res = {}
    for k, a in got.items():
        b = g["o_" + k]
        a = np.asarray(a)
        res[k] = bool(a.shape == b.shape and np.array_equal(a, b, equal_nan=True))
    import ctypes
    lib = ctypes.CDLL(tmp + "/meteotools/lib/fastcompute.so")
    lib.mt_launch_count.restype = ctypes.c_int64
    res["_gpu_launches"] = int(lib.mt_launch_count())
    print("RESULT " + json.dumps(res))
""")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "__init__.py")),
                    reason="baseline/_ref/meteotools (copy of the reference made by build()) is absent")
def test_unmodified_reference_python_on_the_gpu_library(tmp_path):
    dst = tmp_path / "meteotools"
    shutil.copytree(REF, dst, ignore=shutil.ignore_patterns("__pycache__", "lib"))
    os.makedirs(dst / "lib")
    shutil.copy(GPU_SO, dst / "lib" / "fastcompute.so")      # what `python build.py` of the drop-in produces
    script = tmp_path / "run_ref.py"
    script.write_text(SCRIPT)
    golden = os.path.join(ROOT, "tests", "golden", "interp.npz")
    env = dict(os.environ, OMP_NUM_THREADS="4")
    p = subprocess.run([sys.executable, str(script), str(tmp_path), golden], capture_output=True, text=True,
                       timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    launches = res.pop("_gpu_launches")
    assert launches > 0, "the reference ran, but not on the CUDA library"
    bad = [k for k, ok in res.items() if not ok]
    assert not bad, f"reference wrappers on the GPU library differ from the golden vectors: {bad}"
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):   # evidence for profiles/ (copied by hand after a gpurun call)
        with open(os.path.join(out, "level1_dropin.json"), "w") as fh:
            json.dump(dict(identical=res, gpu_launches=launches), fh, indent=1)
