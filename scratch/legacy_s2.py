"""The LEGACY symbol interp2D_fast_layers (host pointers, Fortran layout, synchronous) at S2 size: L = 60
levels of a 1000 x 1000 field, N = 108 000 cylindrical targets -- what the UNMODIFIED reference Python calls
(interpolation/interp_2d.py:131-150) when lib/fastcompute.so is this library.  GPU library vs the CPU
oracle library on the same call, bit-for-bit comparison."""
import ctypes as ct
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402
from meteotools_b200 import _lib  # noqa: E402

L, NY, NX, DX = 60, 1000, 1000, 2000.0
rng = np.random.default_rng(0)
src = np.asfortranarray(rng.standard_normal((L, NY, NX)))        # what interp_2d.py:136 hands over
cyl = bench.make_grid()
x = np.ascontiguousarray(np.asarray(cyl.xTR).reshape(-1))
y = np.ascontiguousarray(np.asarray(cyl.yTR).reshape(-1))
N = x.size
p = lambda a: a.ctypes.data_as(ct.POINTER(ct.c_double))  # noqa: E731
out_g = np.empty((L, N), order="F")
out_c = np.empty((L, N), order="F")
glib, clib = _lib.load(), oracle.lib()


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


g_ms = timed(lambda: glib.interp2D_fast_layers(NX, NY, L, DX, DX, N, p(x), p(y), p(src), p(out_g)), 5)
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
c_ms = timed(lambda: clib.interp2D_fast_layers(NX, NY, L, DX, DX, N, p(x), p(y), p(src), p(out_c)), 5)
print(json.dumps(dict(case="legacy interp2D_fast_layers, L=60, 1000x1000, N=108000 (host pointers, synchronous)",
                      gpu_library_ms=g_ms[0], gpu_library_best_ms=g_ms[1], cpu_oracle_library_ms=c_ms[0],
                      cpu_threads=os.cpu_count(), identical=bool(np.array_equal(out_g, out_c)),
                      source_bytes=src.nbytes, output_bytes=out_g.nbytes,
                      note="the GPU library uploads only the window of the source under the targets "
                           "(csrc/legacy.cu target_box)")))
