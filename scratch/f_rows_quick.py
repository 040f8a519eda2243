"""SURVEY 8(f) rows at size: the smoother passes on (80, 1000, 1000), the azimuthal statistics and the range sums on
the S3 grid (60, 360, 300) in the grid's (theta, r, z) layout.  python scratch/f_rows_quick.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteotools_b200 import _lib, _device as D
if os.environ.get("MT_LIB"):
    _lib.LIB_PATH = os.environ["MT_LIB"]
import bench_rows as br
lib = _lib.load()
PEAK = 6536.0
def rep(label, fn, nbytes, reps=20):
    ms, best = br.timed(fn, reps)
    print(json.dumps(dict(row=label, ms=round(ms, 4), best_ms=round(best, 4), gbs=round(nbytes / ms / 1e6, 1),
                          frac=round(nbytes / ms / 1e6 / PEAK, 3))), flush=True)
# smoothers: one Jacobi pass = read cur (+ v0 at the edges) + write out
shape = (80, 1000, 1000)
v0 = torch.randn(shape, device="cuda", dtype=torch.float64)
cur, out = v0.clone(), torch.empty_like(v0)
for naxes, axa, axb, cw, den in ((1, 2, 0, 2.0, 4.0), (1, 0, 0, 2.0, 4.0), (2, 1, 2, 4.0, 8.0), (3, 0, 1, 6.0, 12.0)):
    rep(f"mt_smooth_pass naxes={naxes} (axes {axa},{axb}) {shape}",
        lambda: _lib.check(lib.mt_smooth_pass(D.ptr(v0), D.ptr(cur), D.ptr(out), *shape, naxes, axa, axb, cw, den, D.stream())),
        16.0 * v0.numel())
del v0, cur, out
# S3 grid, TRZ layout
nz, nt, nr = 60, 360, 300
var = torch.randn((nt, nr, nz), device="cuda", dtype=torch.float64)
g = _lib.mt_grid_t(); g.nz, g.nt, g.nr = nz, nt, nr; g.layout = _lib.MT_LAYOUT_TRZ
g.dz, g.dt, g.dr = 250.0, 2 * np.pi / nt, 1000.0; g.is_bottom = g.is_top = 1
sym = torch.empty((nz, nr), device="cuda", dtype=torch.float64); axs = torch.empty_like(sym); asy = torch.empty_like(var)
rep("mt_azimuthal_stats sym+asy+axisym (60,360,300) TRZ",
    lambda: _lib.check(lib.mt_azimuthal_stats(D.ptr(var), g, D.ptr(sym), D.ptr(asy), D.ptr(axs), D.stream())), 16.0 * var.numel())
rep("mt_azimuthal_stats sym only", lambda: _lib.check(lib.mt_azimuthal_stats(D.ptr(var), g, D.ptr(sym), None, None, D.stream())), 8.0 * var.numel())
# range sums: reduce z (contiguous, stride 1), r (stride nz), theta (stride nr*nz)
w = torch.ones((400,), device="cuda", dtype=torch.float64)
for label, (na, nb, nk, sa, sb, sk) in (("reduce z (contiguous)", (nt, nr, nz, nr * nz, nz, 1)),
                                       ("reduce r", (nt, nz, nr, nr * nz, 1, nz)),
                                       ("reduce theta", (nr, nz, nt, nz, 1, nr * nz))):
    o = torch.empty((na, nb), device="cuda", dtype=torch.float64)
    for mode in (0, 2):
        rep(f"mt_axis_wsum {label} mode {mode}",
            lambda: _lib.check(lib.mt_axis_wsum(D.ptr(var), na, nb, nk, sa, sb, sk, D.ptr(w), mode, 1.0, D.ptr(o), nb, 1, D.stream())),
            8.0 * var.numel())
