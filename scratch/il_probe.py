"""mt_interplevel over the full (60,1000,1000) domain in the layouts it serves: level-fastest box rows
(k_interplevel_tile) and the reference's level-major (nlev, ny, nx) result (k_interplevel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteotools_b200 import _device as D, _lib
lib = _lib.load()
nz, ny, nx, nlev = 60, 1000, 1000, 60
gen = torch.Generator(device="cuda").manual_seed(3)
k = torch.arange(nz, device="cuda", dtype=torch.float64).reshape(-1, 1, 1)
hgt = 1500.0 * torch.rand((ny, nx), generator=gen, device="cuda", dtype=torch.float64)
z3d = hgt[None] + (20000.0 - hgt[None]) * (k / (nz - 1)) ** 1.5
lev = np.arange(nlev) * 300.0 + 100.0
lv = D.to_device(lev)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for V in (1, 4, 12):
    flds = [torch.randn((nz, ny, nx), generator=gen, device="cuda", dtype=torch.float64) for _ in range(V)]
    alg = 8.0 * ny * nx * (nz * (V + 1) + nlev * V)
    for name, shape, strides in (("level-fastest rows", (ny, nx, nlev), (1, nx * nlev, nlev)),
                                 ("level-major (nlev, ny, nx)", (nlev, ny, nx), (ny * nx, nx, 1))):
        outs = [torch.empty(shape, device="cuda", dtype=torch.float64) for _ in range(V)]
        call = lambda: _lib.check(lib.mt_interplevel(D.ptr_array(flds), V, D.ptr(z3d), _lib.MT_F64, nz, ny, nx, 0, ny, 0, nx,
                                                      D.ptr(lv), nlev, 1, -1, D.ptr_array(outs), *strides, 0, D.stream()))
        ms = timed(call)
        print(f"V={V:2d} {name:28s} {ms:8.3f} ms  {alg / ms / 1e6:8.1f} GB/s  {alg / ms / 1e6 / 6536:.3f} of peak")
        del outs
    del flds
