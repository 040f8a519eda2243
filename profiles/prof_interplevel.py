"""Profiling driver: mt_interplevel over the full (60, 1000, 1000) domain, 4 variables, level-major result
(the layout wrf.interplevel returns) -- k_interplevel_tile<double, false, true>.  Used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteotools_b200 import _device as D, _lib
lib = _lib.load()
nz, ny, nx, nlev, V = 60, 1000, 1000, 60, 4
gen = torch.Generator(device="cuda").manual_seed(3)
k = torch.arange(nz, device="cuda", dtype=torch.float64).reshape(-1, 1, 1)
hgt = 1500.0 * torch.rand((ny, nx), generator=gen, device="cuda", dtype=torch.float64)
z3d = hgt[None] + (20000.0 - hgt[None]) * (k / (nz - 1)) ** 1.5
lv = D.to_device(np.arange(nlev) * 300.0 + 100.0)
flds = [torch.randn((nz, ny, nx), generator=gen, device="cuda", dtype=torch.float64) for _ in range(V)]
outs = [torch.empty((nlev, ny, nx), device="cuda", dtype=torch.float64) for _ in range(V)]
for _ in range(3):
    _lib.check(lib.mt_interplevel(D.ptr_array(flds), V, D.ptr(z3d), _lib.MT_F64, nz, ny, nx, 0, ny, 0, nx, D.ptr(lv), nlev, 1,
                                  -1, D.ptr_array(outs), ny * nx, nx, 1, 0, D.stream()))
torch.cuda.synchronize()
print("ok")
