"""cartesian_grid.set_data timing (resident sources): 12 variables, (60,1000,1000) -> (60, 400, 400)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from meteotools_b200.grids import cartesian_grid
from meteotools_b200 import _lib
import ctypes as ct
step = bench.make_step(0, pinned=False)
car = cartesian_grid(dict(Nx=400, Ny=400, Nz=60, dx=1500, dy=1500, dz=300, z_start=100, vertical_coord_type="z",
                          SOR_parameter=1.9, max_iteration_times=10, iteration_tolerance=0.05))
cx = ((bench.NX + 1) / 2 - 1) * bench.DX
car.set_horizontal_location(cx, cx)
dev3d = {k: torch.from_numpy(v).cuda() for k, v in step["vars3d"].items()}
devz = torch.from_numpy(step["z3d"]).cuda()
hgt = torch.from_numpy(step["hgt"]).cuda()
lib = _lib.load()
for it in range(4):
    car.varlist3D, car.varlist2D = [], []
    if it == 3:
        lib.mt_profile_enable(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    car.set_data.__wrapped__(car, bench.DX, bench.DX, devz, hgt=hgt, **dev3d)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("set_data", round(dt * 1e3, 3), "ms")
buf = ct.create_string_buffer(1 << 16)
lib.mt_profile_fetch(buf, len(buf)); lib.mt_profile_enable(0)
agg = {}
for ln in buf.value.decode().strip().splitlines():
    p = ln.split(); agg.setdefault(p[0], [0, 0.0]); agg[p[0]][0] += int(p[1]); agg[p[0]][1] += float(p[2])
for k, v in agg.items(): print(k, v[0], round(v[1], 3), "ms")
npts = 60 * 400 * 400
print("points", npts, "alg GB", 8 * npts * 12 * 2 / 1e9)
