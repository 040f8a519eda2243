"""SURVEY 8(f)-3 timing: cylindrical_grid.set_data(ver_interp_order=2) on the S2 workload
(12 3-D variables, 1000x1000x60 float64 source -> 60z x 360theta x 300r) on the GPU, next to the
reference's per-column scipy splrep/splev loop timed on a sample of columns on the host."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import oracle

step = bench.make_step(0, pinned=False)
cyl = bench.make_grid()
dev3d = {k: torch.from_numpy(v).cuda() for k, v in step["vars3d"].items()}
devz = torch.from_numpy(step["z3d"]).cuda()
res = {}
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cyl.set_data(bench.DX, bench.DX, devz, ver_interp_order=2, hgt=step["hgt"], **dev3d)
    torch.cuda.synchronize()
    res[f"gpu_s_run{it}"] = time.perf_counter() - t0
y0, y1, x0, x1 = cyl._target_box(bench.NX, bench.NY, bench.DX, bench.DX)
ncol = (y1 - y0) * (x1 - x0)
# CPU: the reference's loop on 30 x 30 columns of one variable
sub = (slice(None), slice(y0, y0 + 30), slice(x0, x0 + 30))
name = bench.NAMES3D[0]
t0 = time.perf_counter()
ref = oracle.spline_levels(step["vars3d"][name][sub], step["z3d"][sub], cyl.z, 2)
cpu = time.perf_counter() - t0
res.update(box_columns=int(ncol), cpu_s_900_columns_1var=cpu,
           cpu_s_est_full_domain_12var=cpu / 900 * bench.NX * bench.NY * 12,
           cpu_s_est_box_12var=cpu / 900 * ncol * 12)
print(json.dumps(res))
