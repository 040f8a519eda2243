"""Profiling driver: the fused cyl_d (default) or car_d set of profiles/bench_kernels.py, a few
iterations, nothing else.   python profiles/prof_dyn.py [cyl|car] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch

import bench_kernels as bk
from meteotools_b200 import _lib
from meteotools_b200.pipeline import LevelShardDyn

kind = sys.argv[1] if len(sys.argv) > 1 else "cyl"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shape, spacings = ((60, 360, 300), (250.0, 2 * np.pi / 360, 1000.0)) if kind == "cyl" else \
    ((80, 1000, 1000), (250.0, 2000.0, 2000.0))
gen = torch.Generator(device="cuda").manual_seed(1)
names = ["u", "v", "w", "p", "T", "qv", "qc", "qr", "qi", "qs", "qg"]
f = bk.fields(shape, gen, names)
nz, nt, nr = shape
f2d = torch.full((nt, nr), 5e-5, device="cuda", dtype=torch.float64)
tables = None
if kind == "cyl":
    th = np.arange(nt) * spacings[1]
    tables = dict(r=torch.arange(nr, device="cuda", dtype=torch.float64) * spacings[2],
                  sin_t=torch.from_numpy(np.sin(th)).cuda(), cos_t=torch.from_numpy(np.cos(th)).cuda())
outs = [k for k in (_lib.CYL_D_OUT if kind == "cyl" else _lib.CAR_D_OUT)]
job = LevelShardDyn(kind, f, f2d, spacings, (0, nz, 0, 0), nz, outs, tables)
for _ in range(reps):
    job.stage_a()
    job.stage_b()
torch.cuda.synchronize()
print("ok", kind, reps)
