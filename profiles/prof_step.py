"""Profiling driver: the device-resident S2+cyl_td step of bench.py, a few iterations, nothing
else (no e2e, no CPU baseline).  Used under ncu:  see profiles/README.md for the command lines."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from meteotools_b200.pipeline import CylStepPipeline

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dyn = len(sys.argv) > 2 and sys.argv[2] == "dyn"
step = bench.make_step(0, pinned=False)
cyl = bench.make_grid()
dev3d = {k: torch.from_numpy(v).cuda() for k, v in step["vars3d"].items()}
dev2d = {"f": torch.from_numpy(step["f"]).cuda(), "hgt": torch.from_numpy(step["hgt"]).cuda()}
devz = torch.from_numpy(step["z3d"]).cuda()
from meteotools_b200 import _lib
pipe = CylStepPipeline(cyl, (bench.NZ, bench.NY, bench.NX), bench.DX, bench.DX, bench.NAMES3D, names2d=("f",),
                       hgt=True, resident=True,
                       dyn_outputs=[k for k in _lib.CYL_D_OUT if k != "Th"] if dyn else None)
pipe.bind_device(devz, dev3d, dev2d, direction=0)
for _ in range(steps):
    pipe.run()
torch.cuda.synchronize()
print("ok", steps)
