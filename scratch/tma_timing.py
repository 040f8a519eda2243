"""Phase clocks of k_setdata_tma (consumer warps 0, 7, 14; debug library built with -DMT_TMA_TIMING):
run on the GPU box after copying scratch/fastcompute_timing.so over lib/fastcompute.so."""
import ctypes as ct, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from meteotools_b200 import _lib
from meteotools_b200.pipeline import CylStepPipeline
step = bench.make_step(0, pinned=False)
cyl = bench.make_grid()
dev3d = {k: torch.from_numpy(v).cuda() for k, v in step["vars3d"].items()}
dev2d = {"f": torch.from_numpy(step["f"]).cuda(), "hgt": torch.from_numpy(step["hgt"]).cuda()}
devz = torch.from_numpy(step["z3d"]).cuda()
pipe = CylStepPipeline(cyl, (bench.NZ, bench.NY, bench.NX), bench.DX, bench.DX, bench.NAMES3D, names2d=("f",),
                       hgt=True, resident=True, td_outputs=None)
pipe.bind_device(devz, dev3d, dev2d, direction=0)
lib = ct.CDLL(_lib.LIB_PATH)
for _ in range(3):
    pipe.run()
buf = (ct.c_ulonglong * 8)()
lib.mt_setdata_tma_timing(buf, 1)
N = 10
for _ in range(N):
    pipe.run()
lib.mt_setdata_tma_timing(buf, 1)
names = ["wait coordinate tile", "monotone + brackets", "wait variable tile", "B", "barrier after B", "C"]
tot = sum(buf[i] for i in range(6))
nit = pipe.f_nitems
print("items", nit, "per SM", nit / 148)
for i, nm in enumerate(names):
    per_launch_warp = buf[i] / N / 3 / 148       # cycles per launch, per sampled warp, per CTA
    print(f"{nm:24s} {100 * buf[i] / tot:5.1f} %   {per_launch_warp / 1.965e3:8.1f} us per CTA and launch")
