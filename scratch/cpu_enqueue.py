"""Host time to enqueue one CylStepPipeline.run() (no synchronisation inside the loop): is the device-resident
bench leg GPU-bound or launch-bound?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from meteotools_b200.pipeline import CylStepPipeline
step = bench.make_step(0, pinned=False)
cyl = bench.make_grid()
dev3d = {k: torch.from_numpy(v).cuda() for k, v in step["vars3d"].items()}
dev2d = {"f": torch.from_numpy(step["f"]).cuda(), "hgt": torch.from_numpy(step["hgt"]).cuda()}
devz = torch.from_numpy(step["z3d"]).cuda()
pipes = [CylStepPipeline(cyl, (bench.NZ, bench.NY, bench.NX), bench.DX, bench.DX, bench.NAMES3D, names2d=("f",),
                         hgt=True, resident=True) for _ in range(2)]
for p in pipes:
    p.bind_device(devz, dev3d, dev2d, direction=0)
streams = [torch.cuda.Stream() for _ in range(2)]
for _ in range(10):
    pipes[0].run()
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for i in range(N):
    with torch.cuda.stream(streams[i % 2]):
        pipes[i % 2].run()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0) / N:.4f} ms per step (host), total {1e3 * (t2 - t0) / N:.4f} ms per step")
