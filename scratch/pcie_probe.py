"""What the PCIe link of the box gives for the e2e leg's traffic: box uploads (strided 3-D DMA),
13 result downloads, both at once."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from meteotools_b200 import _device as D

step = bench.make_step(0, pinned=True)
cyl = bench.make_grid()
y0, y1, x0, x1 = cyl._target_box(bench.NX, bench.NY, bench.DX, bench.DX)
pitch = D.box_pitch(x1 - x0)
shp = (bench.NZ, y1 - y0, pitch)
dev = [torch.empty(shp, dtype=torch.float64, device="cuda") for _ in range(13)]
srcs = [step["z3d"]] + [step["vars3d"][k] for k in bench.NAMES3D]
outs_d = [torch.empty((360, 300, 60), dtype=torch.float64, device="cuda") for _ in range(13)]
outs_h = [torch.empty((360, 300, 60), dtype=torch.float64, pin_memory=True) for _ in range(13)]
dense_h = [torch.empty((bench.NZ, y1 - y0, x1 - x0), dtype=torch.float64, pin_memory=True) for _ in range(13)]
dense_d = [torch.empty((bench.NZ, y1 - y0, x1 - x0), dtype=torch.float64, device="cuda") for _ in range(13)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def up():
    with torch.cuda.stream(s1):
        for a, d in zip(srcs, dev):
            D.upload_box(a, y0, y1, x0, x1, out=d)
def up_dense():
    with torch.cuda.stream(s1):
        for h, d in zip(dense_h, dense_d):
            d.copy_(h, non_blocking=True)
def down():
    with torch.cuda.stream(s2):
        for h, d in zip(outs_h, outs_d):
            h.copy_(d, non_blocking=True)
def timeit(fns, n=10):
    for f in fns: f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        for f in fns: f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
nb_up = 13 * bench.NZ * (y1 - y0) * (x1 - x0) * 8
nb_dn = 13 * 360 * 300 * 60 * 8
for name, fns, nb in (("up box (strided 3-D DMA)", [up], nb_up), ("up dense (same bytes, contiguous)", [up_dense], nb_up),
                      ("down", [down], nb_dn), ("up box + down", [up, down], nb_up + nb_dn),
                      ("up dense + down", [up_dense, down], nb_up + nb_dn)):
    ms = timeit(fns)
    print(f"{name:36s} {ms:7.2f} ms  {nb / ms / 1e6:6.1f} GB/s")
