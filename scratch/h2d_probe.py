"""Box-only strided H2D upload (mt_upload_box_pitched = one 2-D DMA per field) of the 13 fields of a step,
float64 vs float32 host arrays, for several alignments of the box's first column; with and without a
concurrent 674 MB D2H on a second stream (what TimeStepStream does).  python scratch/h2d_probe.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meteotools_b200 import _device as D

NZ, NY, NX, NV = 60, 1000, 1000, 13
def pinned(dtype):
    t = torch.empty((NZ, NY, NX), dtype=dtype, pin_memory=True)
    t.zero_()
    return t.numpy()
res = []
d2h_src = torch.empty(13 * 6480000, dtype=torch.float64, device="cuda")
d2h_dst = torch.empty(13 * 6480000, dtype=torch.float64, pin_memory=True)
side = torch.cuda.Stream()
for dtype, name in ((torch.float64, "f64"), (torch.float32, "f32")):
    arrs = [pinned(dtype) for _ in range(NV)]
    for (x0, w) in ((349, 302), (352, 302), (352, 304), (320, 352), (0, 1000)):
        y0, y1, x1 = 349, 651, x0 + w
        outs = [D.upload_box(arrs[0], y0, y1, x0, x1) for _ in range(NV)]
        for conc in (False, True):
            ts = []
            for rep in range(7):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if conc:
                    with torch.cuda.stream(side):
                        d2h_dst.copy_(d2h_src, non_blocking=True)
                e0.record()
                for a, o in zip(arrs, outs):
                    D.upload_box(a, y0, y1, x0, x1, out=o)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            nbytes = NV * NZ * (y1 - y0) * w * arrs[0].itemsize
            r = dict(dtype=name, x0=x0, width=w, concurrent_d2h=conc, ms=round(ms, 3), min_ms=round(min(ts), 3),
                     gbs=round(nbytes / ms / 1e6, 1), mbytes=round(nbytes / 1e6, 1))
            res.append(r)
            print(json.dumps(r), flush=True)
    del arrs
