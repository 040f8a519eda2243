"""Profiling driver for the secondary rows: FD2 along the three axes of (80, 1000, 1000), S3 cyl_d in the grid's
layout, car_d on (80, 1000, 1000), mt_interplevel with 12 variables -- two launches each, nothing else.
   ncu --set full ... -k regex:'k_fd|k_stencil_tma|k_stage_a|k_stage_b|k_interplevel_tile' python profiles/prof_rows.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_rows as br

br.bench_fd((80, 1000, 1000), 1)
for kind, dims, label in (("cyl_trz", (60, 360, 300), "S3 cyl_d"), ("car", (80, 1000, 1000), "car_d (80,1000,1000)")):
    case = br.DynCase(kind, *dims)
    for _ in range(2):
        case.run()
    torch.cuda.synchronize()
    del case
    torch.cuda.empty_cache()
br.bench_interplevel(1, V=12)
torch.cuda.synchronize()
print("ok")
