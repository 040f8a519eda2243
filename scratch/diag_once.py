"""Run the full cyl_d / car_d set a few times (for ncu captures): python scratch/diag_once.py kind nz nt nr [n]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scratch"))
import torch
from meteotools_b200 import _device as D, _lib
import diag_exp as X

kind = sys.argv[1]
nz, nt, nr = (int(x) for x in sys.argv[2:5])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 2
f, f2d, g, din, dout, outs, call, keep, n_out = X.setup(kind, nz, nt, nr)
for _ in range(n):
    _lib.check(call(g, din, dout, _lib.MT_STAGE_ALL, D.stream()), "dyn")
torch.cuda.synchronize()
print("ok")
