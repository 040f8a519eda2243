"""Stage B of car_d in one pass (all 10 FD2-based outputs) vs two passes over disjoint output groups
(fewer concurrent output streams per pass), (80, 1000, 1000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles"))
import bench_kernels as bk
from meteotools_b200 import _lib
from meteotools_b200.pipeline import LevelShardDyn
shape = (80, 1000, 1000)
gen = torch.Generator(device="cuda").manual_seed(1)
names = ["u", "v", "w", "p", "T", "qv", "qc", "qr", "qi", "qs", "qg"]
f = bk.fields(shape, gen, names)
f2d = torch.full(shape[1:], 5e-5, device="cuda", dtype=torch.float64)
A = ["wind_speed", "kinetic_energy", "Tv", "rho", "density_factor", "Th", "Th_rho"]
G1 = ["zeta_r", "zeta_t", "div", "PV"]
G2 = ["zeta_z", "abs_zeta_z", "div_hori", "shear_deform", "stretch_deform", "filamentation_time"]
res = {}
for label, outs in (("full", A + G1 + G2), ("G1", A + G1), ("G2", A + G2), ("G1 (3)", A + G1[:3]), ("PV only", A + ["PV"])):
    job = LevelShardDyn("car", f, f2d, (250.0, 2000.0, 2000.0), (0, shape[0], 0, 0), shape[0], outs)
    job.stage_a()
    ms = bk.timed(job.stage_b, 10)
    res[label] = ms
    print(f"{label:8s} stage B {ms:7.3f} ms")
    del job
    torch.cuda.empty_cache()
print("two passes:", res["G1"] + res["G2"], "vs one:", res["full"])
