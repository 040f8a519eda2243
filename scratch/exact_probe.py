"""Which reduction outputs are already bit-identical to the reference's golden vectors?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import load_golden
import meteotools_b200.calc as C
from golden.make_golden import CYL
from meteotools_b200.grids import cylindrical_grid

def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())

g = load_golden("grids")
cyl = cylindrical_grid(dict(CYL))
for k, v in g.items():
    if k.startswith("g_cyl_"):
        setattr(cyl, k[len("g_cyl_"):], v.copy())
for m in ("calc_vr_vt", "calc_wind_speed", "find_rmw_and_speed"):
    getattr(cyl, m)()
for k in ("sym_vt", "asy_vt", "vt_max_RMW"):
    print(k, same(getattr(cyl, k), g[f"o_cyl_{k}"]))
r = load_golden("reductions")
a, ax = r["r_a"], r["r_ax"]
print("Zavg axis1", same(C.calc_Zaverage(a, ax, 1, ax[4], ax[30]), r["o_Zavg_1"]))
print("Zavg axis1 full", same(C.calc_Zaverage(a, ax, 1), r["o_Zavg_1_full"]))
print("Ravg axis1", same(C.calc_Raverage(a, ax, 1, ax[2], ax[35]), r["o_Ravg_1"]))
print("Zavg axis0", same(C.calc_Zaverage(np.ascontiguousarray(np.moveaxis(a, 1, 0)), ax, 0, ax[4], ax[30]), r["o_Zavg_0"]))
a2 = np.ascontiguousarray(np.moveaxis(a, 1, 2))
print("Zavg axis2 (contiguous)", same(C.calc_Zaverage(a2, ax, 2, ax[4], ax[30]), r["o_Zavg_2"]))
