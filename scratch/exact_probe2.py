"""Which assert_close comparisons of tests/test_gpu_reductions.py are in fact bit-identical?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest
import test_gpu_reductions as T
rec = []
orig = conftest.assert_close
def probe(a, b, rtol=1e-10, what="", **kw):
    a_, b_ = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    same = a_.shape == b_.shape and bool(((a_ == b_) | (np.isnan(a_) & np.isnan(b_))).all())
    rec.append((what, same))
    return orig(a, b, rtol, what, **kw)
T.assert_close = probe
g = conftest.load_golden("reductions")
for fn in (T.test_calc_averages, T.test_cyl_reductions, T.test_cyl_IKE10_reference_quirks, T.test_car_averages_and_smoothers):
    fn(g)
for what, same in rec:
    print(("EXACT  " if same else "close  ") + what)
