"""FD2 along the three axes of (80, 1000, 1000) and (80, 1000, 1024): python scratch/fd_quick.py  [MT_LIB=variant.so]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteotools_b200 import _lib
if os.environ.get("MT_LIB"):
    _lib.LIB_PATH = os.environ["MT_LIB"]
import bench_rows
for shape in ((80, 1000, 1000), (80, 1000, 1024)):
    for r in bench_rows.bench_fd(shape, 20):
        print(json.dumps({k: r[k] for k in ("row", "ms", "best_ms", "frac")}))
