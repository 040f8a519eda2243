"""S3 cyl_d (both layouts), car_d on (80,1000,1000) and S5: python scratch/dyn_quick.py [s5]   [MT_LIB=variant.so]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meteotools_b200 import _lib
if os.environ.get("MT_LIB"):
    _lib.LIB_PATH = os.environ["MT_LIB"]
import bench_rows as br
cases = [("cyl_trz", (60, 360, 300), "S3 TRZ"), ("cyl_ztr", (60, 360, 300), "S3 ZTR"), ("car", (80, 1000, 1000), "car 1000^2")]
if len(sys.argv) > 1 and sys.argv[1] == "s5":
    cases.append(("car", (80, 2000, 2000), "S5"))
for kind, dims, label in cases:
    r = br.bench_dyn(kind, dims, 15, label)
    print(json.dumps({k: r[k] for k in ("row", "ms", "frac", "per_kernel_ms")}))
