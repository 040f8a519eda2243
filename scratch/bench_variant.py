"""bench.py with the library replaced by an A/B build: MT_LIB=scratch/variants/fastcompute_x.so python scratch/bench_variant.py [bench args]"""
import os, sys, runpy
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from meteotools_b200 import _lib
if os.environ.get("MT_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MT_LIB"])
sys.argv = [os.path.join(root, "bench.py")] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
