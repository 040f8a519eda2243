"""Per-launch time of k_setdata_tma on the bench step (S2, 12 variables): the quick A/B harness of the
kernel experiments.  python scratch/tma_quick.py [reps]"""
import ctypes as ct, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from meteotools_b200 import _lib
if os.environ.get('MT_LIB'):
    _lib.LIB_PATH = os.environ['MT_LIB']   # A/B builds (scratch/build_variant.sh)
from meteotools_b200.pipeline import CylStepPipeline
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
step = bench.make_step(0, pinned=False)
cyl = bench.make_grid()
dev3d = {k: torch.from_numpy(v).cuda() for k, v in step["vars3d"].items()}
dev2d = {"f": torch.from_numpy(step["f"]).cuda(), "hgt": torch.from_numpy(step["hgt"]).cuda()}
devz = torch.from_numpy(step["z3d"]).cuda()
pipe = CylStepPipeline(cyl, (bench.NZ, bench.NY, bench.NX), bench.DX, bench.DX, bench.NAMES3D, names2d=("f",),
                       hgt=True, resident=True)
pipe.bind_device(devz, dev3d, dev2d, direction=0)
lib = _lib.load()
for _ in range(5):
    pipe.run()
torch.cuda.synchronize()
res = []
for trial in range(3):
    lib.mt_profile_enable(1)
    for _ in range(reps):
        pipe.run()
    buf = ct.create_string_buffer(1 << 16)
    _lib.check(lib.mt_profile_fetch(buf, len(buf)))
    lib.mt_profile_enable(0)
    d = {}
    for line in buf.value.decode().strip().splitlines():
        name, cnt, tot = line.split()
        d[name] = float(tot) / int(cnt)
    res.append(d)
for k in ("k_setdata_tma", "k_td_main"):
    print(k, " ".join(f"{r[k]:.4f}" for r in res if k in r), "ms")
# checksum of the outputs (bit-level): identical across bit-exact variants
import hashlib
h = hashlib.sha256()
for k in sorted(pipe.out3d):
    h.update(pipe.out3d[k].cpu().numpy().tobytes())
print("sha256(out3d)", h.hexdigest()[:16])
