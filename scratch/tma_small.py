import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MT_SETDATA", "tma")
import numpy as np, torch
import oracle
from meteotools_b200.grids import cylindrical_grid
rng = np.random.default_rng(1)
nz, ny, nx, dx = 14, 70, 84, 2000.0
settings = dict(Nr=30, Ntheta=40, Nz=12, dr=1100, dtheta=2 * np.pi / 40, dz=450, z_start=150,
                r_start=0, theta_start=0, vertical_coord_type="z", dt=300)
hgt = 900.0 * rng.random((ny, nx))
k = np.arange(nz).reshape(-1, 1, 1)
z3d = hgt[None] + (7000.0 - hgt[None]) * (k / (nz - 1)) ** 1.3
fields = {f"v{i}": rng.standard_normal((nz, ny, nx)) for i in range(2)}
cyl = cylindrical_grid(settings)
cyl.set_horizontal_location(33 * dx + 517.0, 34 * dx + 1311.0)
cyl.set_data(dx, dx, z3d, **fields)
torch.cuda.synchronize()
for name, a in fields.items():
    ref = oracle.interp2D_fast_layers(dx, dx, cyl.xTR, cyl.yTR, oracle.interplevel(a, z3d, cyl.z))
    got = np.asarray(getattr(cyl, name))
    same = np.array_equal(got, ref, equal_nan=True)
    print(name, "same" if same else "DIFF", np.nanmax(np.abs(got - ref)) if not same else 0, np.isnan(got).sum(), np.isnan(ref).sum())
