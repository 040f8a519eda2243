import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import calc_oracle as co
from meteotools_b200.grids import cylindrical_grid
rng = np.random.default_rng(0)
nz, ny, nx, dx = 20, 96, 96, 2000.0
yy, xx = np.meshgrid(np.arange(ny) * dx, np.arange(nx) * dx, indexing="ij")
hgt = 1200.0 * np.exp(-((xx - 60e3) ** 2 + (yy - 110e3) ** 2) / 30e3 ** 2)
k = np.arange(nz).reshape(-1, 1, 1)
z3d = hgt[None] + (15000.0 - hgt[None]) * (k / (nz - 1)) ** 1.5
src = {}
for name in ("u", "v", "w"):
    src[name] = 10 * rng.standard_normal((nz, ny, nx))
src["p"] = 1e5 * np.exp(-z3d / 8000.0) + rng.standard_normal((nz, ny, nx))
src["T"] = 300 - 6.5e-3 * z3d + rng.standard_normal((nz, ny, nx))
src["qv"] = 1e-3 + 1e-3 * np.abs(rng.standard_normal((nz, ny, nx)))
src["qc"] = 1e-4 * np.abs(rng.standard_normal((nz, ny, nx)))
src["qr"] = 1e-4 * np.abs(rng.standard_normal((nz, ny, nx)))
f2d = 5e-5 + 1e-7 * rng.standard_normal((ny, nx))
settings = dict(Nr=40, Ntheta=48, Nz=16, dr=2000, dtheta=2 * np.pi / 48, dz=500, z_start=100,
                r_start=0, theta_start=0, vertical_coord_type="z", dt=300)
cyl = cylindrical_grid(settings)
cx, cy = ((nx + 1) / 2 - 1) * dx, ((ny + 1) / 2 - 1) * dx
cyl.set_horizontal_location(cx, cy)
cyl.set_data.__wrapped__(cyl, dx, dx, z3d, f=f2d, hgt=hgt, **src)
cyl.calc_all_thermaldynamic()
cyl.calc_diagnostics()
hgt_c = oracle.interp2D_fast(dx, dx, cyl.xTR, cyl.yTR, hgt)
idx = co.cyl_terrain_index(hgt_c, 100.0, 500.0)
ref = {}
for name, a in src.items():
    lev = oracle.interplevel(a, z3d, cyl.z)
    ref[name] = co.cyl_terrain_mask(oracle.interp2D_fast_layers(dx, dx, cyl.xTR, cyl.yTR, lev), idx)
ref["f"] = oracle.interp2D_fast(dx, dx, cyl.xTR, cyl.yTR, f2d)
with np.errstate(all="ignore"):
    dyn = co.cyl_d_set(ref, dict(r=cyl.r, theta=cyl.theta, dr=cyl.dr, dtheta=cyl.dtheta, dz=cyl.dz))
for name in dyn:
    a, b = np.asarray(getattr(cyl, name)), dyn[name]
    if a.dtype == bool:
        print(name, 'bool mismatches', np.count_nonzero(a != b)); continue
    nan_diff = np.isnan(a) != np.isnan(b)
    inf_diff = np.isinf(a) != np.isinf(b)
    fin = np.isfinite(a) & np.isfinite(b)
    err = np.abs(a[fin]-b[fin])/np.maximum(np.abs(b[fin]), 1e-6*np.max(np.abs(b[fin])))
    print(f"{name:20s} nan_diff={nan_diff.sum():6d} inf_diff={inf_diff.sum():6d} maxrel={err.max():.3g} exact={np.array_equal(a[fin],b[fin])}")
    if nan_diff.any():
        ii = np.argwhere(nan_diff)[:5]
        for i in ii:
            i = tuple(i); print('   at', i, 'gpu', a[i], 'ref', b[i])
