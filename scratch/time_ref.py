import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
t0=time.perf_counter(); step = bench.make_step(0, pinned=False); print('make_step', time.perf_counter()-t0)
r = np.arange(0, 300*1000., 1000.); th = np.arange(0, 360*(2*np.pi/360), 2*np.pi/360)
z = np.arange(60)*300.+100
rr, tt = np.meshgrid(r, th); cx = 999000.0
xTR, yTR = rr*np.cos(tt)+cx, rr*np.sin(tt)+cx
t, out = bench.reference_step(step, z, xTR, yTR, ['T']); print(t)
print('nan frac', np.isnan(out['T']).mean())
t0=time.perf_counter(); print('td', bench.reference_td(bench.synthetic_cyl_fields()))
