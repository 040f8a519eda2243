"""A/B of the stencil stage of mt_cyl_d / mt_car_d: TMA plane pipeline (TR = 16 / 8) vs the register-marching
kernels, bit-for-bit comparison of every output + per-kernel CUDA-event times.
    python scratch/diag_exp.py [car|cyl_trz|cyl_ztr] [nz nt nr]"""
import ctypes as ct
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meteotools_b200 import _device as D, _lib  # noqa: E402

if os.environ.get('MT_LIB'):
    _lib.LIB_PATH = os.environ['MT_LIB']   # A/B builds (scratch/build_variant.sh)
lib = _lib.load()
PEAK = 6536.0


def make(kind, nz, nt, nr, seed=1):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    trz = kind == "cyl_trz"
    shape = (nt, nr, nz) if trz else (nz, nt, nr)
    zz = torch.arange(nz, device="cuda", dtype=torch.float64) * 250.0 + 100.0
    zz = zz.reshape(1, 1, -1) if trz else zz.reshape(-1, 1, 1)
    rn = lambda: torch.randn(shape, generator=gen, device="cuda", dtype=torch.float64)  # noqa: E731
    f = {}
    for k in ("u", "v", "w"):
        f[k] = 10 * rn()
    f["p"] = 1e5 * torch.exp(-zz / 8000.0) + rn()
    f["T"] = 300 - 6.5e-3 * zz + rn()
    f["qv"] = 1e-3 + 1e-3 * rn().abs()
    for k in ("qc", "qr", "qi", "qs", "qg"):
        f[k] = 1e-4 * rn().abs()
    return f


def setup(kind, nz, nt, nr):
    cyl = kind.startswith("cyl")
    f = make(kind, nz, nt, nr)
    f2d = torch.full((nt, nr), 5e-5, device="cuda", dtype=torch.float64) + \
        1e-9 * torch.arange(nt * nr, device="cuda", dtype=torch.float64).reshape(nt, nr)
    g = _lib.mt_grid_t()
    g.nz, g.nt, g.nr = nz, nt, nr
    g.layout = _lib.MT_LAYOUT_TRZ if kind == "cyl_trz" else _lib.MT_LAYOUT_ZTR
    g.dz, g.dt, g.dr = (250.0, 2 * np.pi / nt, 1000.0) if cyl else (250.0, 2000.0, 2000.0)
    keep = []
    if cyl:
        th = np.arange(nt) * (2 * np.pi / nt)
        r = torch.arange(nr, device="cuda", dtype=torch.float64) * 1000.0
        st, ctb = torch.from_numpy(np.sin(th)).cuda(), torch.from_numpy(np.cos(th)).cuda()
        g.r, g.sin_t, g.cos_t = r.data_ptr(), st.data_ptr(), ctb.data_ptr()
        keep += [r, st, ctb]
    g.halo_lo = g.halo_hi = 0
    g.is_bottom = g.is_top = 1
    din = _lib.mt_dyn_in_t()
    for k in ("u", "v", "w", "p", "T", "qv"):
        setattr(din, k, f[k].data_ptr())
    din.f = f2d.data_ptr()
    for i, k in enumerate(("qc", "qr", "qi", "qs", "qg")):
        din.q[i] = f[k].data_ptr()
    names = _lib.CYL_D_OUT if cyl else _lib.CAR_D_OUT
    dout = (_lib.mt_cyl_d_out_t if cyl else _lib.mt_car_d_out_t)()
    outs = {}
    shape = next(iter(f.values())).shape
    for k in names:
        outs[k] = torch.zeros(shape, dtype=torch.uint8 if k == "strain_region" else torch.float64, device="cuda")
        setattr(dout, k, outs[k].data_ptr())
    call = lib.mt_cyl_d if cyl else lib.mt_car_d
    return f, f2d, g, din, dout, outs, call, keep, len(names)


def run(call, g, din, dout, reps):
    def once():
        if os.environ.get("SPLIT") == "1":   # point-wise stage (wind outputs included) and stencil stage as two calls
            _lib.check(call(g, din, dout, _lib.MT_STAGE_A, D.stream()), "dyn A")
            _lib.check(call(g, din, dout, _lib.MT_STAGE_B, D.stream()), "dyn B")
        else:
            _lib.check(call(g, din, dout, _lib.MT_STAGE_ALL, D.stream()), "dyn")
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    lib.mt_profile_enable(1)
    for _ in range(reps):
        once()
    buf = ct.create_string_buffer(1 << 14)
    _lib.check(lib.mt_profile_fetch(buf, len(buf)))
    lib.mt_profile_enable(0)
    per = {ln.split()[0]: round(float(ln.split()[2]) / int(ln.split()[1]), 4)
           for ln in buf.value.decode().strip().splitlines()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, per


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "car"
    dims = [int(x) for x in sys.argv[2:5]] if len(sys.argv) >= 5 else \
        ([80, 1000, 1000] if kind == "car" else [60, 360, 300])
    reps = int(os.environ.get("REPS", "5"))
    nz, nt, nr = dims
    f, f2d, g, din, dout, outs, call, keep, n_out = setup(kind, nz, nt, nr)
    alg = 8.0 * nz * nt * nr * (11 + n_out)
    res = {}
    ref = None
    T = lambda split=0: {"MT_DIAG_TMA": "1", "SPLIT": str(split)}  # noqa: E731
    variants = (("marching", {"MT_DIAG_TMA": "0", "SPLIT": "0"}), ("tma", T(0)), ("tma_split", T(1)))
    only = os.environ.get("VARIANTS")
    if only:
        variants = tuple(v for v in variants if v[0] in only.split(","))
    for label, env in variants:
        os.environ["MT_DIAG_TRIVIAL"] = "0"
        os.environ.update(env)
        for o in outs.values():
            o.zero_()
        ms, per = run(call, g, din, dout, reps)
        snap = outs
        same = None
        if ref is None:
            ref = {k: v.clone() for k, v in outs.items()}
        else:
            bad = []
            for k in snap:
                a, b = snap[k], ref[k]
                if a.dtype == torch.uint8:
                    eq = torch.equal(a, b)
                else:
                    eq = bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
                if not eq:
                    if a.dtype != torch.uint8:
                        ne = ~((a == b) | (torch.isnan(a) & torch.isnan(b)))
                        idx = ne.nonzero()[:3].tolist()
                        bad.append((k, int(ne.sum()), idx))
                    else:
                        bad.append((k, int((a != b).sum()), []))
            same = bad or True
        res[label] = dict(ms=round(ms, 4), frac=round(alg / ms / 1e6 / PEAK, 3), per_kernel=per, identical=same)
        print(label, json.dumps(res[label]), flush=True)
    print(json.dumps(dict(kind=kind, dims=dims, alg_bytes=alg, results=res)))


if __name__ == "__main__":
    main()
