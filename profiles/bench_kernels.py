"""Per-row kernel benchmarks at the SURVEY 8(d) sizes (S1, S3, S5 + generic FD / trilinear), as
evidence for the rows bench.py's headline step does not exercise.  Device-resident inputs, CUDA
events, 3 warm-up + N timed repetitions, algorithmic bytes / time vs the measured HBM peak.

    python profiles/bench_kernels.py [--quick] > profiles/kernels_rNN.json
    torchrun --nproc-per-node N profiles/bench_kernels.py --sharded     (level-sharded car_d, NCCL halo)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from meteotools_b200 import _device as D, _lib  # noqa: E402
from meteotools_b200.pipeline import LevelShardDyn, dyn_level_sharded, level_shard  # noqa: E402

PEAK, PEAK_SRC = bench.peaks()


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def row(name, ms, alg_bytes, points, **extra):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    return dict(kernel=name, ms=ms, alg_bytes=alg_bytes, achieved_gbs=gbs, frac_of_peak=gbs / PEAK,
                points_per_s=points / (ms * 1e-3), **extra)


def fields(shape, gen, names):
    nz = shape[0]
    zz = (torch.arange(nz, device="cuda", dtype=torch.float64) * 250.0 + 100.0).reshape(-1, 1, 1)
    rn = lambda: torch.randn(shape, generator=gen, device="cuda", dtype=torch.float64)  # noqa: E731
    f = {}
    for k in names:
        if k in ("u", "v", "w"):
            f[k] = 10 * rn()
        elif k == "p":
            f[k] = 1e5 * torch.exp(-zz / 8000.0) + rn()
        elif k == "T":
            f[k] = 300 - 6.5e-3 * zz + rn()
        elif k == "qv":
            f[k] = 1e-3 + 1e-3 * rn().abs()
        else:
            f[k] = 1e-4 * rn().abs()
    return f


def bench_dyn(kind, shape, spacings, reps):
    gen = torch.Generator(device="cuda").manual_seed(1)
    names = ["u", "v", "w", "p", "T", "qv", "qc", "qr", "qi", "qs", "qg"]
    f = fields(shape, gen, names)
    nz, nt, nr = shape
    f2d = torch.full((nt, nr), 5e-5, device="cuda", dtype=torch.float64)
    tables = None
    if kind == "cyl":
        th = np.arange(nt) * spacings[1]
        tables = dict(r=torch.arange(nr, device="cuda", dtype=torch.float64) * spacings[2],
                      sin_t=torch.from_numpy(np.sin(th)).cuda(), cos_t=torch.from_numpy(np.cos(th)).cuda())
    outs = [k for k in (_lib.CYL_D_OUT if kind == "cyl" else _lib.CAR_D_OUT)]
    job = LevelShardDyn(kind, f, f2d, spacings, (0, nz, 0, 0), nz, outs, tables)
    lib = _lib.load()
    lib.mt_profile_enable(1)
    for _ in range(reps):
        job.stage_a()
        job.stage_b()
    import ctypes as ct
    buf = ct.create_string_buffer(1 << 14)
    _lib.check(lib.mt_profile_fetch(buf, len(buf)))
    lib.mt_profile_enable(0)
    per = {ln.split()[0]: float(ln.split()[2]) / int(ln.split()[1]) for ln in buf.value.decode().strip().splitlines()}
    ms = timed(lambda: (job.stage_a(), job.stage_b()), reps)
    npts = nz * nt * nr
    n_in, n_out = len(names), len(outs)
    # stage B re-reads the stage-A fields it differentiates (vr vt w Th_rho p rho / u v w Th_rho rho)
    rows = [row(f"{kind}_d fused set (stage A + B), {shape}", ms, 8.0 * npts * (n_in + n_out), npts,
                n_in=n_in, n_out=n_out, per_kernel_ms=per)]
    return rows


def bench_fd(shape, reps):
    v = torch.randn(shape, device="cuda", dtype=torch.float64)
    out = torch.empty_like(v)
    lib = _lib.load()
    rows = []
    n0, n1, n2 = shape
    for axis, (outer, n, inner) in enumerate(((1, n0, n1 * n2), (n0, n1, n2), (n0 * n1, n2, 1))):
        ms = timed(lambda: _lib.check(lib.mt_fd(D.ptr(v), D.ptr(out), outer, n, inner, 250.0, 0, D.stream())), reps)
        rows.append(row(f"FD2 axis {axis}, {shape}", ms, 16.0 * v.numel(), v.numel()))
    return rows


def bench_interplevel(reps):
    """SURVEY 8(d) S2 variant A / the "pressure-level interp" of S4: mt_interplevel over ALL columns of a
    (60, 1000, 1000) field, 4 variables per launch, to 60 height levels and to 60 pressure levels
    (decreasing coordinate).  Bytes = 8 * columns * (nz * (V + 1) + nlev * V)."""
    lib = _lib.load()
    nz, ny, nx, V, nlev = 60, 1000, 1000, 4, 60
    gen = torch.Generator(device="cuda").manual_seed(3)
    k = torch.arange(nz, device="cuda", dtype=torch.float64).reshape(-1, 1, 1)
    hgt = 1500.0 * torch.rand((ny, nx), generator=gen, device="cuda", dtype=torch.float64)
    z3d = hgt[None] + (20000.0 - hgt[None]) * (k / (nz - 1)) ** 1.5
    p3d = 1e5 * torch.exp(-z3d / 8000.0)
    flds = [torch.randn((nz, ny, nx), generator=gen, device="cuda", dtype=torch.float64) for _ in range(V)]
    outs = [torch.empty((ny, nx, nlev), device="cuda", dtype=torch.float64) for _ in range(V)]
    rows = []
    for name, vert, lev in (("height levels", z3d, np.arange(nlev) * 300.0 + 100.0),
                            ("pressure levels (decreasing coordinate)", p3d, np.linspace(98000.0, 9000.0, nlev))):
        lv = D.to_device(lev)
        order = 1 if lev[1] > lev[0] else -1
        call = lambda: _lib.check(lib.mt_interplevel(  # noqa: E731
            D.ptr_array(flds), V, D.ptr(vert), _lib.MT_F64, nz, ny, nx, 0, ny, 0, nx, D.ptr(lv), nlev, order, -1,
            D.ptr_array(outs), 1, nx * nlev, nlev, 0, D.stream()))
        ms = timed(call, reps)
        rows.append(row(f"mt_interplevel, {V} variables, ({nz},{ny},{nx}) -> {nlev} {name}", ms,
                        8.0 * ny * nx * (nz * (V + 1) + nlev * V), V * nlev * ny * nx))
    return rows


def bench_S1_legacy(reps):
    """config[0]: 400x400 -> 200r x 360theta through the LEGACY host-pointer symbol (H2D+kernel+D2H)."""
    import ctypes as ct
    lib = _lib.load()
    rng = np.random.default_rng(0)
    f = np.asfortranarray(rng.standard_normal((400, 400)))
    r = np.arange(0, 200 * 2000., 2000.)
    th = np.arange(0, 360 * (2 * np.pi / 360), 2 * np.pi / 360)
    rr, tt = np.meshgrid(r, th)
    x = np.ascontiguousarray((rr * np.cos(tt) + 399000.0).reshape(-1))
    y = np.ascontiguousarray((rr * np.sin(tt) + 399000.0).reshape(-1))
    out = np.empty(x.size)
    p = lambda a: a.ctypes.data_as(ct.POINTER(ct.c_double))  # noqa: E731
    call = lambda: lib.interp2D_fast(400, 400, 2000.0, 2000.0, x.size, p(x), p(y), p(f), p(out))  # noqa: E731
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    ms = (time.perf_counter() - t0) / reps * 1e3
    import oracle
    olib = oracle.lib()
    out2 = np.empty(x.size)
    t0 = time.perf_counter()
    for _ in range(reps):
        olib.interp2D_fast(400, 400, 2000.0, 2000.0, x.size, p(x), p(y), p(f), p(out2))
    ms_cpu = (time.perf_counter() - t0) / reps * 1e3
    assert np.array_equal(out, out2)
    return [dict(kernel="S1 interp2D_fast legacy symbol (host pointers, e2e)", ms=ms, points_per_s=x.size / ms * 1e3,
                 cpu_oracle_ms=ms_cpu, note="tiny problem: dominated by the synchronous H2D/D2H round trip")]


def sharded(reps):
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    nz, ny, nx = 80, 2000, 2000
    lo, hi, hlo, hhi = level_shard(nz, rank, world)
    names = ["u", "v", "w", "p", "T", "qv", "qc", "qr", "qi", "qs", "qg"]

    def level(k):   # deterministic per global level, so that every rank can rebuild its neighbours
        gen = torch.Generator(device="cuda").manual_seed(1000 + k)
        one = fields((1, ny, nx), gen, names)
        zz = k * 250.0 + 100.0
        one["p"] = one["p"] - 1e5 * np.exp(-100.0 / 8000.0) + 1e5 * np.exp(-zz / 8000.0)
        one["T"] = one["T"] + 6.5e-3 * 100.0 - 6.5e-3 * zz
        return one
    own = [level(k) for k in range(lo, hi)]
    local = {n: torch.cat([lv[n] for lv in own]) for n in names}
    f2d = torch.full((ny, nx), 5e-5, device="cuda", dtype=torch.float64)
    outs = list(_lib.CAR_D_OUT)
    spac = (250.0, 2000.0, 2000.0)
    res = dyn_level_sharded("car", local, f2d, spac, (lo, hi, hlo, hhi), nz, outs, rank, world, dist=dist)
    # check the shard-boundary levels against a local run that is GIVEN the true neighbouring levels
    ext_lo, ext_hi = max(lo - 1, 0), min(hi + 1, nz)
    ext = [level(k) for k in range(ext_lo, min(ext_lo + 4, ext_hi))]
    chk = LevelShardDyn("car", {n: torch.cat([lv[n] for lv in ext]) for n in names}, f2d, spac,
                        (ext_lo, ext_lo + len(ext), 0, 0), nz, ["zeta_r", "div", "PV"])
    # (the check block is only exact for levels whose neighbours are inside the block)
    chk.stage_a(); chk.stage_b()
    c = chk.result()
    k_glob = lo if lo > 0 else lo + 1            # first owned level that has a lower neighbour
    ok = all(torch.equal(res[n][k_glob - lo], c[n][k_glob - ext_lo]) for n in ("zeta_r", "div", "PV")) \
        if (k_glob - ext_lo + 1) < len(ext) else True
    torch.cuda.synchronize()
    dist.barrier()
    job = LevelShardDyn("car", local, f2d, spac, (lo, hi, hlo, hhi), nz, outs)
    from meteotools_b200.pipeline import halo_exchange_levels

    def step():
        job.stage_a()
        if world > 1:
            torch.cuda.current_stream().synchronize()
            halo_exchange_levels(job.halo_fields, hlo, hhi, rank, world, dist)
        job.stage_b()
    for _ in range(3):
        step()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    okt = torch.tensor([1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        npts = nz * ny * nx
        print(json.dumps(dict(kernel=f"car_d level-sharded ({nz},{ny},{nx}), NCCL halo of v,w,Th_rho", n_gpus=world,
                              ms=float(ms), points_per_s=npts / float(ms) * 1e3,
                              alg_bytes=8.0 * npts * (len(names) + len(outs)),
                              achieved_gbs_aggregate=8.0 * npts * (len(names) + len(outs)) / float(ms) / 1e6,
                              boundary_levels_match_unsharded=bool(okt.item() == 1.0), scaling="strong")))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--sharded", action="store_true")
    a = ap.parse_args()
    reps = 3 if a.quick else 10
    if a.sharded:
        return sharded(reps)
    rows = []
    rows += bench_S1_legacy(20)
    rows += bench_dyn("cyl", (60, 360, 300), (250.0, 2 * np.pi / 360, 1000.0), reps * 3)
    rows += bench_fd((80, 1000, 1000), reps)
    rows += bench_interplevel(reps)
    rows += bench_dyn("car", (80, 2000, 2000), (250.0, 2000.0, 2000.0), reps)
    print(json.dumps(dict(peak_gbs=PEAK, peak_source=PEAK_SRC, rows=rows), indent=1))


if __name__ == "__main__":
    main()
