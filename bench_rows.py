"""Secondary rows of bench.py: the SURVEY 8(d) configurations the headline step does not exercise, timed in
the same driver run with CUDA events on device-resident synthetic data.

    S3   cyl_d, all 25 outputs, (60 z, 360 theta, 300 r): in the grid's own (theta, r, z) layout and as a
         (z, theta, r) array (the layout level shards use)
    S5   car_d, all 17 outputs, (80, 2000, 2000)
    FD2  along each axis of (80, 1000, 1000)
    IL   full-domain vertical interpolation (60, 1000, 1000) -> 60 levels, 12 variables per launch
    N>1  S5 split into level shards over the ranks, halo planes through mt_halo_exchange (NCCL) overlapped
         with the interior levels (pipeline.LevelShardDyn.run): strong scaling against the unsharded run

Every row: ms (median of the repetitions), alg_bytes (algorithmic bytes), frac of the measured
HBM peak.  Used by bench.py (``secondary``) and by profiles/bench_kernels.py (stand-alone).
"""
import ctypes as ct
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peak():
    import json
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"])
    return 6650.0


def timed(fn, reps, warm=2):
    """Median and minimum over `reps` individually event-timed calls (ms)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def row(name, ms, alg_bytes, points, **extra):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    return dict(row=name, ms=round(ms, 4), alg_bytes=alg_bytes, achieved_gbs=round(gbs, 1),
                frac=round(gbs / _peak(), 4), points_per_s=points / (ms * 1e-3), **extra)


def synth_fields(shape, zaxis, gen, names):
    """SURVEY 8(d) S3 / S5 distributions, generated on the device."""
    import torch
    nz = shape[zaxis]
    zz = torch.arange(nz, device="cuda", dtype=torch.float64) * 250.0 + 100.0
    view = [1, 1, 1]
    view[zaxis] = -1
    zz = zz.reshape(view)
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


DYN_IN = ["u", "v", "w", "p", "T", "qv", "qc", "qr", "qi", "qs", "qg"]


class DynCase:
    """mt_cyl_d / mt_car_d on one device-resident synthetic grid.  kind: "car" (ZTR), "cyl_trz" (the
    cylindrical grid's own layout), "cyl_ztr"."""

    def __init__(self, kind, nz, nt, nr, seed=1):
        import torch
        from meteotools_b200 import _lib
        self.lib, self._lib = _lib.load(), _lib
        cyl = kind.startswith("cyl")
        trz = kind == "cyl_trz"
        gen = torch.Generator(device="cuda").manual_seed(seed)
        shape = (nt, nr, nz) if trz else (nz, nt, nr)
        self.f = synth_fields(shape, 2 if trz else 0, gen, DYN_IN)
        self.f2d = torch.full((nt, nr), 5e-5, device="cuda", dtype=torch.float64) + \
            1e-9 * torch.arange(nt * nr, device="cuda", dtype=torch.float64).reshape(nt, nr)
        g = _lib.mt_grid_t()
        g.nz, g.nt, g.nr = nz, nt, nr
        g.layout = _lib.MT_LAYOUT_TRZ if trz else _lib.MT_LAYOUT_ZTR
        g.dz, g.dt, g.dr = (250.0, 2 * np.pi / nt, 1000.0) if cyl else (250.0, 2000.0, 2000.0)
        self.keep = []
        if cyl:
            th = np.arange(nt) * (2 * np.pi / nt)
            r = torch.arange(nr, device="cuda", dtype=torch.float64) * 1000.0
            st, ctb = torch.from_numpy(np.sin(th)).cuda(), torch.from_numpy(np.cos(th)).cuda()
            g.r, g.sin_t, g.cos_t = r.data_ptr(), st.data_ptr(), ctb.data_ptr()
            self.keep += [r, st, ctb]
        g.is_bottom = g.is_top = 1
        din = _lib.mt_dyn_in_t()
        for k in ("u", "v", "w", "p", "T", "qv"):
            setattr(din, k, self.f[k].data_ptr())
        din.f = self.f2d.data_ptr()
        for i, k in enumerate(("qc", "qr", "qi", "qs", "qg")):
            din.q[i] = self.f[k].data_ptr()
        self.names = list(_lib.CYL_D_OUT if cyl else _lib.CAR_D_OUT)
        dout = (_lib.mt_cyl_d_out_t if cyl else _lib.mt_car_d_out_t)()
        self.outs = {}
        for k in self.names:
            self.outs[k] = torch.empty(shape, dtype=torch.uint8 if k == "strain_region" else torch.float64,
                                       device="cuda")
            setattr(dout, k, self.outs[k].data_ptr())
        self.g, self.din, self.dout = g, din, dout
        self.call = self.lib.mt_cyl_d if cyl else self.lib.mt_car_d
        self.npts = nz * nt * nr
        self.alg_bytes = 8.0 * self.npts * (len(DYN_IN) + len(self.names))

    def run(self):
        from meteotools_b200 import _device as D
        self._lib.check(self.call(self.g, self.din, self.dout, self._lib.MT_STAGE_ALL, D.stream()), "dyn")

    def per_kernel(self, reps=3):
        import torch
        for _ in range(2):   # first launches of a process: module load, tensor maps, function attributes
            self.run()
        torch.cuda.synchronize()
        self.lib.mt_profile_enable(1)
        for _ in range(reps):
            self.run()
        buf = ct.create_string_buffer(1 << 14)
        self._lib.check(self.lib.mt_profile_fetch(buf, len(buf)))
        self.lib.mt_profile_enable(0)
        return {ln.split()[0]: round(float(ln.split()[2]) / int(ln.split()[1]), 4)
                for ln in buf.value.decode().strip().splitlines()}


def bench_dyn(kind, dims, reps, label):
    import torch
    case = DynCase(kind, *dims)
    per = case.per_kernel()
    ms, best = timed(case.run, reps)
    r = row(label, ms, case.alg_bytes, case.npts, best_ms=round(best, 4), n_in=len(DYN_IN), n_out=len(case.names),
            per_kernel_ms=per)
    del case
    torch.cuda.empty_cache()
    return r


def bench_fd(shape, reps):
    import torch
    from meteotools_b200 import _device as D, _lib
    lib = _lib.load()
    v = torch.randn(shape, device="cuda", dtype=torch.float64)
    out = torch.empty_like(v)
    rows = []
    n0, n1, n2 = shape
    for axis, (outer, n, inner) in enumerate(((1, n0, n1 * n2), (n0, n1, n2), (n0 * n1, n2, 1))):
        ms, best = timed(lambda: _lib.check(lib.mt_fd(D.ptr(v), D.ptr(out), outer, n, inner, 250.0, 0, D.stream())), reps)
        rows.append(row(f"FD2 axis {axis}, {tuple(shape)}", ms, 16.0 * v.numel(), v.numel(), best_ms=round(best, 4)))
    del v, out
    torch.cuda.empty_cache()
    return rows


def bench_interplevel(reps, V=12):
    """SURVEY 8(d) S2 variant A: vertical interpolation over ALL 10^6 columns of (60, 1000, 1000) fields to 60
    height levels, V variables per launch, result in the (nlev, ny, nx) order wrf.interplevel returns.
    Bytes = 8 * columns * (nz * (V + 1) + nlev * V)."""
    import torch
    from meteotools_b200 import _device as D, _lib
    lib = _lib.load()
    nz, ny, nx, nlev = 60, 1000, 1000, 60
    gen = torch.Generator(device="cuda").manual_seed(3)
    k = torch.arange(nz, device="cuda", dtype=torch.float64).reshape(-1, 1, 1)
    hgt = 1500.0 * torch.rand((ny, nx), generator=gen, device="cuda", dtype=torch.float64)
    z3d = hgt[None] + (20000.0 - hgt[None]) * (k / (nz - 1)) ** 1.5
    flds = [torch.randn((nz, ny, nx), generator=gen, device="cuda", dtype=torch.float64) for _ in range(V)]
    outs = [torch.empty((nlev, ny, nx), device="cuda", dtype=torch.float64) for _ in range(V)]
    lv = D.to_device(np.arange(nlev) * 300.0 + 100.0)
    fp, op = D.ptr_array(flds), D.ptr_array(outs)
    call = lambda: _lib.check(lib.mt_interplevel(  # noqa: E731
        fp, V, D.ptr(z3d), _lib.MT_F64, nz, ny, nx, 0, ny, 0, nx, D.ptr(lv), nlev, 1, -1,
        op, ny * nx, nx, 1, 0, D.stream()))
    ms, best = timed(call, reps)
    r = row(f"mt_interplevel, {V} variables, ({nz},{ny},{nx}) -> {nlev} height levels, level-major result", ms,
            8.0 * ny * nx * (nz * (V + 1) + nlev * V), V * nlev * ny * nx, best_ms=round(best, 4))
    del flds, outs, z3d
    torch.cuda.empty_cache()
    return r


# ------------------------------------------------------------------------------ level shards (N > 1)
CAR_S5 = (80, 2000, 2000)


def _level(k, ny, nx, names):
    """Global level k of the S5 fields, deterministic per level: every rank can rebuild any level."""
    import torch
    gen = torch.Generator(device="cuda").manual_seed(1000 + k)
    one = synth_fields((1, ny, nx), 0, gen, names)
    zz = k * 250.0 + 100.0
    one["p"] = one["p"] - 1e5 * np.exp(-100.0 / 8000.0) + 1e5 * np.exp(-zz / 8000.0)
    one["T"] = one["T"] + 6.5e-3 * 100.0 - 6.5e-3 * zz
    return one


def sharded_car_d(rank, world, dist, reps=10, dims=CAR_S5, unsharded_ms=None, transport="nccl"):
    """S5 car_d split into `world` level shards: the real NCCL halo (mt_halo_exchange on a side stream),
    overlapped with the interior levels.  Returns the row on rank 0 (None elsewhere)."""
    import torch
    from meteotools_b200 import _lib
    from meteotools_b200.pipeline import HaloComm, LevelShardDyn, PeerHalo, level_shard
    nz, ny, nx = dims
    lo, hi, hlo, hhi = level_shard(nz, rank, world)
    names = list(DYN_IN)
    own = [_level(k, ny, nx, names) for k in range(lo, hi)]
    local = {n: torch.cat([lv[n] for lv in own]) for n in names}
    del own
    f2d = torch.full((ny, nx), 5e-5, device="cuda", dtype=torch.float64)
    outs = list(_lib.CAR_D_OUT)
    spac = (250.0, 2000.0, 2000.0)
    comm = PeerHalo(rank, world, ny * nx, 3, dist) if transport == "peer" else HaloComm(rank, world, dist)
    job = LevelShardDyn("car", local, f2d, spac, (lo, hi, hlo, hhi), nz, outs)
    del local
    job.run(comm)
    torch.cuda.synchronize()
    # boundary levels against a local run that is GIVEN the true neighbouring levels: a block of 5 levels
    # around each shard boundary; exact for the block's inner levels (centred z-differences)
    ok = True
    chk_out = ["zeta_r", "zeta_t", "div", "PV", "wind_speed"]
    res = job.result()
    for glob in ([lo] if hlo else []) + ([hi - 1] if hhi else []):
        b0 = max(glob - 2, 0)
        b1 = min(b0 + 5, nz)
        blk = [_level(k, ny, nx, names) for k in range(b0, b1)]
        chk = LevelShardDyn("car", {n: torch.cat([lv[n] for lv in blk]) for n in names}, f2d, spac,
                            (b0, b1, 0, 0), nz, chk_out)
        chk.run(None)
        c = chk.result()
        for n in chk_out:
            ok = ok and bool(torch.equal(res[n][glob - lo], c[n][glob - b0]))
        del chk, blk, c
    torch.cuda.synchronize()
    dist.barrier()

    def step():
        job.run(comm)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record()
        step()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    # the exchange alone (all planes of one step, nothing to overlap with)
    hs = []
    for _ in range(5):
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        comm.exchange(job.halo_fields, hlo, hhi)
        e1.record()
        e1.synchronize()
        hs.append(e0.elapsed_time(e1))
    # the same work without overlap: exchange everything first, then the whole shard
    ns = []
    for _ in range(max(3, reps // 2)):
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if transport == "peer":      # (one exchange per step: the sequence numbers of the ranks must agree)
            job.stage_a()
            comm.exchange(job.halo_fields, hlo, hhi)
        else:
            comm.exchange(job.halo_inputs, hlo, hhi)
            job.stage_a()
            comm.exchange(job.halo_derived, hlo, hhi)
        job.stage_b()
        e1.record()
        e1.synchronize()
        ns.append(e0.elapsed_time(e1))
    # the shard's kernels alone, no exchange: as run() issues them (interior + boundary launches) and as one launch
    def only(fn, n=5):
        out = []
        for _ in range(n):
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return float(np.median(out))
    # timeline of the overlapped run: when each phase ends, ms after the start (median of 5, max over ranks)
    tl = {}
    for _ in range(5):
        dist.barrier()
        marks = {}
        job.run(comm, marks=marks)
        torch.cuda.synchronize()
        for k, ev in marks.items():
            if k != "start":
                tl.setdefault(k, []).append(marks["start"].elapsed_time(ev))
    tl_keys = ["inputs_exchanged", "stage_a", "theta_rho_exchanged", "interior", "end"]
    tl_t = torch.tensor([float(np.median(tl[k])) if k in tl else 0.0 for k in tl_keys], device="cuda", dtype=torch.float64)
    dist.all_reduce(tl_t, op=dist.ReduceOp.MAX)
    c_split = only(lambda: (job.stage_a(), job.stage_b_split()))
    c_whole = only(lambda: (job.stage_a(), job.stage_b()))
    stats = torch.tensor([float(np.median(ts)), float(np.median(hs)), float(np.median(ns)), 0.0 if ok else 1.0,
                          c_split, c_whole], device="cuda", dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    comm.close()
    del job
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    ms, halo_ms, serial_ms, bad, c_split, c_whole = (float(x) for x in stats.tolist())
    npts = nz * ny * nx
    alg = 8.0 * npts * (len(names) + len(outs))
    how = ("halo planes (v, w, theta_rho) by the copy engines into the neighbour's memory (NVLink peer memory, "
           "stream memory operations)" if transport == "peer" else "NCCL halo (v, w, theta_rho) overlapped")
    out = dict(row=f"car_d {dims} level-sharded over {world} ranks, {how}", n_gpus=world, transport=transport,
               ms=round(ms, 4), halo_ms=round(halo_ms, 4), ms_without_overlap=round(serial_ms, 4),
               ms_kernels_only_split_launches=round(c_split, 4), ms_kernels_only_one_launch=round(c_whole, 4),
               timeline_ms_after_start={k: round(float(v), 4) for k, v in zip(tl_keys, tl_t.tolist())},
               alg_bytes=alg, achieved_gbs_aggregate=round(alg / ms / 1e6, 1), points_per_s=npts / ms * 1e3,
               halo_bytes_per_rank=int(3 * ny * nx * 8 * ((1 if hlo else 0) + (1 if hhi else 0)) * 2),
               boundary_levels_match_unsharded=bool(bad == 0.0), scaling="strong",
               nccl_version=int(_lib.load().mt_nccl_version()))
    if unsharded_ms:
        out["unsharded_ms"] = round(unsharded_ms, 4)
        out["strong_scaling_efficiency"] = round(unsharded_ms / (world * ms), 4)
    return out


def secondary_single_gpu(reps=10, quick=False):
    rows = []
    rows.append(bench_dyn("cyl_trz", (60, 360, 300), reps * 3, "S3 cyl_d, 25 outputs, (60,360,300), grid layout (theta,r,z)"))
    rows.append(bench_dyn("cyl_ztr", (60, 360, 300), reps * 3, "S3 cyl_d, 25 outputs, (60,360,300), (z,theta,r) array"))
    rows += bench_fd((80, 1000, 1000), reps)
    rows.append(bench_interplevel(reps))
    if not quick:
        rows.append(bench_dyn("car", CAR_S5, reps, "S5 car_d, 17 outputs, (80,2000,2000)"))
    return rows
