#!/usr/bin/env python
"""bench.py -- headline benchmark of meteotools_b200 (contract: see the task statement / DESIGN.md).

Workload "S2+cyl_td" (BASELINE.json configs[1] + the north star's target): ONE step = one WRF
time step of a 1000x1000x60 float64 field:
    cart->cyl ingest (cylindrical_grid.set_data: vertical interpolation to 60 height levels +
    bilinear gather + terrain mask) of the 12 3-D fields and 2 2-D fields of the reference's
    test_cyl_thermaldynamics.py:44-56 call, onto the (60 z, 360 theta, 300 r) grid, followed by
    the full cyl_td diagnostic set (13 outputs, incl. the calc_Tc fixed point).
metric  = output grid points/s: target-grid points (60*360*300 = 6.48 M per step) per second
          through the whole step.
value   = device-resident: sources already in HBM (full 1000x1000x60 arrays), CUDA-event timed.
e2e     = the public grid API with HOST (pinned) buffers: H2D of the touched boxes, kernels, D2H
          of the 13 thermodynamic outputs, wall clock around synchronised steps.
N > 1   = one process per GPU, every rank processes its own time steps (weak scaling, no
          data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NZ, NY, NX, DX = 60, 1000, 1000, 2000.0
CYL = dict(Nr=300, Ntheta=360, Nz=60, dr=1000, dtheta=2 * np.pi / 360, dz=300, z_start=100, r_start=0,
           theta_start=0, vertical_coord_type="z", dt=300)
NAMES3D = ["u", "v", "w", "p", "T", "Th", "qv", "qc", "qr", "qi", "qs", "qg"]
NAMES2D = ["f", "hgt"]
NPTS = CYL["Nz"] * CYL["Ntheta"] * CYL["Nr"]
METRIC = "output grid points/sec (cart->cyl set_data + cyl_td diagnostics, 1000x1000x60 f64 WRF field)"
UNIT = "grid points/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ synthetic WRF step
def make_step(seed, pinned):
    """SURVEY 8(d) S2-style source: terrain-following heights over Gaussian hills (so that low
    target levels fall below ground: NaN path), smooth + noise fields.  Returns host arrays
    (pinned when `pinned`), all float64 C-order."""
    rng = np.random.default_rng(seed)
    if pinned:
        import torch

        def alloc(shape):
            return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    else:
        def alloc(shape):
            return np.empty(shape)
    yy, xx = np.meshgrid(np.arange(NY) * DX, np.arange(NX) * DX, indexing="ij")
    cx, cy = (NX - 1) * DX / 2, (NY - 1) * DX / 2
    hgt = alloc((NY, NX))
    hgt[:] = (1500.0 * np.exp(-((xx - cx - 150e3) ** 2 + (yy - cy + 80e3) ** 2) / 120e3 ** 2)
              + 900.0 * np.exp(-((xx - cx + 200e3) ** 2 + (yy - cy - 100e3) ** 2) / 90e3 ** 2))
    f2d = alloc((NY, NX))
    f2d[:] = 5e-5 + 2e-11 * (yy - cy)
    k = (np.arange(NZ) / (NZ - 1)) ** 1.5
    z3d = alloc((NZ, NY, NX))
    np.multiply((20000.0 - hgt)[None], k[:, None, None], out=z3d)
    z3d += hgt[None]
    noise = [rng.standard_normal((NZ, NY, NX), dtype=np.float32) for _ in range(3)]
    v = {n: alloc((NZ, NY, NX)) for n in NAMES3D}
    swirl = 30.0 * np.sin(2 * np.pi * xx / (NX * DX)) * np.cos(2 * np.pi * yy / (NY * DX))
    np.multiply(noise[0], 3.0, out=v["u"]); v["u"] += swirl[None]
    np.multiply(noise[1], 3.0, out=v["v"]); v["v"] -= swirl.T[None]
    np.multiply(noise[2], 0.5, out=v["w"])
    np.divide(z3d, -8000.0, out=v["p"]); np.exp(v["p"], out=v["p"]); v["p"] *= 1e5; v["p"] += noise[1]
    np.multiply(z3d, -6.5e-3, out=v["T"]); v["T"] += 300.0; v["T"] += noise[2]
    np.divide(1e5, v["p"], out=v["Th"]); np.power(v["Th"], 287.0 / 1004.0, out=v["Th"]); v["Th"] *= v["T"]
    np.abs(noise[0], out=noise[0]); np.abs(noise[1], out=noise[1]); np.abs(noise[2], out=noise[2])
    np.multiply(noise[0], 1e-3, out=v["qv"]); v["qv"] += 1e-3
    for i, n in enumerate(("qc", "qr", "qi", "qs", "qg")):
        np.multiply(noise[i % 3][::-1] if i >= 3 else noise[i], 1e-4, out=v[n])
    return dict(z3d=z3d, hgt=hgt, f=f2d, vars3d=v)


def make_grid():
    from meteotools_b200.grids import cylindrical_grid
    cyl = cylindrical_grid(dict(CYL))
    cyl.set_horizontal_location(((NX + 1) / 2 - 1) * DX, ((NY + 1) / 2 - 1) * DX)   # wrfout_grid.py:69-70
    return cyl


# ------------------------------------------------------------------------------ clocks sampler
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU reference step
TD_CHAIN = ["calc_theta", "calc_saturated_vapor", "calc_saturated_qv", "calc_vapor", "calc_RH", "calc_Td",
            "calc_theta_es", "calc_theta_v", "calc_Tv", "calc_rho", "calc_Tc", "calc_theta_e", "calc_moist_dTdz"]
TD_NAMES = ["Th", "vapor_s", "qvs", "vapor", "RH", "Td", "Th_es", "Th_v", "Tv", "rho", "Tc", "Th_e", "moist_dTdz"]


class ReferenceStep:
    """ONE WHOLE step of the workload on the host, nothing sampled or extrapolated: the 12 3-D + 2 2-D
    fields through cylindrical_grid.set_data (grids/cylindrical_grid.py:95-137: full-domain
    wrf.interplevel per variable, asfortranarray copy, interp2D_fast_layers, terrain index + mask) and
    then the cyl_td chain of test_cyl_thermaldynamics.py:96-260 ON THAT STEP'S OUTPUT.

    kind "reference": the UNMODIFIED reference Python (baseline/_ref/meteotools, copied from
    /root/reference by build()) drives the step; its lib/fastcompute.so is the oracle's C/OpenMP
    restatement of fastcompute.f90 (gfortran is absent) and wrf.interplevel the oracle's DINTERP3DZ.
    kind "port": baseline/_ref is absent -> the oracle's restatement of the same Python layer."""

    def __init__(self, threads):
        import oracle
        from oracle import refpkg
        oracle.build()
        os.environ["OMP_NUM_THREADS"] = str(threads)
        self.threads = threads
        self.mt = refpkg.import_reference()
        self.kind = "reference" if self.mt is not None else "port"
        if self.mt is None:
            r = np.arange(0, CYL["Nr"] * float(CYL["dr"]), float(CYL["dr"]))
            th = np.arange(0, CYL["Ntheta"] * float(CYL["dtheta"]), float(CYL["dtheta"]))
            self.z = np.arange(0, CYL["Nz"] * float(CYL["dz"]), float(CYL["dz"])) + float(CYL["z_start"])
            rr, tt = np.meshgrid(r, th)
            cx = ((NX + 1) / 2 - 1) * DX
            self.xTR, self.yTR = rr * np.cos(tt) + cx, rr * np.sin(tt) + cx

    def describe(self):
        if self.kind == "reference":
            return ("unmodified reference Python (cylindrical_grid.set_data + the 13 calc_* methods) from baseline/_ref, "
                    "fastcompute.so = gcc/OpenMP restatement of fastcompute.f90 (no gfortran in the image), "
                    "wrf.interplevel = DINTERP3DZ restatement")
        return "oracle port of the reference's Python layer + fastcompute.f90 restatement (baseline/_ref absent)"

    def run(self, step):
        """Returns (seconds, dict of the 13 outputs)."""
        if self.kind == "reference":
            cyl = self.mt.grids.cylindrical_grid(dict(CYL))      # untimed: once per run in real use
            cyl.set_horizontal_location(((NX + 1) / 2 - 1) * DX, ((NY + 1) / 2 - 1) * DX)
            t0 = time.perf_counter()
            with np.errstate(all="ignore"):
                cyl.set_data(DX, DX, step["z3d"], f=step["f"], hgt=step["hgt"], **step["vars3d"])
                for m in TD_CHAIN:
                    getattr(cyl, m)()
            dt = time.perf_counter() - t0
            return dt, {k: getattr(cyl, k) for k in TD_NAMES}
        import oracle
        from oracle import calc_oracle as co
        t0 = time.perf_counter()
        hgt_c = oracle.interp2D_fast(DX, DX, self.xTR, self.yTR, step["hgt"])
        oracle.interp2D_fast(DX, DX, self.xTR, self.yTR, step["f"])
        idx = co.cyl_terrain_index(hgt_c, float(CYL["z_start"]), float(CYL["dz"]))
        lev = [oracle.interplevel(step["vars3d"][n], step["z3d"], self.z) for n in NAMES3D]   # set_data:109-113
        f = {}
        for n, l in zip(NAMES3D, lev):
            f[n] = co.cyl_terrain_mask(oracle.interp2D_fast_layers(DX, DX, self.xTR, self.yTR, l), idx)
        del lev
        with np.errstate(all="ignore"):
            o = co.cyl_td_set(f["T"], f["p"], f["qv"], f["qc"], f["qr"])
        dt = time.perf_counter() - t0
        return dt, {k: o[k] for k in TD_NAMES}


def cpu_baseline(step, threads, steps=2):
    """The repo arm's CPU leg: `steps` WHOLE steps (after one untimed one) of the same workload."""
    ref = ReferenceStep(threads)
    ref.run(step)
    ts = [ref.run(step)[0] for _ in range(steps)]
    sec = float(np.mean(ts))
    return {"value": NPTS / sec, "unit": UNIT, "cores": threads, "kind": ref.kind,
            "sample": f"{steps} whole steps (all 12 3-D + 2 2-D fields + the 13 cyl_td outputs of that step's own "
                      f"output), nothing extrapolated; {ref.describe()}; {sec:.2f} s/step",
            "seconds_per_step": sec}


def run_reference(args):
    """Reference arm: every timed step is the WHOLE step on the host (ReferenceStep), all host cores."""
    quiet_stdout()   # the reference's @timer prints
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = ReferenceStep(threads)
    step = make_step(0, pinned=False)
    times, t_start = [], time.perf_counter()
    for i in range(args.warmup + args.steps):
        dt, out = ref.run(step)
        if i >= args.warmup:
            times.append(dt)
        # safety net far above the expected 25 x ~10 s: never run into the driver's 1800 s limit; the line
        # then reports the number of steps that WERE timed (nothing is extrapolated either way)
        if time.perf_counter() - t_start > 1300 and len(times) >= 2:
            break
    assert out["Th_e"].shape == (CYL["Nz"], CYL["Ntheta"], CYL["Nr"])
    sec = float(np.mean(times))
    val = NPTS / sec
    sample_txt = (f"{len(times)} timed whole steps after {args.warmup} warm-up steps, nothing sampled or extrapolated; "
                  + ref.describe())
    emit(({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(), "timed_seconds": float(np.sum(times)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": ref.kind, "sample": sample_txt},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def config_dict():
    return {"workload": "S2+cyl_td: cylindrical_grid.set_data (12 3-D + 2 2-D fields, 1000x1000x60 f64 -> 60z x 360theta x 300r, "
                        "terrain-masked) + full cyl_td set (13 outputs)",
            "source": [NZ, NY, NX], "target": [CYL["Nz"], CYL["Ntheta"], CYL["Nr"]],
            "points_per_step": NPTS, "l2_policy": "inputs larger than L2 (touched source boxes 13 x 43 MB + 25 x 52 MB outputs per step >> 126 MB)",
            "parallelism": "time steps sharded over ranks, no collective"}


# ------------------------------------------------------------------------------ main (b200 arm)
_REAL_STDOUT = None


def quiet_stdout():
    """Everything that libraries print to stdout (NCCL's version banner, torchrun notices) goes to
    stderr; the ONE JSON line of the contract is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def bind_near_gpu(local_rank):
    """Multi-rank runs: keep this rank's threads (and so the first-touch placement of its pinned host
    buffers) on the CPUs that NVML reports as local to its GPU.  On a multi-socket box the e2e leg
    otherwise crosses the socket interconnect for half the ranks.  Best effort: silently skipped when
    NVML or the affinity call is not available or the container restricts the CPU set."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def link_probe(h2d_bytes, d2h_bytes, barrier, reps=10):
    """ms for moving exactly one step's bytes over the host link with nothing else going on: one
    cudaMemcpyAsync per direction, pinned buffers, two streams, both directions at once."""
    import torch
    hs = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    hd = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    ds = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dd = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ts = []
    for i in range(reps + 2):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            ds.copy_(hs, non_blocking=True)
        with torch.cuda.stream(s2):
            hd.copy_(dd, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        if i >= 2:
            ts.append((time.perf_counter() - t0) * 1e3)
    del hs, hd, ds, dd
    torch.cuda.empty_cache()
    return float(np.median(ts))


def placement(local_rank, cpus_bound):
    """Where this rank's GPU and host threads sit: PCI bus id, the GPU's NUMA node (sysfs), the CPUs the
    process may run on."""
    info = {"local_rank": local_rank, "cpus_bound_near_gpu": cpus_bound,
            "cpus_allowed": len(os.sched_getaffinity(0))}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        addr = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        info["pci"] = addr
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as fh:
            info["gpu_numa_node"] = int(fh.read().strip())
    except Exception as e:   # not every VM exposes it
        info["gpu_numa_node"] = None
        info["note"] = str(e)[:80]
    try:
        with open("/sys/devices/system/node/online") as fh:
            info["numa_nodes_online"] = fh.read().strip()
    except OSError:
        pass
    return info


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--streams", type=int, default=2, help="time steps in flight in the device-resident leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary rows (S3 / S5 / FD2 / interplevel / level shards)")
    ap.add_argument("--min-timed-s", type=float, default=1.0, help="repeat the timed block of K steps until this much GPU time")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from meteotools_b200 import _device as D, _lib
    from meteotools_b200.pipeline import CylStepPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    near = bind_near_gpu(local) if world > 1 and os.environ.get("MT_BENCH_NO_BIND") != "1" else 0
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = make_step(rank, pinned=True)
    cyl = make_grid()
    # ---------------- value: device-resident sources -------------------------------------------
    dev3d = {k: torch.from_numpy(v).cuda() for k, v in step["vars3d"].items()}
    dev2d = {"f": torch.from_numpy(step["f"]).cuda(), "hgt": torch.from_numpy(step["hgt"]).cuda()}
    devz = torch.from_numpy(step["z3d"]).cuda()
    # Two pipeline slots on two CUDA streams: consecutive time steps are independent, so step k+1's
    # memory-bound ingest kernels may overlap step k's FP64-bound thermodynamics (--streams 1: off)
    nslots = max(1, args.streams)
    pipes = [CylStepPipeline(cyl, (NZ, NY, NX), DX, DX, NAMES3D, names2d=("f",), hgt=True, resident=True)
             for _ in range(nslots)]
    streams = [torch.cuda.Stream() for _ in range(nslots)]
    for p_ in pipes:
        p_.bind_device(devz, dev3d, dev2d, direction=0)
    pipe = pipes[0]

    def run_steps(k):
        main = torch.cuda.current_stream()
        for s_ in streams:
            s_.wait_stream(main)
        for i in range(k):
            with torch.cuda.stream(streams[i % nslots]):
                pipes[i % nslots].run()
        for s_ in streams:
            main.wait_stream(s_)
    clocks = Clocks(local)   # sampled over warm-up + timed region + roofline pass (same kernels)
    clocks.start()
    run_steps(args.warmup * nslots)
    barrier()
    launches0 = lib.mt_launch_count()
    # The timed block is EXACTLY K steps between barrier + synchronize on both sides; the block is repeated
    # until >= --min-timed-s of GPU time has been measured (sustained behaviour, >= 10 clock samples under
    # load) and the MEDIAN block is reported (every block is listed in `timed_blocks_ms`).
    blocks = []
    while True:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(args.steps)
        e1.record()
        torch.cuda.synchronize()
        blocks.append(e0.elapsed_time(e1))
        barrier()
        done = torch.tensor([1.0 if (sum(blocks) >= args.min_timed_s * 1e3 or len(blocks) >= 400) else 0.0],
                            device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(done, op=dist.ReduceOp.MIN)   # every rank runs the same number of blocks
        if done.item() == 1.0:
            break
    launches = (lib.mt_launch_count() - launches0) // len(blocks)
    blk = torch.tensor(blocks, device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(blk, op=dist.ReduceOp.MAX)        # per block: the slowest rank
    blocks = [float(x) for x in blk.tolist()]
    ms = float(np.median(blocks)) / args.steps
    value = world * NPTS / (ms * 1e-3)

    # ---------------- roofline: per-launch CUDA events over the same steps ----------------------
    lib.mt_profile_enable(1)
    for _ in range(args.steps):
        pipe.run()
    buf = ct.create_string_buffer(1 << 16)
    _lib.check(lib.mt_profile_fetch(buf, len(buf)))
    lib.mt_profile_enable(0)
    tc_iters = int(D.to_host(pipe.scratch[:16])[15])
    kern = {}
    for line in buf.value.decode().strip().splitlines():
        name, cnt, tot = line.split()
        kern[name] = (int(cnt), float(tot))
    roof = roofline(kern, pipe, cyl, tc_iters, args.steps)
    clk = clocks.stop()

    # ---------------- e2e: public API, host (pinned) buffers ------------------------------------
    # (a) one step at a time through the grid methods; (b) a stream of time steps through
    # pipeline.TimeStepStream, which overlaps step k's D2H with step k+1's H2D (full-duplex PCIe).
    # Every step uploads its source boxes and downloads its 13 thermodynamic outputs in both.
    del pipe, pipes, dev3d, devz
    torch.cuda.empty_cache()
    td_names = list(_lib.TD_OUT)

    def e2e_step():
        cyl.set_data.__wrapped__(cyl, DX, DX, step["z3d"], f=step["f"], hgt=step["hgt"], **step["vars3d"])
        cyl.calc_all_thermaldynamic()
        return cyl.to_host(*td_names)
    for _ in range(2):
        res = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        res = e2e_step()
    torch.cuda.synchronize()
    e2e_single_s = (time.perf_counter() - t0) / args.e2e_steps
    assert all(res[k].shape == (CYL["Nz"], CYL["Ntheta"], CYL["Nr"]) for k in td_names)
    del res
    from meteotools_b200.pipeline import TimeStepStream
    ts = TimeStepStream(cyl, (NZ, NY, NX), DX, DX, NAMES3D, names2d=("f",), hgt=True, outputs=td_names,
                        slots=int(os.environ.get("MT_E2E_SLOTS", 3)))
    n_stream = max(4 * args.e2e_steps, 8)
    for _ in ts.process([step] * 3):
        pass
    barrier()
    t0 = time.perf_counter()
    got = 0
    for _, r in ts.process([step] * n_stream):
        got += 1
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / n_stream
    assert got == n_stream and r["Th_e"].shape == (CYL["Nz"], CYL["Ntheta"], CYL["Nr"])
    # (c) the same stream with float32 HOST sources -- what a wrfout file holds: half
    # the H2D bytes, widened on the device; results as the reference gets them from float32 input
    # (wrf-python rounds the level-interpolated field to float32).  Reported beside the headline.
    del ts
    torch.cuda.empty_cache()

    def pinned32(a):
        h = torch.empty(a.shape, dtype=torch.float32, pin_memory=True).numpy()
        np.copyto(h, a, casting="same_kind")
        return h
    step32 = dict(z3d=pinned32(step["z3d"]), hgt=step["hgt"], f=step["f"],
                  vars3d={k: pinned32(v) for k, v in step["vars3d"].items()})
    ts32 = TimeStepStream(cyl, (NZ, NY, NX), DX, DX, NAMES3D, names2d=("f",), hgt=True, outputs=td_names,
                          dtype=np.float32)
    for _ in ts32.process([step32] * 3):
        pass
    barrier()
    t0 = time.perf_counter()
    for _, r in ts32.process([step32] * n_stream):
        pass
    torch.cuda.synchronize()
    e2e32_s = (time.perf_counter() - t0) / n_stream
    del ts32
    torch.cuda.empty_cache()
    # (d) float32 in AND out: the results are rounded to float32 on the device, as the reference's
    # netCDF cache stores them ('f4', grid_general.py:119-129) -- half the download, which is what
    # bounds the stream.  Reported beside the headline, which keeps the reference's float64 arrays.
    ts32io = TimeStepStream(cyl, (NZ, NY, NX), DX, DX, NAMES3D, names2d=("f",), hgt=True, outputs=td_names,
                            dtype=np.float32, out_dtype=np.float32)
    for _ in ts32io.process([step32] * 3):
        pass
    barrier()
    t0 = time.perf_counter()
    for _, r in ts32io.process([step32] * n_stream):
        pass
    torch.cuda.synchronize()
    e2e32io_s = (time.perf_counter() - t0) / n_stream
    assert r["Th_e"].dtype == np.float32
    del ts32io, step32
    y0, y1, x0, x1 = cyl._target_box(NX, NY, DX, DX)
    h2d = (len(NAMES3D) + 1) * NZ * (y1 - y0) * (x1 - x0) * 8 + 2 * NY * NX * 8
    d2h = len(td_names) * NPTS * 8
    # (e) everything the reference's set_data + calc_* leave on the grid: the 12 interpolated fields too
    all_names = td_names + [n for n in NAMES3D if n not in td_names]
    ts_all = TimeStepStream(cyl, (NZ, NY, NX), DX, DX, NAMES3D, names2d=("f",), hgt=True, outputs=all_names,
                            slots=int(os.environ.get("MT_E2E_SLOTS", 3)))
    for _ in ts_all.process([step] * 3):
        pass
    barrier()
    t0 = time.perf_counter()
    for _, r in ts_all.process([step] * n_stream):
        pass
    torch.cuda.synchronize()
    e2e_all_s = (time.perf_counter() - t0) / n_stream
    assert r["u"].shape == (CYL["Nz"], CYL["Ntheta"], CYL["Nr"])
    d2h_all = len(all_names) * NPTS * 8
    del ts_all, r
    torch.cuda.empty_cache()
    # (f) what the host<->device link alone gives for exactly these bytes, all ranks at once
    link_ms = link_probe(h2d, d2h, barrier)
    link_all_ms = link_probe(h2d, d2h_all, barrier)
    if world > 1:
        te = torch.tensor([e2e_s, e2e_single_s, e2e32_s, e2e32io_s, e2e_all_s, link_ms, link_all_ms], device="cuda",
                          dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s, e2e_single_s, e2e32_s, e2e32io_s, e2e_all_s, link_ms, link_all_ms = (float(x) for x in te.tolist())
    place = placement(local, near)
    if world > 1:
        places = [None] * world
        dist.all_gather_object(places, place)
    else:
        places = [place]
    # ---------------- secondary rows (SURVEY 8d S3 / S5 / FD2 / interplevel; level shards at N > 1) ----------
    secondary = []
    if not args.no_secondary:
        import bench_rows
        if world == 1:
            secondary = bench_rows.secondary_single_gpu(reps=10)
        else:
            unsharded = None
            if rank == 0:   # the strong-scaling reference: the same S5 on one GPU
                r5 = bench_rows.bench_dyn("car", bench_rows.CAR_S5, 5, "S5 car_d, 17 outputs, (80,2000,2000), one GPU (rank 0)")
                unsharded = r5["ms"]
                secondary.append(r5)
            barrier()
            # the level-sharded S5 with both transports of the halo planes: NVLink peer memory (copy engines +
            # stream memory operations, the default of pipeline.LevelShardDyn users on one node) and NCCL
            for transport in ("peer", "nccl"):
                sh = bench_rows.sharded_car_d(rank, world, dist, reps=10, unsharded_ms=unsharded, transport=transport)
                if sh is not None:
                    secondary.append(sh)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(),
            "clocks": clk, "gpu_launches": int(launches), "steps_in_flight": nslots, "cpus_bound_near_gpu": near,
            "timed_reps": len(blocks), "timed_region_s": sum(blocks) * 1e-3,
            "timed_blocks_ms": {"median": float(np.median(blocks)), "min": float(np.min(blocks)),
                                "max": float(np.max(blocks)), "steps_per_block": args.steps},
            "e2e": {"value": world * NPTS / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": n_stream,
                    "link_floor_ms": link_ms, "frac_of_link": link_ms / (e2e_s * 1e3),
                    "link_probe": "one cudaMemcpyAsync of h2d_bytes_per_step (pinned -> device) and one of "
                                  "d2h_bytes_per_step (device -> pinned) on two streams at once, all ranks together; "
                                  "median of 10",
                    "placement": places,
                    "all_fields": {"value": world * NPTS / e2e_all_s, "ms_per_step": e2e_all_s * 1e3,
                                   "d2h_bytes_per_step": int(d2h_all), "link_floor_ms": link_all_ms,
                                   "frac_of_link": link_all_ms / (e2e_all_s * 1e3),
                                   "note": "the 13 cyl_td outputs AND the 12 interpolated set_data fields downloaded: "
                                           "everything the reference's set_data + calc_* leave on the grid as host "
                                           "arrays (the headline downloads the 13 outputs; the others stay on the "
                                           "device until read: 52 MB each)"},
                    "path": "pipeline.TimeStepStream.process (pinned host arrays in -> pinned host arrays out, "
                            "three slots: D2H of step k overlaps H2D of step k+1; the link alone takes 13.9 ms for these bytes, "
                            "scratch/pcie_probe.py)",
                    "single_step": {"value": world * NPTS / e2e_single_s, "ms_per_step": e2e_single_s * 1e3,
                                    "path": "cylindrical_grid.set_data + calc_all_thermaldynamic + to_host, "
                                            "one step at a time"},
                    "float32_sources": {"value": world * NPTS / e2e32_s, "ms_per_step": e2e32_s * 1e3,
                                        "h2d_bytes_per_step": int((h2d - 2 * NY * NX * 8) // 2 + 2 * NY * NX * 8),
                                        "note": "same stream fed with float32 host arrays (what wrfout files hold): "
                                                "half the upload, widened on the device"},
                    "float32_in_and_out": {"value": world * NPTS / e2e32io_s, "ms_per_step": e2e32io_s * 1e3,
                                           "h2d_bytes_per_step": int((h2d - 2 * NY * NX * 8) // 2 + 2 * NY * NX * 8),
                                           "d2h_bytes_per_step": int(d2h // 2),
                                           "note": "float32 sources and results rounded to float32 on the device "
                                                   "(the precision of the reference's netCDF cache): half the "
                                                   "bytes both ways; NOT the headline, which keeps float64 results"}},
            "roofline": roof, "tc_iterations": tc_iters, "secondary": secondary,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(step, os.cpu_count() or 1)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def roofline(kern, pipe, cyl, tc_iters, steps):
    """Dominant kernel of the step: algorithmic bytes per launch / CUDA-event duration per launch."""
    peak, which = peaks()
    y0, y1, x0, x1 = pipe.box
    box = (y1 - y0) * (x1 - x0)
    V, nlev, nz, n = len(NAMES3D), CYL["Nz"], NZ, pipe.n
    # unique source columns touched by the four-corner stencils (SURVEY 8d "U")
    ix = D_to_host(pipe.plan[0])
    nx_full = NX
    cells = np.unique(np.concatenate([ix[:, 2] * nx_full + ix[:, 0], ix[:, 2] * nx_full + ix[:, 1],
                                      ix[:, 3] * nx_full + ix[:, 0], ix[:, 3] * nx_full + ix[:, 1]]))
    U = int(cells.size)
    alg = {
        # reads (V fields + coordinate) x nz x box columns, writes V x nlev x box columns
        "k_interplevel": 8.0 * box * (nz * (V + 1) + nlev * V),
        "k_interplevel_tile": 8.0 * box * (nz * (V + 1) + nlev * V),
        # 8*(L*N + L*U) per variable + the plan (32 B/point) once per variable slice
        "k_gather2d_lf": V * 8.0 * (nlev * n + nlev * U) + V * 32.0 * n,
        # fused set_data (SURVEY 8d formula B): touched source columns of V fields + the coordinate,
        # outputs, plan
        "k_setdata_fused": 8.0 * nz * U * (V + 1) + 8.0 * nlev * n * V + 32.0 * n,
        "k_setdata_tma": 8.0 * nz * U * (V + 1) + 8.0 * nlev * n * V + 32.0 * n,
        # the stop-test kernel: with NaN in the data (terrain) the ten certain sweeps were already
        # produced speculatively by k_td_main and this launch only reads the flag
        "k_tc_finish": None,
        # reads T,p,qv,qc,qr; writes the 13 outputs of the cyl_td set (Tc after the ten certain
        # sweeps and theta_e included) + the sweep-invariant g
        "k_td_main": 8.0 * NPTS * (5 + 14),
        # reads g, Tc; writes Tc
        "k_tc_iter": 8.0 * NPTS * 3,
    }
    rows = []
    for name, (cnt, tot) in kern.items():
        per = tot / cnt
        launches_per_step = cnt / steps
        eff_cnt = cnt
        a = alg.get(name)
        rows.append({"kernel": name, "launches_per_step": launches_per_step, "ms_per_step": tot / steps,
                     "avg_ms_per_launch": per,
                     "achieved_gbs": (a / (per * 1e-3) / 1e9) if a else None, "alg_bytes_per_launch": a})
    rows.sort(key=lambda r: -r["ms_per_step"])
    top = rows[0]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(top["kernel"])
    return {"bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_gbs"], "peak": peak,
            "peak_source": which, "unit": "GB/s",
            "frac": (top["achieved_gbs"] / peak) if top["achieved_gbs"] else None, "traffic": traffic,
            "unique_source_columns": U, "box_columns": box, "kernels": rows,
            "step_alg_bytes": sum(r["alg_bytes_per_launch"] * r["launches_per_step"] for r in rows
                                  if r["alg_bytes_per_launch"]),
            "note": "per-launch CUDA events (mt_profile_*) over the same steps, re-run after the timed region"}


def D_to_host(t):
    return t.detach().cpu().numpy().astype(np.int64)


if __name__ == "__main__":
    main()
