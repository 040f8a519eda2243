"""Per-row kernel benchmarks at the SURVEY 8(d) sizes (S1, S3, S5 + generic FD / vertical interpolation),
stand-alone form of bench.py's `secondary` rows (bench_rows.py).  Device-resident inputs, CUDA events,
algorithmic bytes / time vs the measured HBM peak.

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
import bench_rows  # noqa: E402

PEAK, PEAK_SRC = bench.peaks()


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
    unsharded = None
    if rank == 0 and os.environ.get("MT_SHARD_REF", "1") == "1":
        unsharded = bench_rows.bench_dyn("car", bench_rows.CAR_S5, 5, "S5 car_d, one GPU")["ms"]
    dist.barrier()
    for transport in os.environ.get("MT_SHARD_TRANSPORT", "nccl,peer").split(","):
        r = bench_rows.sharded_car_d(rank, world, dist, reps=reps, unsharded_ms=unsharded, transport=transport)
        if rank == 0:
            print(json.dumps(r))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--sharded", action="store_true")
    a = ap.parse_args()
    reps = 3 if a.quick else 10
    if a.sharded:
        return sharded(reps)
    rows = bench_S1_legacy(20) + bench_rows.secondary_single_gpu(reps, quick=a.quick)
    print(json.dumps(dict(peak_gbs=PEAK, peak_source=PEAK_SRC, rows=rows), indent=1))


if __name__ == "__main__":
    main()
