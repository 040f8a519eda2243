"""Level-sharded cyl_d / car_d on ONE GPU: the shards of several ranks are emulated
in one process (stage A everywhere, halo levels copied by hand exactly as
pipeline.halo_exchange_levels would send them, stage B everywhere) and must reproduce the
unsharded result bit for bit -- including the levels next to every shard boundary."""
import numpy as np
import pytest

from conftest import assert_same
from golden.make_golden import CAR, CAR_OUT, CYL, CYL_OUT

pytestmark = pytest.mark.gpu


def _emulate(kind, fields, f2d, spacings, outputs, world, tables=None, split=False):
    import torch
    from meteotools_b200.pipeline import LevelShardDyn, level_shard
    nz = next(iter(fields.values())).shape[0]
    jobs = []
    for r in range(world):
        shard = level_shard(nz, r, world)
        local = {k: torch.from_numpy(v[shard[0]:shard[1]].copy()).cuda() for k, v in fields.items()}
        jobs.append((shard, LevelShardDyn(kind, local, f2d, spacings, shard, nz, outputs, tables)))
    for _, j in jobs:
        j.stage_a()
    for r, (shard, j) in enumerate(jobs):       # what halo_exchange_levels does, by hand
        lo, hi, hlo, hhi = shard
        for idx, f in enumerate(j.halo_fields):
            if hlo:
                nb = jobs[r - 1][1]
                f[0].copy_(nb.halo_fields[idx][nb.hlo + nb.nz_own - 1])
            if hhi:
                nb = jobs[r + 1][1]
                f[-1].copy_(nb.halo_fields[idx][nb.hlo])
    for _, j in jobs:
        if split:   # interior levels, then the boundary levels: the launches of LevelShardDyn.run()
            j.stage_b_split()
        else:
            j.stage_b()
    res = [j.result() for _, j in jobs]
    return {k: torch.cat([x[k] for x in res]).cpu().numpy() for k in outputs}


@pytest.mark.parametrize("world", [1, 2, 3])
def test_car_d_level_sharded(golden_grids, world):
    import torch
    g = golden_grids
    fields = {k[len("g_car_"):]: v for k, v in g.items() if k.startswith("g_car_") and v.ndim == 3}
    f2d = torch.from_numpy(g["g_car_f"]).cuda()
    outs = [k for k in CAR_OUT if k not in ("Th",)]
    got = _emulate("car", fields, f2d, (CAR["dz"], CAR["dy"], CAR["dx"]), outs, world)
    ref = _emulate("car", fields, f2d, (CAR["dz"], CAR["dy"], CAR["dx"]), outs, 1)
    for k in outs:
        assert_same(got[k], ref[k], f"car {k} world={world}")
    for k in ("zeta_r", "zeta_t", "div", "wind_speed"):      # and the reference itself
        assert_same(got[k], g[f"o_car_{k}"], f"car {k} vs reference")


@pytest.mark.parametrize("world", [2, 3])
def test_car_d_level_sharded_split_launches(golden_grids, world):
    """Stage B issued per level sub-range (mt_grid_t.sub_lo / sub_hi), as the overlapped run does."""
    import torch
    g = golden_grids
    fields = {k[len("g_car_"):]: v for k, v in g.items() if k.startswith("g_car_") and v.ndim == 3}
    f2d = torch.from_numpy(g["g_car_f"]).cuda()
    outs = [k for k in CAR_OUT if k not in ("Th",)]
    got = _emulate("car", fields, f2d, (CAR["dz"], CAR["dy"], CAR["dx"]), outs, world, split=True)
    ref = _emulate("car", fields, f2d, (CAR["dz"], CAR["dy"], CAR["dx"]), outs, 1)
    for k in outs:
        assert_same(got[k], ref[k], f"car {k} world={world} split")


@pytest.mark.parametrize("world", [2, 3])
def test_cyl_d_level_sharded(golden_grids, world):
    import torch
    g = golden_grids
    fields = {k[len("g_cyl_"):]: v for k, v in g.items() if k.startswith("g_cyl_") and v.ndim == 3}
    f2d = torch.from_numpy(g["g_cyl_f"]).cuda()
    r = np.arange(0, CYL["Nr"] * float(CYL["dr"]), float(CYL["dr"]))
    th = np.arange(0, CYL["Ntheta"] * float(CYL["dtheta"]), float(CYL["dtheta"]))
    tables = dict(r=torch.from_numpy(r).cuda(), sin_t=torch.from_numpy(np.sin(th)).cuda(),
                  cos_t=torch.from_numpy(np.cos(th)).cuda())
    outs = ["zeta_r", "zeta_t", "zeta_z", "div", "vr", "vt", "shear_deform", "kinetic_energy"]
    for split in (False, True):
        got = _emulate("cyl", fields, f2d, (CYL["dz"], CYL["dtheta"], CYL["dr"]), outs, world, tables, split=split)
        for k in outs:
            assert_same(got[k], g[f"o_cyl_{k}"], f"cyl {k} world={world} split={split}")


def _random_fields(nz, nt, nr, seed=11):
    rng = np.random.default_rng(seed)
    shp = (nz, nt, nr)
    zz = (np.arange(nz) * 250.0 + 100.0).reshape(-1, 1, 1)
    f = {k: 10 * rng.standard_normal(shp) for k in ("u", "v", "w")}
    f["p"] = 1e5 * np.exp(-zz / 8000.0) + rng.standard_normal(shp)
    f["T"] = 300 - 6.5e-3 * zz + rng.standard_normal(shp)
    f["qv"] = 1e-3 + 1e-3 * np.abs(rng.standard_normal(shp))
    for k in ("qc", "qr", "qi"):
        f[k] = 1e-4 * np.abs(rng.standard_normal(shp))
    return f


@pytest.mark.parametrize("kind,dims", [("car", (12, 20, 40)), ("car", (9, 17, 70)), ("cyl", (10, 24, 36))])
def test_level_sharded_tma_path(kind, dims):
    """Even row lengths: stage B runs as the TMA plane pipeline (csrc/diag_tma.cu).  Shards of 2-4 ranks,
    whole-shard and split (interior / boundary) launches, against the unsharded run -- bit for bit."""
    import torch
    nz, nt, nr = dims
    fields = _random_fields(nz, nt, nr)
    f2d = torch.from_numpy(5e-5 + 1e-9 * np.arange(nt * nr, dtype=np.float64).reshape(nt, nr)).cuda()
    tables, spac = None, (250.0, 2000.0, 2000.0)
    if kind == "cyl":
        th = np.arange(nt) * (2 * np.pi / nt)
        tables = dict(r=torch.from_numpy(np.arange(nr) * 1000.0).cuda(), sin_t=torch.from_numpy(np.sin(th)).cuda(),
                      cos_t=torch.from_numpy(np.cos(th)).cuda())
        spac = (250.0, 2 * np.pi / nt, 1000.0)
    from meteotools_b200 import _lib
    outs = [k for k in (_lib.CYL_D_OUT if kind == "cyl" else _lib.CAR_D_OUT) if k != "Th"]
    ref = _emulate(kind, fields, f2d, spac, outs, 1, tables)
    for world in (2, 3, 4):
        if nz // world < 2:
            continue
        for split in (False, True):
            got = _emulate(kind, fields, f2d, spac, outs, world, tables, split=split)
            for k in outs:
                assert_same(got[k], ref[k], f"{kind} {k} world={world} split={split}")


@pytest.mark.parametrize("kind,dims", [("car", (6, 24, 64)), ("cyl", (6, 24, 64))])
def test_stencil_tma_special_values_match_the_ieee_kernels(kind, dims, monkeypatch):
    """The TMA stencil kernel divides through mt::FDiv (reciprocal + exact correction, special operands selected
    branch-free); the register-marching kernels use IEEE divisions.  Zero fields, negative zeros, NaN blocks
    (terrain), +-inf points and the r = 0 column must come out BIT for bit the same, the sign of zero included."""
    import torch
    nz, nt, nr = dims
    fields = _random_fields(nz, nt, nr, seed=21)
    fields["w"][:] = 0.0                       # an identically zero field: every quotient of it is +-0
    fields["u"][1, 3:9, 5:20] = -0.0
    fields["u"][2, 4, :] = 0.0
    fields["v"][:2, 10:14, 30:50] = np.nan     # terrain-masked block
    fields["v"][4, 2, 7] = np.inf
    fields["u"][3, 20, 40] = -np.inf
    fields["T"][1, 12, 12] = np.nan
    f2d = torch.from_numpy(5e-5 + 1e-9 * np.arange(nt * nr, dtype=np.float64).reshape(nt, nr)).cuda()
    tables, spac = None, (250.0, 2000.0, 2000.0)
    if kind == "cyl":
        th = np.arange(nt) * (2 * np.pi / nt)
        tables = dict(r=torch.from_numpy(np.arange(nr) * 1000.0).cuda(), sin_t=torch.from_numpy(np.sin(th)).cuda(),
                      cos_t=torch.from_numpy(np.cos(th)).cuda())
        spac = (250.0, 2 * np.pi / nt, 1000.0)
    from meteotools_b200 import _lib
    outs = [k for k in (_lib.CYL_D_OUT if kind == "cyl" else _lib.CAR_D_OUT) if k != "Th"]
    monkeypatch.setenv("MT_DIAG_TMA", "1")
    got = _emulate(kind, fields, f2d, spac, outs, 1, tables)
    monkeypatch.setenv("MT_DIAG_TMA", "0")
    ref = _emulate(kind, fields, f2d, spac, outs, 1, tables)
    for k in outs:
        a, b = got[k], ref[k]
        if a.dtype != np.float64:
            assert np.array_equal(a, b), k
            continue
        nan = np.isnan(a) & np.isnan(b)
        same_bits = a.view(np.int64) == b.view(np.int64)        # +0 and -0 differ here
        assert (same_bits | nan).all(), f"{kind} {k}: {np.count_nonzero(~(same_bits | nan))} points differ in bits"
