"""Guard rails of the C ABI: a caller that breaks a precondition gets MT_EINVAL, never a lost CUDA context."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_setdata_tma_bad_items_are_reported_and_the_context_survives():
    """a misaligned tile origin used to end in cudaErrorIllegalInstruction (sticky)."""
    import torch
    from meteotools_b200 import _device as D, _lib
    from meteotools_b200.grids import cylindrical_grid
    from meteotools_b200.pipeline import CylStepPipeline
    lib = _lib.load()
    rng = np.random.default_rng(4)
    nz, ny, nx, dx = 10, 64, 64, 2000.0
    st = dict(Nr=20, Ntheta=24, Nz=8, dr=2000, dtheta=2 * np.pi / 24, dz=500, z_start=100, r_start=0,
              theta_start=0, vertical_coord_type="z", dt=300)
    cyl = cylindrical_grid(st)
    cyl.set_horizontal_location(63000.0, 63000.0)
    z3d = torch.from_numpy(np.broadcast_to((np.arange(nz) * 600.0).reshape(-1, 1, 1), (nz, ny, nx)).copy()).cuda()
    fld = torch.from_numpy(rng.standard_normal((nz, ny, nx))).cuda()
    pipe = CylStepPipeline(cyl, (nz, ny, nx), dx, dx, ["u"], names2d=(), hgt=False, td_outputs=None, resident=True)
    assert pipe.mode == "tma"
    pipe.bind_device(z3d, {"u": fld}, {}, direction=0)
    pipe.run()
    torch.cuda.synchronize()
    good = pipe.out3d["u"].clone()
    assert pipe.f_nitems >= 2

    def launch(items):
        g, s = pipe.g, D.stream()
        sny, snx, oy, ox = pipe.region
        return lib.mt_setdata_tma(pipe._src_ptrs, 1, D.ptr(pipe._vert), pipe.mt_dtype, pipe.nz, sny, snx, snx, oy, ox,
                                  D.ptr(pipe.levels), g.Nz, pipe.direction, D.ptr(pipe.plan[0]), D.ptr(pipe.plan[1]),
                                  pipe.n, D.ptr(pipe.f_order), D.ptr(items), pipe.f_nitems, pipe.f_nflag, None,
                                  pipe._out_ptrs, g.Nz, pipe.flags, D.ptr(pipe.f_counter), s)

    for corrupt in ("misaligned origin", "count too large", "start beyond order"):
        bad = pipe.f_items.clone()
        if corrupt == "misaligned origin":
            bad[0, 0] += 1
        elif corrupt == "count too large":
            bad[1, 3] = 500
        else:
            bad[1, 2] = 2 ** 30
        assert launch(bad) == 0                                   # asynchronous: the launch itself succeeds
        rc = lib.mt_setdata_tma_status(D.ptr(pipe.f_counter), D.stream())
        assert rc == _lib.MT_EINVAL if hasattr(_lib, "MT_EINVAL") else rc == -1, corrupt
        assert b"items" in lib.mt_last_error()
        assert lib.mt_setdata_tma_status(D.ptr(pipe.f_counter), D.stream()) == 0   # the record was cleared
        torch.cuda.synchronize()                                  # the context is alive ...
        pipe.run()                                                # ... and a correct call still works
        torch.cuda.synchronize()
        now = pipe.out3d["u"]
        assert bool(((now == good) | (torch.isnan(now) & torch.isnan(good))).all()), corrupt
