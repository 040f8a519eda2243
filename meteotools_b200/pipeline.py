"""Persistent per-time-step pipeline and the multi-GPU driver.

``CylStepPipeline`` is the production form of "set_data + diagnostics for one WRF time step":
everything that does not depend on the step (target plan, bounding box, level table, workspace,
output buffers) is prepared once; per step only ``upload`` (box-only strided DMA from the host
arrays), ``run`` (3-4 C calls: terrain index, mt_setdata, mt_cyl_td and/or mt_cyl_d) and
``download`` remain.  Results are identical to ``cylindrical_grid.set_data`` + ``calc_*``.

Multi-GPU: WRF time steps are independent, so ``shard_steps`` deals them round-robin to the
ranks of a ``torch.distributed`` job (one process per GPU, no data-path collective); vertical
levels couple only through FD2(., dz, axis=0), for which ``halo_exchange_levels`` sends one
boundary level to each neighbour (NCCL send/recv; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import _device as D
from . import _lib


def shard_steps(n_steps: int, rank: int, world: int):
    """Time steps of this rank: round-robin, so that ranks stay balanced for any n_steps."""
    return list(range(rank, n_steps, world))


def level_shard(nz: int, rank: int, world: int):
    """Contiguous level block [lo, hi) of `rank` (sizes differ by at most one level) and the
    halo levels it needs for a centred z-difference: (lo, hi, halo_lo, halo_hi)."""
    base, rem = divmod(nz, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi, (1 if rank > 0 else 0), (1 if rank < world - 1 else 0)


def halo_exchange_levels(fields, halo_lo, halo_hi, rank, world, dist=None):
    """One-level halo exchange along axis 0 of level-sharded (nz_local, ny, nx) C-order tensors
    (owned levels are [halo_lo, nz_local-halo_hi)).  Batched isend/irecv pairs with rank +- 1;
    works on CUDA tensors with NCCL and on CPU tensors with gloo."""
    import torch.distributed as dist_mod
    dist = dist or dist_mod
    ops = []
    for f in fields:
        n = f.shape[0]
        if halo_lo:   # talk to rank-1: send my first owned level, receive its last owned level
            ops.append(dist.P2POp(dist.isend, f[halo_lo], rank - 1))
            ops.append(dist.P2POp(dist.irecv, f[0], rank - 1))
        if halo_hi:   # talk to rank+1
            ops.append(dist.P2POp(dist.isend, f[n - 1 - halo_hi], rank + 1))
            ops.append(dist.P2POp(dist.irecv, f[n - 1], rank + 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class CylStepPipeline:
    """cart -> cyl ingest (+ terrain mask) and fused diagnostics of one time step on one GPU."""

    def __init__(self, grid, src_shape, dx, dy, names3d, names2d=("f",), hgt=True, dtype=np.float64,
                 td_outputs=tuple(_lib.TD_OUT), dyn_outputs=None, inclusive=False, resident=False):
        t = D.require_cuda()
        self.g, self.lib = grid, _lib.load()
        self.nz, self.ny, self.nx = src_shape
        self.dx, self.dy = float(dx), float(dy)
        self.names3d, self.names2d, self.has_hgt = list(names3d), list(names2d), bool(hgt)
        assert len(self.names3d) <= _lib.MT_MAX_VARS
        self.np_dtype = np.dtype(dtype)
        self.mt_dtype = _lib.MT_F32 if self.np_dtype == np.float32 else _lib.MT_F64
        self.t_dtype = t.float32 if self.np_dtype == np.float32 else t.float64
        self.flags = (_lib.MT_IL_INCLUSIVE if inclusive else 0) | \
                     (_lib.MT_IL_ROUND_F32 if self.np_dtype == np.float32 else 0)
        self.box = grid._target_box(self.nx, self.ny, self.dx, self.dy)
        y0, y1, x0, x1 = self.box
        self.resident = resident
        # host sources are staged box-only with rows padded to D.box_pitch (16-byte aligned rows)
        self.region = (self.ny, self.nx, 0, 0) if resident else (y1 - y0, D.box_pitch(x1 - x0), y0, x0)
        g = grid
        self.levels = D.to_device(g.z)
        self.order = 1 if np.all(np.diff(g.z) > 0) else (-1 if np.all(np.diff(g.z) < 0) else 0)
        self.xd, self.yd = g._targets_dev()
        self.n = self.xd.numel()
        self.plan = g._plan_for(self.dx, self.dy, self.nx, self.ny)
        self.wsb = self.lib.mt_setdata_workspace_bytes(len(self.names3d), g.Nz, y0, y1, x0, x1)
        self.ws = D.empty(((self.wsb + 7) // 8,))
        self.out3d = {k: g._new() for k in self.names3d}
        self.out2d = {k: g._new(2) for k in self.names2d + (["hgt"] if hgt else [])}
        self.hgt_idx = D.empty((g.Ntheta, g.Nr), t.int32) if hgt else None
        self.td_names = list(td_outputs or [])
        for dep in ("Tc", "Th"):
            if "Th_e" in self.td_names and dep not in self.td_names:
                self.td_names.append(dep)
        self.td_out = {k: g._new() for k in self.td_names}
        self.dyn_names = list(dyn_outputs or [])
        self.dyn_out = {k: g._new(dtype=t.uint8 if k == "strain_region" else None) for k in self.dyn_names}
        self.scratch = D.empty((16 + self.n * g.Nz,))
        self.direction = -1
        import os
        mode = os.environ.get("MT_SETDATA", "tma")   # tma (default) | split | fused
        fits = self.nz <= 254
        self.mode = "split"
        if fits and mode == "tma" and self.lib.mt_setdata_tma_supported(self.mt_dtype, self.nz, g.Nz,
                                                                          self.region[1]):
            self.mode = "tma"
            tw_, th_, al_, mt_ = _lib.tma_tile(self.mt_dtype)
            self.f_order, self.f_items, self.f_nitems, self.f_nflag, _ = \
                g._fused_items_for(self.plan, tw_, th_, max_targets=mt_, spatial=True, align=al_,
                                   org_x=self.region[3])
        elif fits and g.Nz % 2 == 0 and mode == "fused":
            self.mode = "fused"
            self.tile_h = 2 if os.environ.get("MT_SETDATA_TILE") == "16x2" else 4
            self.f_order, self.f_items, self.f_nitems, self.f_nflag, _ = \
                g._fused_items_for(self.plan, 16, self.tile_h)
        # the work-item counter is per pipeline: pipelines of one grid may run on different streams
        self.f_counter = t.zeros((2,), dtype=t.int32, device="cuda")   # [work-item counter, status word]
        self._items_checked = False
        # the 2-D part of the step (bilinear of the 2-D fields, terrain index + fix-up) is one launch
        self.f_ticket = t.zeros((1,), dtype=t.int32, device="cuda")
        self._keys2d = list(self.out2d)
        self._out2d_ptrs = D.ptr_array([self.out2d[k] for k in self._keys2d]) if self._keys2d else None
        if not resident:   # persistent device staging for the uploaded boxes
            shp = (self.nz, y1 - y0, self.region[1])
            self.stage3d = {k: t.empty(shp, dtype=self.t_dtype, device="cuda") for k in self.names3d}
            self.stage_vert = t.empty(shp, dtype=self.t_dtype, device="cuda")
            self.stage2d = {k: t.empty((self.ny, self.nx), dtype=t.float64, device="cuda")
                            for k in self.out2d}
        self._build_structs()

    # ---- per-step input binding -----------------------------------------------------------
    def h2d_bytes(self):
        y0, y1, x0, x1 = self.box
        return (len(self.names3d) + 1) * self.nz * (y1 - y0) * (x1 - x0) * self.np_dtype.itemsize + \
            len(self.out2d) * self.ny * self.nx * 8

    def upload(self, vert, vars3d, vars2d):
        """Host arrays -> device staging: box-only strided DMA for the 3-D fields."""
        y0, y1, x0, x1 = self.box
        self.direction = 1 if float(vert[0, 0, 0]) > float(vert[self.nz - 1, 0, 0]) else 0
        D.upload_box(vert, y0, y1, x0, x1, out=self.stage_vert)
        for k in self.names3d:
            D.upload_box(vars3d[k], y0, y1, x0, x1, out=self.stage3d[k])
        t = D.torch()
        for k in self.out2d:
            self.stage2d[k].copy_(t.from_numpy(np.ascontiguousarray(vars2d[k], dtype=np.float64)),
                                  non_blocking=True)
        self._bind(self.stage_vert, self.stage3d, self.stage2d)

    def bind_device(self, vert, vars3d, vars2d, direction):
        """Use resident full-size device arrays (no copy)."""
        self.direction = int(direction)
        self._bind(vert, vars3d, vars2d)

    def _bind(self, vert, vars3d, vars2d):
        self._vert = vert
        self._src3d = [vars3d[k] for k in self.names3d]
        self._src2d = vars2d
        self._src_ptrs = D.ptr_array(self._src3d)
        self._src2d_ptrs = D.ptr_array([vars2d[k] for k in self._keys2d]) if self._keys2d else None

    def _build_structs(self):
        g = self.g
        self._out_ptrs = D.ptr_array([self.out3d[k] for k in self.names3d])
        if self.td_names:
            self._tin, self._tout = _lib.mt_td_in_t(), _lib.mt_td_out_t()
            for k in _lib.TD_IN:
                setattr(self._tin, k, self.out3d[k].data_ptr() if k in self.out3d else None)
            for k in self.td_names:
                setattr(self._tout, k, self.td_out[k].data_ptr())
        if self.dyn_names:
            self._grid = g._grid_struct()
            self._din, self._dout = _lib.mt_dyn_in_t(), _lib.mt_cyl_d_out_t()
            for k in ("u", "v", "w", "p", "T", "Th", "qv"):
                setattr(self._din, k, self.out3d[k].data_ptr() if k in self.out3d else None)
            self._din.f = self.out2d["f"].data_ptr() if "f" in self.out2d else None
            hyd = [k for k in _lib.HYDRO if k in self.out3d]
            for i in range(6):
                self._din.q[i] = self.out3d[hyd[i]].data_ptr() if i < len(hyd) else None
            for k in self.dyn_names:
                setattr(self._dout, k, self.dyn_out[k].data_ptr())

    # ---- the step ---------------------------------------------------------------------------
    def run(self):
        lib, g, s = self.lib, self.g, D.stream()
        y0, y1, x0, x1 = self.box
        sny, snx, oy, ox = self.region
        if 1 <= len(self._keys2d) <= 4:   # 2-D fields (bilinear only) + terrain index: one launch
            _lib.check(lib.mt_cyl_prelude(self._src2d_ptrs, self._out2d_ptrs, len(self._keys2d), self.ny, self.nx,
                                          D.ptr(self.plan[0]), D.ptr(self.plan[1]), g.Ntheta, g.Nr,
                                          self._keys2d.index("hgt") if self.has_hgt else -1, g.z_start, g.dz,
                                          D.ptr(self.hgt_idx), D.ptr(self.f_ticket), s), "mt_cyl_prelude")
        else:
            for k, o in self.out2d.items():
                _lib.check(lib.mt_interp2d_layers(D.ptr(self._src2d[k]), 1, self.ny, self.nx, 0, self.nx, 1,
                                                  self.dx, self.dy, D.ptr(self.xd), D.ptr(self.yd), self.n,
                                                  D.ptr(o), 0, 1, s), "mt_interp2d_layers")
            if self.has_hgt:
                _lib.check(lib.mt_terrain_index_cyl(D.ptr(self.out2d["hgt"]), g.Ntheta, g.Nr, g.z_start, g.dz,
                                                    D.ptr(self.hgt_idx), s), "mt_terrain_index_cyl")
        if self.mode == "tma":
            _lib.check(lib.mt_setdata_tma(self._src_ptrs, len(self.names3d), D.ptr(self._vert), self.mt_dtype,
                                          self.nz, sny, snx, snx, oy, ox, D.ptr(self.levels), g.Nz, self.direction,
                                          D.ptr(self.plan[0]), D.ptr(self.plan[1]), self.n, D.ptr(self.f_order),
                                          D.ptr(self.f_items), self.f_nitems, self.f_nflag,
                                          D.ptr(self.hgt_idx), self._out_ptrs, g.Nz, self.flags,
                                          D.ptr(self.f_counter), s), "mt_setdata_tma")
            if not self._items_checked:   # once per plan: the kernel's verdict on the work items it was given
                _lib.check(lib.mt_setdata_tma_status(D.ptr(self.f_counter), s), "mt_setdata_tma_status")
                self._items_checked = True
        elif self.mode == "fused":
            _lib.check(lib.mt_setdata_fused(self._src_ptrs, len(self.names3d), D.ptr(self._vert), self.mt_dtype,
                                            self.nz, sny, snx, oy, ox, D.ptr(self.levels), g.Nz, self.direction,
                                            D.ptr(self.plan[0]), D.ptr(self.plan[1]), self.n, D.ptr(self.f_order),
                                            D.ptr(self.f_items), self.f_nitems, 16, self.tile_h, self.f_nflag,
                                            D.ptr(self.hgt_idx), self._out_ptrs, g.Nz, self.flags,
                                            D.ptr(self.f_counter), s), "mt_setdata_fused")
        else:
            _lib.check(lib.mt_setdata(self._src_ptrs, len(self.names3d), D.ptr(self._vert), self.mt_dtype, self.nz,
                                      sny, snx, oy, ox, self.ny, self.nx, self.dx, self.dy, y0, y1, x0, x1,
                                      D.ptr(self.levels), g.Nz, self.order, self.direction, D.ptr(self.xd),
                                      D.ptr(self.yd), D.ptr(self.plan[0]), D.ptr(self.plan[1]), self.n,
                                      D.ptr(self.hgt_idx), self._out_ptrs, 1, g.Nz, self.flags, D.ptr(self.ws),
                                      self.wsb, s), "mt_setdata")
        if self.td_names:
            _lib.check(lib.mt_cyl_td(self._tin, self._tout, self.n * g.Nz, D.ptr(self.scratch), s), "mt_cyl_td")
        if self.dyn_names:
            _lib.check(lib.mt_cyl_d(self._grid, self._din, self._dout, _lib.MT_STAGE_ALL, s), "mt_cyl_d")

    def download(self, names=None):
        """Results as NumPy (Nz, Ntheta, Nr) views of pinned host copies."""
        pool = {**self.out3d, **self.td_out, **self.dyn_out}
        names = list(names) if names is not None else list(self.td_out) + list(self.dyn_out)
        t = D.torch()
        hosts = {}
        for k in names:
            h = t.empty(pool[k].shape, dtype=pool[k].dtype, pin_memory=True)
            h.copy_(pool[k], non_blocking=True)
            hosts[k] = h
        t.cuda.current_stream().synchronize()
        return {k: h.numpy().transpose(2, 0, 1) for k, h in hosts.items()}

    def d2h_bytes(self, names=None):
        names = list(names) if names is not None else list(self.td_out) + list(self.dyn_out)
        return len(names) * self.n * self.g.Nz * 8

    def store_into_grid(self):
        """Attach the step's device buffers to the grid object as fields (no copy)."""
        for pool, nd in ((self.out3d, 3), (self.td_out, 3), (self.dyn_out, 3), (self.out2d, 2)):
            for k, v in pool.items():
                self.g._put(k, v, nd)


class HaloComm:
    """The library's own NCCL communicator for the level-shard halo exchange (csrc/halo.cu,
    ``mt_halo_exchange``): rank 0 creates the NCCL unique id, ``torch.distributed`` carries its 128 bytes
    to the other ranks, every rank joins.  The library binds to the NCCL copy torch already loaded.

        comm = HaloComm(rank, world)            # torch.distributed must be initialised (any backend)
        comm.exchange([v, w], hlo, hhi, stream) # asynchronous on `stream`
    """

    def __init__(self, rank, world, dist=None):
        import ctypes as ct
        import torch.distributed as dist_mod
        t = D.require_cuda()
        dist = dist or dist_mod
        self.lib, self.rank, self.world = _lib.load(), rank, world
        path = None
        try:   # only needed when no NCCL is resident yet (torch loads it with torch.cuda / distributed)
            import os
            import nvidia.nccl as _n
            cand = os.path.join(list(_n.__path__)[0], "lib", "libnccl.so.2")
            path = cand.encode() if os.path.exists(cand) else None
        except Exception:
            pass
        _lib.check(self.lib.mt_nccl_load(path), "mt_nccl_load")
        idbuf = ct.create_string_buffer(128)
        if rank == 0:
            _lib.check(self.lib.mt_nccl_unique_id(idbuf), "mt_nccl_unique_id")
        backend = dist.get_backend()
        carrier = t.frombuffer(bytearray(idbuf.raw), dtype=t.uint8).clone()
        if backend == "nccl":
            carrier = carrier.cuda()
        dist.broadcast(carrier, src=0)
        raw = bytes(carrier.cpu().numpy().tobytes())
        self.comm = ct.c_void_p()
        _lib.check(self.lib.mt_nccl_comm_init(ct.byref(self.comm), world, rank, raw), "mt_nccl_comm_init")

    def exchange(self, tensors, halo_lo, halo_hi, stream=None):
        """One halo level of every (nz_local, ny, nx) C-order tensor, asynchronous on `stream`."""
        if not (halo_lo or halo_hi) or not tensors:
            return
        first = tensors[0]
        plane = int(first.shape[1]) * int(first.shape[2])
        for x in tensors:
            assert x.is_contiguous() and x.shape == first.shape and x.dtype == first.dtype
        s = stream.cuda_stream if stream is not None else D.stream()
        _lib.check(self.lib.mt_halo_exchange(self.comm, D.ptr_array(tensors), len(tensors), plane,
                                             int(first.shape[0]), int(halo_lo), int(halo_hi),
                                             self.rank - 1 if halo_lo else -1, self.rank + 1 if halo_hi else -1, s),
                   "mt_halo_exchange")

    def close(self):
        if self.comm:
            _lib.check(self.lib.mt_nccl_comm_destroy(self.comm), "mt_nccl_comm_destroy")
            self.comm = None


class PeerHalo:
    """The level-shard halo exchange over NVLink / NVSwitch PEER MEMORY (csrc/peer.cu): the planes travel by
    the copy engines straight into a mailbox in the neighbour's memory, the processes order themselves with
    stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32) -- no kernel, no SM, no host
    synchronisation.  Same call as ``HaloComm.exchange``; all ranks must sit on one node with peer access.

        halo = PeerHalo(rank, world, plane_elems=ny * nx, nfields=3)   # collective: IPC handles are exchanged
        halo.exchange([w, v, th_rho], hlo, hhi, stream)
    """
    HEADER = 256   # bytes: [0] data_seq, [64] ack_seq

    def __init__(self, rank, world, plane_elems, nfields, dist=None):
        import ctypes as ct
        import torch.distributed as dist_mod
        D.require_cuda()
        dist = dist or dist_mod
        self.lib, self.rank, self.world = _lib.load(), rank, world
        if not self.lib.mt_peer_supported():
            raise RuntimeError("cuStreamWaitValue32 / cuStreamWriteValue32 are not available in this driver")
        self.plane_bytes, self.nfields, self.seq = int(plane_elems) * 8, int(nfields), 0
        nbytes = self.HEADER + self.nfields * self.plane_bytes
        self.box, handles = {}, {}
        for d, nb in (("lo", rank - 1), ("hi", rank + 1)):
            if 0 <= nb < world:
                ptr, h = ct.c_void_p(), ct.create_string_buffer(64)
                _lib.check(self.lib.mt_peer_box_create(nbytes, ct.byref(ptr), h), "mt_peer_box_create")
                self.box[d], handles[d] = ptr.value, h.raw
        everyone = [None] * world
        dist.all_gather_object(everyone, handles)
        self.peer = {}
        for d, nb, facing in (("lo", rank - 1, "hi"), ("hi", rank + 1, "lo")):
            if 0 <= nb < world:     # the neighbour's box that faces me
                ptr = ct.c_void_p()
                _lib.check(self.lib.mt_peer_box_open(everyone[nb][facing], ct.byref(ptr)), "mt_peer_box_open")
                self.peer[d] = ptr.value
        dist.barrier()

    def exchange(self, tensors, halo_lo, halo_hi, stream=None):
        """One halo level of every (nz_local, ny, nx) C-order tensor, asynchronous on `stream`."""
        if not (halo_lo or halo_hi) or not tensors:
            return
        lib, pb = self.lib, self.plane_bytes
        assert len(tensors) <= self.nfields and all(x.is_contiguous() and x[0].numel() * 8 == pb for x in tensors)
        s = stream.cuda_stream if stream is not None else D.stream()
        self.seq += 1
        seq, nzl = self.seq, int(tensors[0].shape[0])
        sides = ([("lo", halo_lo, 0)] if halo_lo else []) + ([("hi", nzl - 1 - halo_hi, nzl - 1)] if halo_hi else [])
        for d, own_level, _ in sides:      # my boundary planes -> the neighbour's box
            _lib.check(lib.mt_stream_wait_geq32(self.box[d] + 64, seq - 1, s), "wait ack")
            for k, x in enumerate(tensors):
                _lib.check(lib.mt_peer_copy(self.peer[d] + self.HEADER + k * pb, x.data_ptr() + own_level * pb, pb, s),
                           "mt_peer_copy")
            _lib.check(lib.mt_stream_write32(self.peer[d], seq, s), "write data_seq")
        for d, _, halo_level in sides:     # the neighbour's planes -> my halo levels
            _lib.check(lib.mt_stream_wait_geq32(self.box[d], seq, s), "wait data")
            for k, x in enumerate(tensors):
                _lib.check(lib.mt_peer_copy(x.data_ptr() + halo_level * pb, self.box[d] + self.HEADER + k * pb, pb, s),
                           "mt_peer_copy")
            _lib.check(lib.mt_stream_write32(self.peer[d] + 64, seq, s), "write ack_seq")

    def close(self):
        for ptr in self.peer.values():
            self.lib.mt_peer_box_close(ptr)
        for ptr in self.box.values():
            self.lib.mt_peer_box_destroy(ptr)
        self.peer, self.box = {}, {}


class LevelShardDyn:
    """cyl_d / car_d of ONE time step on ONE level shard.

    kind      "car" or "cyl"
    local     dict name -> CUDA tensor (nz_own, nt, nr) C-order ("ZTR") with the shard's OWNED levels
              of u, v, w, p, T, qv (, Th, qc, qr, qi, qs, qg, qh)
    f2d       (nt, nr) Coriolis tensor or None
    spacings  (dz, dt, dr)  i.e. (dz, dy, dx) or (dz, dtheta, dr)
    shard     (lo, hi, halo_lo, halo_hi) from ``level_shard``
    tables    cyl only: dict(r=, sin_t=, cos_t=) CUDA tensors

    Only FD2(., dz, axis=0) couples levels: one halo plane of the z-differentiated INPUTS
    (``halo_inputs``: v, w -- and u on a cylindrical grid, whose rotation mixes u and v) and of the derived
    theta_rho (``halo_derived``) is needed from each neighbour; the one-sided formulas are used only at
    the global first / last level.  ``run(comm)`` overlaps the exchange with the work that does not need it:

        side stream:  exchange(halo_inputs) ........ [wait stage A] exchange(theta_rho)
        main stream:  stage A (thermodynamic point-wise fields) -> stage B on the levels that need no halo
                      -> [wait side] stage B on the one or two boundary levels

    ``stage_a`` / ``stage_b`` / ``halo_fields`` remain for callers that exchange by other means.
    """

    def __init__(self, kind, local, f2d, spacings, shard, nz_global, outputs, tables=None):
        t = D.require_cuda()
        self.t = t
        self.lib, self.kind, self.outputs = _lib.load(), kind, list(outputs)
        lo, hi, hlo, hhi = shard
        self.hlo, self.hhi, self.nz_own = hlo, hhi, hi - lo
        any_field = next(iter(local.values()))
        nt, nr = int(any_field.shape[1]), int(any_field.shape[2])
        nzl = self.nz_own + hlo + hhi

        def padded(x):   # owned levels placed between the halo slots
            buf = t.empty((nzl, nt, nr), dtype=t.float64, device="cuda")
            buf[hlo:hlo + self.nz_own].copy_(x)
            return buf
        self.fields = {k: padded(v) for k, v in local.items()}
        self._keep = (f2d, tables)
        g = _lib.mt_grid_t()
        g.nz, g.nt, g.nr, g.layout = nzl, nt, nr, _lib.MT_LAYOUT_ZTR
        g.dz, g.dt, g.dr = (float(s) for s in spacings)
        if kind == "cyl":
            g.r, g.sin_t, g.cos_t = (tables[k].data_ptr() for k in ("r", "sin_t", "cos_t"))
        g.halo_lo, g.halo_hi = hlo, hhi
        g.is_bottom, g.is_top = int(lo == 0), int(hi == nz_global)
        din = _lib.mt_dyn_in_t()
        for k in ("u", "v", "w", "p", "T", "Th", "qv"):
            setattr(din, k, self.fields[k].data_ptr() if k in self.fields else None)
        din.f = f2d.data_ptr() if f2d is not None else None
        hyd = [k for k in _lib.HYDRO if k in self.fields]
        for i in range(6):
            din.q[i] = self.fields[hyd[i]].data_ptr() if i < len(hyd) else None
        names = _lib.CYL_D_OUT if kind == "cyl" else _lib.CAR_D_OUT
        dout = (_lib.mt_cyl_d_out_t if kind == "cyl" else _lib.mt_car_d_out_t)()
        self.outs = {}
        stage_a_needed = ["Th_rho", "rho"] + (["vr", "vt"] if kind == "cyl" else [])
        for k in list(dict.fromkeys(self.outputs + stage_a_needed)):
            if k not in names:
                raise ValueError(f"unknown output {k!r}")
            self.outs[k] = t.empty((nzl, nt, nr), dtype=t.uint8 if k == "strain_region" else t.float64,
                                   device="cuda")
            setattr(dout, k, self.outs[k].data_ptr())
        self._g, self._din, self._dout = g, din, dout
        self._call = self.lib.mt_cyl_d if kind == "cyl" else self.lib.mt_car_d
        self._plane, self._hyd, self._f2d = nt * nr, hyd, f2d
        self._views = {}
        # the z-differentiated fields: inputs first (they can travel at once), theta_rho after stage A
        self.halo_inputs = [self.fields["w"], self.fields["v"]] + ([self.fields["u"]] if kind == "cyl" else [])
        self.halo_derived = [self.outs["Th_rho"]]
        self.halo_fields = self.halo_inputs + self.halo_derived
        self._side = None

    def _stage(self, stages, sub=(0, 0)):
        self._g.sub_lo, self._g.sub_hi = sub
        _lib.check(self._call(self._g, self._din, self._dout, stages, D.stream()), f"stages {stages}")
        self._g.sub_lo, self._g.sub_hi = 0, 0

    def stage_a(self):
        """Thermodynamic point-wise fields of all local levels (Tv, rho, density factor, theta, theta_rho)."""
        self._stage(_lib.MT_STAGE_A | _lib.MT_STAGE_WIND_IN_B)

    def stage_a_levels(self, level0, nlev):
        """The same for the local levels [level0, level0 + nlev) only: the point-wise stage has no neighbours, so
        a level range is the same call on pointers moved to its first level."""
        if nlev <= 0:
            return
        key = (level0, nlev)
        if key not in self._views:
            off = level0 * self._plane * 8
            g = _lib.mt_grid_t()
            for name, _ in _lib.mt_grid_t._fields_:
                setattr(g, name, getattr(self._g, name))
            g.nz, g.halo_lo, g.halo_hi, g.is_bottom, g.is_top, g.sub_lo, g.sub_hi = nlev, 0, 0, 1, 1, 0, 0
            din = _lib.mt_dyn_in_t()
            for k in ("u", "v", "w", "p", "T", "Th", "qv"):
                setattr(din, k, self.fields[k].data_ptr() + off if k in self.fields else None)
            din.f = self._f2d.data_ptr() if self._f2d is not None else None
            for i in range(6):
                din.q[i] = self.fields[self._hyd[i]].data_ptr() + off if i < len(self._hyd) else None
            dout = (_lib.mt_cyl_d_out_t if self.kind == "cyl" else _lib.mt_car_d_out_t)()
            for k in ("Tv", "rho", "density_factor", "Th", "Th_rho"):
                if k in self.outs:
                    setattr(dout, k, self.outs[k].data_ptr() + off)
            self._views[key] = (g, din, dout)
        g, din, dout = self._views[key]
        _lib.check(self._call(g, din, dout, _lib.MT_STAGE_A | _lib.MT_STAGE_WIND_IN_B, D.stream()), "stage A (levels)")

    def stage_b(self, sub=(0, 0)):
        """Stencil outputs + the wind-derived point-wise outputs of the owned levels [sub[0], sub[1])
        (indices into the owned block; (0, 0) = all of it)."""
        self._stage(_lib.MT_STAGE_B | _lib.MT_STAGE_WIND_IN_B, sub)

    def interior(self):
        """Owned levels whose z-stencil stays inside the shard: they need no halo plane."""
        return (1 if self.hlo else 0, self.nz_own - (1 if self.hhi else 0))

    def stage_b_split(self):
        """stage_b() as run() issues it: interior levels first, then the boundary levels."""
        a, b = self.interior()
        if b > a:
            self.stage_b((a, b))
        if self.hlo:
            self.stage_b((0, 1))
        if self.hhi and (self.nz_own > 1 or not self.hlo):
            self.stage_b((self.nz_own - 1, self.nz_own))

    def run(self, comm=None, marks=None):
        """The whole set with the halo exchange overlapped (see the class docstring).  `comm`: a HaloComm,
        or None for a single shard.  `marks`: a dict that receives CUDA events at the phase boundaries
        (start, inputs_exchanged, stage_a, theta_rho_exchanged, interior, end) for timeline measurements."""
        t = self.t

        def mark(name, stream):
            if marks is not None:
                ev = t.cuda.Event(enable_timing=True)
                ev.record(stream)
                marks[name] = ev
        if comm is None or not (self.hlo or self.hhi):
            self.stage_a()
            self.stage_b()
            return
        if isinstance(comm, PeerHalo):
            return self._run_peer(comm, mark)
        main = t.cuda.current_stream()
        if self._side is None:
            # HIGH priority: the stencil and point-wise kernels are persistent and fill every SM; when the
            # exchange and a compute kernel become runnable together, NCCL's few CTAs must be placed first,
            # or the planes only start to move when the compute kernel has finished (measured: no overlap)
            self._side = t.cuda.Stream(priority=-1)
        side = self._side
        mark("start", main)
        side.wait_stream(main)                      # the inputs are ready
        comm.exchange(self.halo_inputs, self.hlo, self.hhi, side)
        mark("inputs_exchanged", side)
        self.stage_a()
        mark("stage_a", main)
        ev_a = t.cuda.Event()
        ev_a.record(main)
        side.wait_event(ev_a)                       # theta_rho of the boundary levels exists
        comm.exchange(self.halo_derived, self.hlo, self.hhi, side)
        ev_h = t.cuda.Event()
        ev_h.record(side)
        mark("theta_rho_exchanged", side)
        a, b = self.interior()
        if b > a:
            self.stage_b((a, b))                    # overlaps the planes in flight
        mark("interior", main)
        main.wait_event(ev_h)
        if self.hlo:
            self.stage_b((0, 1))
        if self.hhi and (self.nz_own > 1 or not self.hlo):
            self.stage_b((self.nz_own - 1, self.nz_own))
        mark("end", main)

    def _run_peer(self, halo, mark):
        """run() over peer memory: ONE exchange of all three planes per side, started as soon as theta_rho of
        the boundary levels exists; it needs no SM, so the point-wise stage of the other levels and the
        interior stencil levels run beside it undisturbed.

            side:  [theta_rho of the boundary levels] -> put planes -> wait neighbours' planes -> halo levels
            main:  point-wise stage of the other levels -> stencil interior -> [wait side] -> boundary levels
        """
        t = self.t
        main = t.cuda.current_stream()
        if self._side is None:
            self._side = t.cuda.Stream(priority=-1)
        side = self._side
        first, last = self.hlo, self.hlo + self.nz_own - 1          # local indices of the owned end levels
        send = sorted({first} if self.hlo else set()) + sorted(({last} if self.hhi else set()) - ({first} if self.hlo else set()))
        mark("start", main)
        for lvl in send:                                            # tiny launches: one level each
            self.stage_a_levels(lvl, 1)
        ev0 = t.cuda.Event()
        ev0.record(main)
        side.wait_event(ev0)
        halo.exchange(self.halo_fields, self.hlo, self.hhi, side)
        ev_h = t.cuda.Event()
        ev_h.record(side)
        mark("theta_rho_exchanged", side)
        lo = first + (1 if first in send else 0)
        hi = last - (1 if last in send else 0)
        self.stage_a_levels(lo, hi - lo + 1)                        # never touches the halo levels
        mark("stage_a", main)
        a, b = self.interior()
        if b > a:
            self.stage_b((a, b))
        mark("interior", main)
        main.wait_event(ev_h)
        if self.hlo:
            self.stage_b((0, 1))
        if self.hhi and (self.nz_own > 1 or not self.hlo):
            self.stage_b((self.nz_own - 1, self.nz_own))
        mark("end", main)

    def result(self):
        return {k: self.outs[k][self.hlo:self.hlo + self.nz_own] for k in self.outputs}


def dyn_level_sharded(kind, local, f2d, spacings, shard, nz_global, outputs, rank, world, tables=None,
                      dist=None, comm=None):
    """Distributed form of ``LevelShardDyn``.  With a ``HaloComm`` the exchange runs through the library's
    NCCL communicator, overlapped with the interior levels; without one (CPU tests with gloo, or a caller
    that has no NCCL) the planes travel through ``torch.distributed`` point-to-point calls."""
    job = LevelShardDyn(kind, local, f2d, spacings, shard, nz_global, outputs, tables)
    if comm is not None or world == 1:
        job.run(comm if world > 1 else None)
        return job.result()
    job.stage_a()
    D.torch().cuda.current_stream().synchronize()
    halo_exchange_levels(job.halo_fields, shard[2], shard[3], rank, world, dist)
    job.stage_b()
    return job.result()


class TimeStepStream:
    """Process a sequence of WRF time steps end to end (host arrays in, host arrays out) with the
    PCIe traffic of consecutive steps overlapped: ``slots`` ``CylStepPipeline`` slots, each on its own
    CUDA stream; while step k's results travel device->host, step k+1's source boxes already
    travel host->device (PCIe is full duplex) and its kernels run.  Results are identical to
    calling ``cylindrical_grid.set_data`` + ``calc_all_thermaldynamic`` per step.

        ts = TimeStepStream(grid, (nz, ny, nx), dx, dy, names3d, names2d=("f",), outputs=[...])
        for k, result in ts.process(steps):      # steps: iterable of dict(z3d=, vars3d=, f=, hgt=)
            ...                                  # result: dict name -> (Nz, Ntheta, Nr) ndarray

    ``out_dtype=np.float32``: the results are rounded to float32 on the device (as the reference's
    netCDF cache does when it stores them, grid_general.py:119-129) and travel as float32: half the
    download, which is what bounds the stream.  The default keeps the reference's float64 arrays.
    """

    def __init__(self, grid, src_shape, dx, dy, names3d, names2d=("f",), hgt=True, outputs=None,
                 dtype=np.float64, dyn_outputs=None, slots=3, out_dtype=np.float64):
        t = D.require_cuda()
        self.t = t
        self.outputs = list(outputs) if outputs is not None else list(_lib.TD_OUT)
        self.slots = []
        for _ in range(slots):
            pipe = CylStepPipeline(grid, src_shape, dx, dy, names3d, names2d=names2d, hgt=hgt, dtype=dtype,
                                   td_outputs=[o for o in self.outputs if o in _lib.TD_OUT] or None,
                                   dyn_outputs=dyn_outputs, resident=False)
            pool = {**pipe.out3d, **pipe.td_out, **pipe.dyn_out}
            self.out32 = np.dtype(out_dtype) == np.float32
            narrow = {}
            if self.out32:   # float64 results are narrowed on the device before the download
                narrow = {k: t.empty(pool[k].shape, dtype=t.float32, device="cuda") for k in self.outputs
                          if pool[k].dtype == t.float64}
            host = {k: t.empty(pool[k].shape, dtype=narrow[k].dtype if k in narrow else pool[k].dtype,
                               pin_memory=True) for k in self.outputs}
            self.slots.append(dict(pipe=pipe, pool=pool, host=host, narrow=narrow, stream=t.cuda.Stream(),
                                   event=None, key=None))

    def _collect(self, slot):
        slot["event"].synchronize()
        res = {k: h.numpy().transpose(2, 0, 1) for k, h in slot["host"].items()}
        key, slot["event"], slot["key"] = slot["key"], None, None
        return key, res

    def process(self, steps):
        """Generator over (step_key, results).  Result arrays are views of the slot's pinned
        buffers: they stay valid until ``len(slots)`` further steps have been issued."""
        i = 0
        for key, step in (steps.items() if isinstance(steps, dict) else enumerate(steps)):
            slot = self.slots[i % len(self.slots)]
            i += 1
            if slot["event"] is not None:
                yield self._collect(slot)
            with self.t.cuda.stream(slot["stream"]):
                v2d = {"f": step["f"]} if "f" in slot["pipe"].out2d else {}
                if slot["pipe"].has_hgt:
                    v2d["hgt"] = step["hgt"]
                for k in slot["pipe"].out2d:
                    if k not in v2d:
                        v2d[k] = step[k]
                slot["pipe"].upload(step["z3d"], step["vars3d"], v2d)
                slot["pipe"].run()
                for k, h in slot["host"].items():
                    src = slot["pool"][k]
                    if k in slot["narrow"]:
                        _lib.check(slot["pipe"].lib.mt_cast_f64_f32(D.ptr(src), D.ptr(slot["narrow"][k]), src.numel(),
                                                                    D.stream()), "mt_cast_f64_f32")
                        src = slot["narrow"][k]
                    h.copy_(src, non_blocking=True)
                ev = self.t.cuda.Event()
                ev.record(slot["stream"])
            slot["event"], slot["key"] = ev, key
        # drain in issue order
        n = len(self.slots)
        for j in range(n):
            slot = self.slots[(i + j) % n]
            if slot["event"] is not None:
                yield self._collect(slot)
