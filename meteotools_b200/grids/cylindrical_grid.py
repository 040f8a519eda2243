"""Cylindrical (z, theta, r) grid -- same methods and attribute names as the reference's
``meteotools/grids/cylindrical_grid.py`` (:14-597), with every array operation executed by the
CUDA library:

* ``set_data``  (:95-137)  -> ``mt_setdata``: vertical interpolation (wrf.interplevel semantics,
  parity unpinned) of the bounding box of the targets, bilinear gather, terrain mask -- fused,
  batched over variables, only the touched source box is uploaded;
* ``calc_*``    (:264-508) -> ``mt_cyl_d`` with the output set the method needs (the reference's
  attribute-presence dispatch and AttributeError contract are kept);
* ``calc_diagnostics()``   -> the whole cyl_d set in two HBM passes (stage A + stage B).

Fields live on the device in (theta, r, z) memory order ("TRZ": z contiguous), which is what the
gather produces and what the reference's own interpolation returns (a layer-fastest strided
array); through the attribute protocol they appear as (Nz, Ntheta, Nr) arrays.
Spline vertical interpolation (``ver_interp_order=2``, :83-93, :123-127) runs FITPACK's
interpolating spline per column on the device (``mt_spline_levels``, csrc/spline.cu).
"""
import os

import numpy as np

from .. import _device as D
from .. import _lib
from ..calc import Make_cyclinder_coord
from ..exceptions import DimensionError
from ..tools import timer
from .grid_general import gridsystem
from .thermaldynamic_calculation import calc_thermaldynamic

_HYDRO_SETS = (("qc", "qr", "qi", "qs", "qg", "qh"), ("qc", "qr", "qi", "qs", "qg"),
               ("qc", "qr", "qi", "qs"), ("qc", "qr", "qi"), ("qc", "qr"), ("qc",), ("qr",), ())


class cylindrical_grid(gridsystem, calc_thermaldynamic):
    _layout = _lib.MT_LAYOUT_TRZ
    # the stencil plan / work items follow the targets, the sin / cos / r tables follow the axes
    _derived_caches = {'xTR': ('_plan', '_fused_items', '_items_checked'), 'yTR': ('_plan', '_fused_items', '_items_checked'),
                       'theta': ('_grid_cache',), 'r': ('_grid_cache',)}

    def init_grid(self):  # :19-25
        self.meta_convert_to_float('dr', 'dtheta', 'dz', 'dt', 'z_start', 'r_start', 'theta_start')
        self.meta_convert_to_int('Nr', 'Ntheta', 'Nz')
        self.create_coord()

    def create_coord(self):  # :27-40
        self.r = np.arange(self.r_start, self.Nr * self.dr, self.dr)
        self.theta = np.arange(self.theta_start, self.Ntheta * self.dtheta, self.dtheta)
        self.z = np.arange(0, self.Nz * self.dz, self.dz) + self.z_start
        self.set_coord_meta('z', 'vertical', self.z, unit='m', description='Vertical coordinate')
        self.set_coord_meta('theta', 'tangential', self.theta, unit='rad',
                            description='Tangential coordinate')
        self.set_coord_meta('r', 'radial', self.r, unit='m', description='Radial coordinate')
        self.__dict__.pop('_grid_cache', None)

    def _grid_shape3(self):
        return (self.Nz, self.Ntheta, self.Nr)

    def set_horizontal_location(self, offset_x, offset_y):  # :42-49
        xTR, yTR = Make_cyclinder_coord([offset_y, offset_x], self.r, self.theta)
        self.__dict__.pop('_plan', None)
        self.xTR, self.yTR = xTR, yTR

    # ------------------------------------------------------------------ grid descriptor for C
    def _grid_struct(self, nz=None, halo=(0, 0), ends=(True, True)):
        cache = self.__dict__.get('_grid_cache')
        if cache is None:
            # sin/cos tables from NumPy on the host (bit-identical to the reference's
            # np.sin(theta) / np.cos(theta), :266-268)
            cache = dict(r=D.to_device(self.r), sin=D.to_device(np.sin(self.theta)),
                         cos=D.to_device(np.cos(self.theta)))
            self.__dict__['_grid_cache'] = cache
        g = _lib.mt_grid_t()
        g.nz, g.nt, g.nr = (self.Nz if nz is None else nz), self.Ntheta, self.Nr
        g.layout = self._layout
        g.dz, g.dt, g.dr = self.dz, self.dtheta, self.dr
        g.r, g.sin_t, g.cos_t = cache['r'].data_ptr(), cache['sin'].data_ptr(), cache['cos'].data_ptr()
        g.halo_lo, g.halo_hi = halo
        g.is_bottom, g.is_top = int(ends[0]), int(ends[1])
        return g

    # ------------------------------------------------------------------------- set_data (F1)
    def _target_box(self, nx, ny, dx, dy):
        """Bounding box [y0,y1) x [x0,x1) (source indices) of all four-corner stencils."""
        with np.errstate(all='ignore'):
            xi, yi = self.xTR / dx, self.yTR / dy
        if np.nanmax(self.xTR) > dx * (nx - 1) or np.nanmin(self.xTR) < 0:
            raise ValueError(f"X target outside the grid range 0 ~ {dx * (nx - 1)}")
        if np.nanmax(self.yTR) > dy * (ny - 1) or np.nanmin(self.yTR) < 0:
            raise ValueError(f"Y target outside the grid range 0 ~ {dy * (ny - 1)}")
        x0, x1 = int(np.nanmin(xi)), min(int(np.nanmax(xi)) + 2, nx)
        y0, y1 = int(np.nanmin(yi)), min(int(np.nanmax(yi)) + 2, ny)
        return y0, y1, x0, x1

    def _targets_dev(self):
        return self.device('xTR'), self.device('yTR')

    def _interp2d_var(self, dx, dy, var, map_factor=None):
        """2-D variable -> (theta, r) device tensor (interp2D_fast, :104-105)."""
        var = np.asarray(var) if not D.is_tensor(var) else var
        ny, nx = var.shape
        self._target_box(nx, ny, dx, dy)
        if map_factor is not None:   # u10 *= MAPFAC_MX in the field's own dtype, on a device copy
            raw = D.torch().from_numpy(np.ascontiguousarray(var)).cuda() if not D.is_tensor(var) else var.clone()
            var = D.scale_box(raw.reshape(1, ny, nx), map_factor, 0, 0, nx).reshape(ny, nx)
        src = D.to_device(var)
        xd, yd = self._targets_dev()
        out = self._new(2)
        _lib.check(_lib.load().mt_interp2d_layers(D.ptr(src), 1, ny, nx, 0, nx, 1, float(dx), float(dy),
                                                  D.ptr(xd), D.ptr(yd), xd.numel(), D.ptr(out), 0, 1,
                                                  D.stream()), 'mt_interp2d_layers')
        return out

    @timer
    def set_data(self, wrf_dx, wrf_dy, wrf_vertical, ver_interp_order=1, inclusive=False, map_factors=None,
                 **vardict):
        """Interpolate source fields (z, y, x) / (y, x) onto this grid (:95-137).

        ``inclusive``: bracket test of the vertical interpolation, see include/meteotools_b200.h
        (MT_IL_INCLUSIVE); default = the strict test of wrf-python's DINTERP3DZ.
        ``map_factors``: dict name -> (ny, nx) factor, e.g. ``dict(u=MAPFAC_MX, v=MAPFAC_MY, u10=MAPFAC_MX,
        v10=MAPFAC_MY)``: the source field is multiplied by it on the device while it is staged --
        ``wrfout_grid.correct_map_scale`` (grids/wrfout_grid.py:48-55) without a pass over the host arrays,
        in NumPy's arithmetic for the two dtypes."""
        map_factors = dict(map_factors or {})
        if ver_interp_order not in (1, 2):   # the reference leaves `tmp` undefined -> NameError (:110-127)
            raise NameError(f"ver_interp_order must be 1 (linear) or 2 (quadratic spline), got {ver_interp_order}")
        self.varlist3D, self.varlist2D = [], []
        for varname, var in vardict.items():
            if len(var.shape) == 3:
                self.varlist3D.append(varname)
            elif len(var.shape) == 2:
                self._put(varname, self._interp2d_var(wrf_dx, wrf_dy, var, map_factors.get(varname)), ndim=2)
                self.varlist2D.append(varname)
        hgt_idx = None
        if 'hgt' in self.varlist2D:
            self.set_terrain()
            hgt_idx = self._hgt_idx_dev
        if self.varlist3D:
            self._setdata3d(wrf_dx, wrf_dy, wrf_vertical, [vardict[n] for n in self.varlist3D],
                            self.varlist3D, hgt_idx, inclusive, spline_order=ver_interp_order,
                            map_factors=map_factors)

    def _setdata3d(self, dx, dy, vert, sources, names, hgt_idx, inclusive, spline_order=1, map_factors=None):
        """3-D variables: one ``mt_setdata`` per batch of <= 16 variables.  Host arrays are
        uploaded box-only with one strided DMA each (``mt_upload_box``); CUDA tensors are used in
        place (full array, no copy).  ``spline_order`` 2 / 3: the vertical stage is the per-column
        interpolating spline of the reference's ``spline`` method (splrep + splev) instead of
        wrf.interplevel; the result is float64 whatever the input dtype, as in the reference."""
        lib = _lib.load()
        t = D.torch()
        nz, ny, nx = vert.shape
        y0, y1, x0, x1 = self._target_box(nx, ny, dx, dy)
        on_dev = D.is_tensor(vert)
        if any(D.is_tensor(s) != on_dev for s in sources):
            raise TypeError("sources and the vertical coordinate must all be host arrays or all CUDA tensors")

        def is_f32(a):
            return a.dtype == (t.float32 if D.is_tensor(a) else np.float32)
        # direction of DINTERP3DZ from column (0,0) of the FULL coordinate array
        direction = 1 if float(vert[0, 0, 0]) > float(vert[nz - 1, 0, 0]) else 0
        levels = D.to_device(self.z)
        order = 1 if np.all(np.diff(self.z) > 0) else (-1 if np.all(np.diff(self.z) < 0) else 0)
        xd, yd = self._targets_dev()
        n = xd.numel()
        plan = self._plan_for(dx, dy, nx, ny)
        trz = self._layout == _lib.MT_LAYOUT_TRZ
        # resident sources are used in place (region = the full array) unless their rows are not
        # 16-byte aligned for some field (odd nx): then the box is staged device-to-device like a
        # host source, so that the TMA kernel can still read it
        in_place = on_dev and all((nx * (4 if is_f32(a) else 8)) % 16 == 0 for a in [vert] + list(sources))
        if in_place:
            sny, snx, oy, ox = ny, nx, 0, 0
        else:        # region = the box, rows padded to D.box_pitch (never read)
            sny, snx, oy, ox = y1 - y0, D.box_pitch(x1 - x0), y0, x0

        def stage(a, want32, fac=None):
            """Source -> device array the kernels read; `fac`: (ny, nx) map factor applied to the raw
            staged data in ITS dtype, before any widening (wrfout_grid.correct_map_scale)."""
            if on_dev and in_place:
                raw, org, width = a.contiguous(), (0, 0), nx
                if fac is not None:
                    raw = raw.clone()          # never modify the caller's array
            elif on_dev:
                raw, org, width = D.upload_box(a, y0, y1, x0, x1), (y0, x0), x1 - x0
            else:
                a = np.asarray(a)
                if a.dtype not in (np.float32, np.float64):
                    a = a.astype(np.float64)
                raw, org, width = D.upload_box(a, y0, y1, x0, x1), (y0, x0), x1 - x0
            if fac is not None:
                D.scale_box(raw, fac, org[0], org[1], width)
            return raw if (raw.dtype == t.float32 and want32) else D.cast_f64(raw)
        vert_cache = {}
        # float32 sources: wrf-python returns float32 (result rounded), reproduced by ROUND_F32;
        # they are uploaded as float32 (half the PCIe bytes) when the coordinate is float32 too.
        for want32 in (False, True):
            group = [(nm, s) for nm, s in zip(names, sources) if is_f32(s) == want32]
            if not group:
                continue
            use32 = want32 and is_f32(vert)
            dtype = _lib.MT_F32 if use32 else _lib.MT_F64
            if dtype not in vert_cache:
                vert_cache[dtype] = stage(vert, use32)
            vert_d = vert_cache[dtype]
            flags = (_lib.MT_IL_INCLUSIVE if inclusive else 0) | (_lib.MT_IL_ROUND_F32 if want32 else 0)
            if spline_order != 1:
                flags = _lib.MT_IL_SPLINE2 if spline_order == 2 else _lib.MT_IL_SPLINE3
            for b0 in range(0, len(group), _lib.MT_MAX_VARS):
                batch = group[b0:b0 + _lib.MT_MAX_VARS]
                srcs = [stage(s, use32, (map_factors or {}).get(nm)) for nm, s in batch]
                outs = [self._new() for _ in batch]
                mode = os.environ.get("MT_SETDATA", "tma")   # tma (default) | split | fused
                fits = nz <= 254 and spline_order == 1
                if fits and mode == "tma" and lib.mt_setdata_tma_supported(dtype, nz, self.Nz, snx):
                    # one warp-specialised TMA kernel, no intermediate in HBM (csrc/setdata_tma.cu).
                    # It writes target-major, level-fastest rows; a (z, y, x) grid takes them through
                    # the tiled transpose of mt_permute3 (two near-copy passes instead of one
                    # bilinear launch per variable on the level-major path).
                    tw_, th_, al_, mt_ = _lib.tma_tile(dtype)
                    order_t, items_t, nitems, nflag, counter = self._fused_items_for(
                        plan, tw_, th_, max_targets=mt_, spatial=True, align=al_, org_x=ox)
                    rows = outs if trz else [D.empty((n, self.Nz)) for _ in batch]
                    _lib.check(lib.mt_setdata_tma(D.ptr_array(srcs), len(batch), D.ptr(vert_d), dtype, nz, sny,
                                                  snx, snx, oy, ox, D.ptr(levels), self.Nz, direction,
                                                  D.ptr(plan[0]), D.ptr(plan[1]), n, D.ptr(order_t),
                                                  D.ptr(items_t), nitems, nflag, D.ptr(hgt_idx),
                                                  D.ptr_array(rows), self.Nz, flags, D.ptr(counter),
                                                  D.stream()), 'mt_setdata_tma')
                    checked = self.__dict__.setdefault('_items_checked', set())
                    if counter.data_ptr() not in checked:   # once per work-item list: the kernel's verdict on it
                        _lib.check(lib.mt_setdata_tma_status(D.ptr(counter), D.stream()), 'mt_setdata_tma_status')
                        checked.add(counter.data_ptr())
                    if not trz and self.Nz <= 96:
                        # (z, y, x) grid: one batched transpose for the whole batch, the Cartesian
                        # terrain mask applied on the way when the grid has a terrain height
                        zh = self._fused_terrain_mask([nm for nm, _ in batch])
                        _lib.check(lib.mt_rows_to_levels(D.ptr_array(rows), D.ptr_array(outs), len(batch), n, self.Nz,
                                                         D.ptr(zh[0]) if zh else None, D.ptr(zh[1]) if zh else None,
                                                         D.stream()), 'mt_rows_to_levels')
                    elif not trz:
                        n1, n2 = self._grid_shape3()[1:]
                        outs = [D.permute3(r_, (self.Nz, n1, n2), (1, n2 * self.Nz, self.Nz)) for r_ in rows]
                    for (name, _), o in zip(batch, outs):
                        self._put(name, o)
                    continue
                fits = fits and trz and self.Nz % 2 == 0
                if fits and mode == "fused":   # first-generation fused kernel (cp.async), kept for A/B runs
                    th_ = 2 if os.environ.get("MT_SETDATA_TILE") == "16x2" else 4
                    order_t, items_t, nitems, nflag, counter = self._fused_items_for(plan, 16, th_)
                    _lib.check(lib.mt_setdata_fused(D.ptr_array(srcs), len(batch), D.ptr(vert_d), dtype, nz, sny,
                                                    snx, oy, ox, D.ptr(levels), self.Nz, direction,
                                                    D.ptr(plan[0]), D.ptr(plan[1]), n, D.ptr(order_t),
                                                    D.ptr(items_t), nitems, 16, th_, nflag, D.ptr(hgt_idx),
                                                    D.ptr_array(outs), self.Nz, flags, D.ptr(counter),
                                                    D.stream()), 'mt_setdata_fused')
                    for (name, _), o in zip(batch, outs):
                        self._put(name, o)
                    continue
                wsb = lib.mt_setdata_workspace_bytes(len(batch), self.Nz, y0, y1, x0, x1)
                ws = D.empty(((wsb + 7) // 8,))
                _lib.check(lib.mt_setdata(D.ptr_array(srcs), len(batch), D.ptr(vert_d), dtype, nz, sny, snx,
                                          oy, ox, ny, nx, float(dx), float(dy), y0, y1, x0, x1,
                                          D.ptr(levels), self.Nz, order, direction, D.ptr(xd), D.ptr(yd),
                                          D.ptr(plan[0]), D.ptr(plan[1]), n, D.ptr(hgt_idx),
                                          D.ptr_array(outs), 1 if trz else n, self.Nz if trz else 1, flags,
                                          D.ptr(ws), wsb, D.stream()), 'mt_setdata')
                if spline_order != 1 and lib.mt_spline_status(D.stream()) != 0:
                    # scipy.interpolate.splrep: ier = 10 -> ValueError (cylindrical_grid.py:90)
                    raise ValueError("Error on input data")
                for (name, _), o in zip(batch, outs):
                    self._put(name, o)

    def _fused_terrain_mask(self, names):
        """(z, hgt) device arrays when the transpose of a (z, y, x) grid should mask terrain on the way
        (Cartesian grids only; this grid masks by level index inside the kernel)."""
        return None

    def _plan_for(self, dx, dy, nx, ny):
        """Per-target stencil (indices + weights), built once per (grid location, source grid)."""
        key = (float(dx), float(dy), int(nx), int(ny))
        plan = self.__dict__.get('_plan')
        if plan is None or plan[2] != key:
            xd, yd = self._targets_dev()
            n = xd.numel()
            ix = D.empty((n, 4), D.torch().int32)
            w = D.empty((n, 2))
            _lib.check(_lib.load().mt_interp2d_plan(D.ptr(xd), D.ptr(yd), n, float(dx), float(dy), nx, ny,
                                                    D.ptr(ix), D.ptr(w), D.stream()), 'mt_interp2d_plan')
            plan = (ix, w, key)
            self.__dict__['_plan'] = plan
            self.__dict__.pop('_fused_items', None)
            self.__dict__.pop('_items_checked', None)
        return plan

    def _fused_items_for(self, plan, tile_w=16, tile_h=4, max_targets=128, spatial=False, align=1, org_x=0):
        """Targets bucketed by source tile for ``mt_setdata_tma`` / ``mt_setdata_fused`` (built once
        per plan on the host: an argsort of n keys): (order, items, nitems, nflagged, counter).
        ``spatial``: items stay in row-major tile order, so that CTAs working at the same time share
        their tile halos in L2 (TMA kernel); otherwise big items first (short tail).
        ``align``/``org_x``: tile origins satisfy (tile_x - org_x) % align == 0 (the TMA unit needs
        16-byte aligned box starts: 2 columns for float64, 4 for float32)."""
        cache = self.__dict__.setdefault('_fused_items', {})
        key_ = (plan[2], tile_w, tile_h, max_targets, spatial, align, org_x)
        if key_ in cache:
            return cache[key_]
        ix = D.to_host(plan[0]).astype(np.int64)
        valid = ix[:, 0] >= 0
        idx = np.nonzero(valid)[0]
        nflagged = int(ix.shape[0] - idx.size)
        if idx.size:
            x0, y0 = ix[idx, 0], ix[idx, 2]
            bx0, by0 = int(x0.min()), int(y0.min())
            bx0 = org_x + (bx0 - org_x) // align * align
            tx, ty = (x0 - bx0) // tile_w, (y0 - by0) // tile_h
            ntx = int(tx.max()) + 1
            key = ty * ntx + tx
            # within a tile, targets are ordered by source cell: consecutive targets then usually
            # share their four corner columns (the TMA kernel loads them once per pair)
            cell = ((y0 - by0) % tile_h) * tile_w + ((x0 - bx0) % tile_w)
            srt = np.argsort(key * (tile_w * tile_h) + cell, kind='stable')
            skey = key[srt]
            starts = np.flatnonzero(np.r_[True, skey[1:] != skey[:-1]])
            ends = np.r_[starts[1:], skey.size]
            rows = []
            scell = cell[srt]
            pair_align = spatial and os.environ.get("MT_TMA_PAIR_ALIGN", "1") != "0"
            for s0, e0 in zip(starts, ends):
                k = int(skey[s0])
                t_x, t_y = bx0 + (k % ntx) * tile_w, by0 + (k // ntx) * tile_h
                nchunk = -(-(e0 - s0) // max_targets)      # equal chunks, not 128 + remainder
                per = -(-(e0 - s0) // nchunk)
                for c0 in range(s0, e0, per):
                    c1 = min(c0 + per, e0)
                    rows.append((t_x, t_y, c0, c1 - c0))
                    if pair_align:
                        # The kernel takes the item's targets two at a time, and a pair in one source
                        # cell shares its corner loads.  In plain cell order one cell with an odd
                        # count shifts every later cell across the pair boundaries: whole pairs of
                        # each cell first, the odd ones out at the end (61 % -> 74 % shared on S2).
                        cc = scell[c0:c1]
                        first = np.r_[True, cc[1:] != cc[:-1]]
                        grp = np.cumsum(first) - 1
                        pos = np.arange(c1 - c0) - np.flatnonzero(first)[grp]     # rank inside the cell
                        cnt = np.bincount(grp)[grp]
                        left = (cnt % 2 == 1) & (pos == cnt - 1)                  # the odd one out
                        perm = np.r_[np.flatnonzero(~left), np.flatnonzero(left)]
                        srt[c0:c1] = srt[c0:c1][perm]
            items = np.array(rows, dtype=np.int32)
            if not spatial:
                items = items[np.argsort(-items[:, 3], kind='stable')]   # big items first: short tail
            else:
                # spatial order, except that the smallest items go last (largest of them first):
                # the persistent CTAs draw items from a counter, and a tail of small items ends
                # the kernel evenly on all SMs
                ntail = min(len(items) // 4, int(os.environ.get("MT_TMA_TAIL", 148)))
                if ntail > 0:
                    by_size = np.argsort(items[:, 3], kind='stable')
                    tail = by_size[:ntail][::-1]
                    keep = np.ones(len(items), bool)
                    keep[tail] = False
                    items = np.concatenate([items[keep], items[tail]])
            order = idx[srt].astype(np.int32)
        else:
            items, order = np.zeros((0, 4), np.int32), np.zeros((1,), np.int32)
        t = D.torch()
        res = (t.from_numpy(order).cuda(), t.from_numpy(np.ascontiguousarray(items)).cuda() if len(items) else
               t.zeros((1, 4), dtype=t.int32, device='cuda'), int(len(items)), nflagged,
               t.zeros((2,), dtype=t.int32, device='cuda'))   # [work-item counter, status word]
        cache[key_] = res
        return res

    def set_terrain(self):  # :139-176
        hgt = self.device('hgt')
        idx = D.empty((self.Ntheta, self.Nr), D.torch().int32)
        _lib.check(_lib.load().mt_terrain_index_cyl(D.ptr(hgt), self.Ntheta, self.Nr, self.z_start, self.dz,
                                                    D.ptr(idx), D.stream()), 'mt_terrain_index_cyl')
        self.__dict__['_hgt_idx_dev'] = idx

    @property
    def hgt_idx(self):
        return D.to_host(self._hgt_idx_dev)

    @property
    def hgt_mask(self):  # :171-176
        return np.where(self.hgt_idx.reshape(1, self.Ntheta, self.Nr)
                        > np.arange(0, self.Nz, 1).reshape(-1, 1, 1), 1, 0)

    def terrain_mask_vars(self):  # :178-181
        lib = _lib.load()
        n = self.Ntheta * self.Nr
        for varname in self.varlist3D:
            t = self.device(varname)
            self._fields[varname].host = None
            sz, sn = (1, self.Nz) if self._layout == _lib.MT_LAYOUT_TRZ else (n, 1)
            _lib.check(lib.mt_terrain_mask_cyl(D.ptr(t), self.Nz, n, sz, sn, D.ptr(self._hgt_idx_dev),
                                               D.stream()), 'mt_terrain_mask_cyl')

    # ------------------------------------------------------------------ fused dynamics (F2)
    def _hydro_names(self):  # :429-447: which hydrometeors enter the density factor
        for names in _HYDRO_SETS:
            if self._has(*names):
                return names
        return ()

    def _run_cyl_d(self, outputs, stages=_lib.MT_STAGE_ALL, existing=()):
        """One ``mt_cyl_d`` call producing `outputs`.  `existing`: point-wise fields of earlier calls
        (Th_rho, rho) that a stage-B-only call reads from the grid.

        When the grid already holds ``vr`` / ``vt`` and the call does not (re)compute them, they are
        handed in as INPUTS (``mt_dyn_in_t.vr / .vt``) and used as they are -- exactly what the
        reference's methods do with ``self.vr`` / ``self.vt`` (user-assigned or storm-relative winds
        included); u and v are then not needed at all."""
        lib = _lib.load()
        g = self._grid_struct()
        din = _lib.mt_dyn_in_t()
        keep = []   # broadcast temporaries of non-field attributes (a scalar f, ...) stay alive until the launch
        for k in ("u", "v", "w", "p", "T", "Th", "qv", "f"):
            if self._has(k):
                keep.append(self.device(k, ndim=2 if k == "f" else 3))
                setattr(din, k, keep[-1].data_ptr())
        if 'vr' not in outputs and 'vt' not in outputs:
            for k in ("vr", "vt"):
                if self._has(k):
                    keep.append(self.device(k))
                    setattr(din, k, keep[-1].data_ptr())
        hyd = self._hydro_names()
        if not hyd and ('density_factor' in outputs or 'Th_rho' in outputs):
            print('Warning: no liquid / solid hydrometeor mixing ratio set; Th_rho will equal Th_v')
        for i in range(6):
            din.q[i] = self.device(hyd[i]).data_ptr() if i < len(hyd) else None
        dout = _lib.mt_cyl_d_out_t()
        bufs = {}
        for k in outputs:
            bufs[k] = self._new(dtype=D.torch().uint8 if k == 'strain_region' else None)
            setattr(dout, k, bufs[k].data_ptr())
        for k in existing:
            keep.append(self.device(k))
            setattr(dout, k, keep[-1].data_ptr())
        _lib.check(lib.mt_cyl_d(g, din, dout, stages, D.stream()), 'mt_cyl_d')
        for k, b in bufs.items():
            self._put(k, b)

    def _stencil(self, outputs, existing=()):
        """Stencil outputs from the grid's vr / vt (stage B alone) or, when the grid does not hold them
        yet, from u, v in ONE fused call that also leaves vr / vt on the grid -- the reference's
        ``if 'vr' not in dir(self): self.calc_vr_vt()`` followed by the method body."""
        if self._has('vr', 'vt'):
            self._run_cyl_d(outputs, _lib.MT_STAGE_B, existing=existing)
        elif existing:   # inputs of stage B that are fields of the grid: two calls
            self.calc_vr_vt()
            self._run_cyl_d(outputs, _lib.MT_STAGE_B, existing=existing)
        elif self._has('u', 'v'):
            self._run_cyl_d(['vr', 'vt'] + list(outputs), _lib.MT_STAGE_ALL)
        else:
            raise AttributeError('u and v must be set first')

    def calc_diagnostics(self, outputs=None):
        """The whole cyl_d diagnostic set (test_cyl_dynamics.py:148-227) in two HBM passes."""
        names = list(outputs) if outputs is not None else \
            [k for k in _lib.CYL_D_OUT if k not in ('Th',) or not self._has('Th')]
        if not self._has('u', 'v'):
            raise AttributeError('u and v must be set first')
        self._run_cyl_d(names)
        return self

    # ---- per-method API of the reference -------------------------------------------------------
    def calc_vr_vt(self):  # :264-270
        if self._has('u', 'v'):
            self._run_cyl_d(['vr', 'vt'], _lib.MT_STAGE_A)
        else:
            raise AttributeError('u and v must be set first')

    def _vr_vt_2d(self, uname, vname, vrname, vtname):
        """2-D (theta, r) rotation through the same kernel with Nz = 1."""
        lib = _lib.load()
        g = self._grid_struct(nz=1)
        g.layout = _lib.MT_LAYOUT_ZTR
        din = _lib.mt_dyn_in_t()
        din.u, din.v = self.device(uname).data_ptr(), self.device(vname).data_ptr()
        dout = _lib.mt_cyl_d_out_t()
        vr, vt = self._new(2), self._new(2)
        dout.vr, dout.vt = vr.data_ptr(), vt.data_ptr()
        _lib.check(lib.mt_cyl_d(g, din, dout, _lib.MT_STAGE_A, D.stream()), 'mt_cyl_d')
        self._put(vrname, vr, 2)
        self._put(vtname, vt, 2)

    def calc_vr_vt10(self):  # :272-278
        if self._has('u10', 'v10'):
            self._vr_vt_2d('u10', 'v10', 'vr10', 'vt10')
        else:
            raise AttributeError('u10 and v10 must be set first')

    def calc_fric_vr_vt(self):  # :280-288
        if self._has('fric_u', 'fric_v'):
            lib = _lib.load()
            din = _lib.mt_dyn_in_t()
            din.u, din.v = self.device('fric_u').data_ptr(), self.device('fric_v').data_ptr()
            dout = _lib.mt_cyl_d_out_t()
            vr, vt = self._new(), self._new()
            dout.vr, dout.vt = vr.data_ptr(), vt.data_ptr()
            _lib.check(lib.mt_cyl_d(self._grid_struct(), din, dout, _lib.MT_STAGE_A, D.stream()), 'mt_cyl_d')
            self._put('fric_vr', vr)
            self._put('fric_vt', vt)
        else:
            raise AttributeError('fric_u and fric_v must be set first')

    def calc_aam(self):  # :290-297
        if not self._has('f'):
            raise AttributeError('f must be set first')
        if not self._has('vt') or not self._has('vr'):
            self.calc_vr_vt()
        self._run_cyl_d(['aam'], _lib.MT_STAGE_A)   # from the grid's vt, as it is

    def calc_agradient_force(self):  # :335-348
        if not self._has('vt'):
            self.calc_vr_vt()
        if not self._has('rho'):
            self.calc_rho()
        if self._has('p', 'f'):
            self._run_cyl_d(['coriolis_force', 'centrifugal_force'], _lib.MT_STAGE_A)
            self._run_cyl_d(['radial_PG_force', 'agradient_force'], _lib.MT_STAGE_B, existing=('rho',))
        else:
            raise AttributeError('p and f must be set first')

    def calc_wind_speed(self):  # :354-358
        if self._has('u', 'v'):
            self._run_cyl_d(['wind_speed'], _lib.MT_STAGE_A)
        else:
            raise AttributeError('u and v must be set first')

    def calc_wind_speed10(self):  # :360-364
        if self._has('u10', 'v10'):
            lib = _lib.load()
            g = self._grid_struct(nz=1)
            g.layout = _lib.MT_LAYOUT_ZTR
            din = _lib.mt_dyn_in_t()
            din.u, din.v = self.device('u10').data_ptr(), self.device('v10').data_ptr()
            dout = _lib.mt_cyl_d_out_t()
            ws = self._new(2)
            dout.wind_speed = ws.data_ptr()
            _lib.check(lib.mt_cyl_d(g, din, dout, _lib.MT_STAGE_A, D.stream()), 'mt_cyl_d')
            self._put('wind_speed10', ws, 2)
        else:
            raise AttributeError('u10 and v10 must be set first')

    def calc_kinetic_energy10(self):  # :375-381 (raises when vr10 AND vt10 already exist, as written there)
        if not self._has('vr10') or not self._has('vt10'):
            self.calc_vr_vt10()
            lib = _lib.load()
            g = self._grid_struct(nz=1)
            g.layout = _lib.MT_LAYOUT_ZTR
            din = _lib.mt_dyn_in_t()
            zero = D.to_device(np.zeros((self.Ntheta, self.Nr)))     # (vr10^2 + vt10^2 + 0^2)/2: same bits
            din.u, din.v, din.w = self.device('u10').data_ptr(), self.device('v10').data_ptr(), zero.data_ptr()
            dout = _lib.mt_cyl_d_out_t()
            ke = self._new(2)
            dout.kinetic_energy = ke.data_ptr()
            _lib.check(lib.mt_cyl_d(g, din, dout, _lib.MT_STAGE_A, D.stream()), 'mt_cyl_d')
            self._put('kinetic_energy10', ke, 2)
        else:
            raise AttributeError('w must be set first')

    def calc_kinetic_energy(self):  # :366-373
        if not self._has('vr') or not self._has('vt'):
            self.calc_vr_vt()
        if self._has('w'):
            self._run_cyl_d(['kinetic_energy'], _lib.MT_STAGE_A)
        else:
            raise AttributeError('w must be set first')

    def _need_vr_vt(self):
        if not self._has('vr') or not self._has('vt'):
            self.calc_vr_vt()

    def calc_vorticity_3D(self):  # :383-394
        if not self._has('w'):
            self._need_vr_vt()
            raise AttributeError('w must be set first')
        self._stencil(['zeta_r', 'zeta_t', 'zeta_z'])

    def calc_abs_vorticity_3D(self):  # :396-402
        if not self._has('zeta_z'):
            self.calc_vorticity_3D()
        if self._has('f'):
            self._stencil(['abs_zeta_z'])
        else:
            raise AttributeError('f must be set first')

    def calc_divergence_hori(self):  # :404-409
        self._stencil(['div_hori'])

    def calc_divergence(self):  # :412-421
        if not self._has('w'):
            self._need_vr_vt()
            raise AttributeError('w must be set first')
        self._stencil(['div'])

    def calc_density_factor(self):  # :423-447
        if not self._has('qv'):
            raise AttributeError('qv must be set first')
        self._run_cyl_d(['density_factor'], _lib.MT_STAGE_A)

    def calc_density_potential_temperature(self):  # :449-453
        self.calc_density_factor()
        if not self._has('Th'):
            self.calc_theta()
        self._run_cyl_d(['Th_rho'], _lib.MT_STAGE_A)

    def calc_PV(self):  # :455-467
        if not self._has('Th_rho'):
            self.calc_density_potential_temperature()
        if not self._has('abs_zeta_z'):
            self.calc_abs_vorticity_3D()
        if not self._has('rho'):
            self.calc_rho()
        self._stencil(['PV'], existing=('Th_rho', 'rho'))

    def calc_vorticity_z(self):  # :473-477
        self._stencil(['zeta_z'])

    def calc_abs_vorticity_z(self):  # :479-485
        if not self._has('zeta_z'):
            self.calc_vorticity_z()
        if self._has('f'):
            self._stencil(['abs_zeta_z'])
        else:
            raise AttributeError('f must be set first')

    def calc_deformation(self):  # :487-493
        self._stencil(['shear_deform', 'stretch_deform'])

    def calc_filamentation(self):  # :495-508
        if not self._has('shear_deform') or not self._has('stretch_deform'):
            self.calc_deformation()
        if not self._has('zeta_z'):
            self.calc_vorticity_z()
        self._stencil(['filamentation_time', 'strain_region'])

    # ------------------------------------------------------- azimuthal statistics (a13, 8f-1)
    def _azimuthal(self, varname, want_axisym):
        """sym_/asy_(/axisymmetry_) of a 3-D (z,theta,r), 2-D (theta,r) or 1-D (theta) variable
        through mt_azimuthal_stats (one pass over theta per (z, r))."""
        if varname not in dir(self):
            raise AttributeError(f'no variable named {varname}')
        f = self._fields.get(varname)
        t = D.torch()
        if f is not None and f.ndim == 3:
            var, g = self.device(varname), self._grid_struct()
            sym, asy = D.empty((self.Nz, self.Nr)), self._new()
        else:   # (theta, r) or (theta,): the same kernel with nz = 1 (and nr = 1)
            host = np.asarray(getattr(self, varname))
            var = self.device(varname) if f is not None else D.to_device(host)
            g = self._grid_struct(nz=1)
            g.layout = _lib.MT_LAYOUT_ZTR
            if host.ndim == 1:
                g.nr = 1
            sym, asy = D.empty((1, int(g.nr))), t.empty_like(var)
        axs = t.empty_like(sym) if want_axisym else None
        _lib.check(_lib.load().mt_azimuthal_stats(D.ptr(var), g, D.ptr(sym), D.ptr(asy), D.ptr(axs), D.stream()),
                   'mt_azimuthal_stats')

        def small(x):   # (Nz, Nr) / (Nr,) / scalar, as the reference returns them
            h = D.to_host(x)
            if f is not None and f.ndim == 3:
                return h
            return h[0] if h.shape[1] > 1 or np.asarray(getattr(self, varname)).ndim == 2 else h[0, 0]
        object.__setattr__(self, f'sym_{varname}', small(sym))
        if f is not None:
            self._put(f'asy_{varname}', asy, f.ndim)
        else:
            object.__setattr__(self, f'asy_{varname}', D.to_host(asy))
        if want_axisym:
            object.__setattr__(self, f'axisymmetry_{varname}', small(axs))

    def calc_axisymmetric_asymmetric(self, varname):  # :214-235
        self._azimuthal(varname, False)

    def calc_axisymmetry(self, varname):  # :187-212
        self._azimuthal(varname, True)

    def find_rmw_and_speed(self):  # :300-314
        if not self._has('sym_vt'):
            if not self._has('vt'):
                self.calc_vr_vt()
            self.calc_axisymmetric_asymmetric('vt')
        if not self._has('sym_wind_speed'):
            if not self._has('wind_speed'):
                self.calc_wind_speed()
            self.calc_axisymmetric_asymmetric('wind_speed')
        max_wind_r_idx = np.nanargmax(self.sym_wind_speed, axis=1)   # (Nz, Nr) host array: tiny
        self.RMW = self.r[max_wind_r_idx]
        self.wspd_max_RMW = self.sym_wind_speed[np.arange(self.Nz), max_wind_r_idx]
        self.vt_max_RMW = self.sym_vt[np.arange(self.Nz), max_wind_r_idx]

    def find_rmw_and_speed10(self):  # :317-331, as written there
        if not self._has('sym_vt10'):
            if not self._has('vt10'):
                self.calc_vr_vt10()
            self.calc_axisymmetric_asymmetric('vt10')
        if not self._has('sym_wind_speed10'):
            if not self._has('wind_speed10'):
                self.calc_wind_speed10()
            self.calc_axisymmetric_asymmetric('wind_speed10')
        # sym_wind_speed10 is 1-D (Nr,): the reference's axis=1 raises numpy's AxisError here too
        max_wind_r_idx = np.nanargmax(self.sym_wind_speed10, axis=1)
        self.RMW10 = self.r[max_wind_r_idx]
        self.wspd_max_RMW10 = self.sym_wind_speed10[max_wind_r_idx]
        self.vt_max_RMW10 = self.sym_vt[max_wind_r_idx]       # (sic) sym_vt, not sym_vt10

    def rotate_to_shear_direction(self, varname, direction, unit='rad'):  # :237-262
        """``shear_relative_<var>``: the variable rolled along theta by int(direction / dtheta) points
        (3-D (z, theta, r): axis 1; 2-D (theta, r) and 1-D (theta): axis 0)."""
        if unit == 'deg':
            direction = np.deg2rad(direction)
        rotate_idx = int(direction / self.dtheta)
        print(rotate_idx)
        var = np.asarray(getattr(self, varname))
        if var.ndim == 3:
            setattr(self, f'shear_relative_{varname}', np.roll(var, rotate_idx, axis=1))
        elif var.ndim in (1, 2):
            setattr(self, f'shear_relative_{varname}', np.roll(var, rotate_idx, axis=0))

    # ------------------------------------------------------------ range averages (8f-1)
    def _wsum3(self, varname, axis, w, mode, denom):
        """mt_axis_wsum over one logical axis (0: z, 1: theta, 2: r) of a 3-D field in the grid's
        layout; returns a dense C-order CUDA tensor over the two kept axes.  Axis 2 is the
        contiguous axis of the reference's (z, theta, r) array: NumPy's pairwise order there."""
        var = self.device(varname)
        n = (self.Nz, self.Ntheta, self.Nr)
        if self._layout == _lib.MT_LAYOUT_TRZ:
            st = (1, self.Nr * self.Nz, self.Nz)
        else:
            st = (self.Ntheta * self.Nr, self.Nr, 1)
        ka, kb = [i for i in range(3) if i != axis]
        out = D.empty((n[ka], n[kb]))
        wd = D.to_device(np.ascontiguousarray(w, dtype=np.float64))
        if axis == 2:
            mode = int(mode) | _lib.MT_WSUM_PAIRWISE
        _lib.check(_lib.load().mt_axis_wsum(D.ptr(var), n[ka], n[kb], n[axis], st[ka], st[kb], st[axis],
                                            D.ptr(wd), int(mode), float(denom), D.ptr(out), n[kb], 1,
                                            D.stream()), 'mt_axis_wsum')
        return out

    @staticmethod
    def _select(axis, vmin, vmax):
        from ..calc.averages import convert_nan_to_default_value
        vmin, vmax = convert_nan_to_default_value(axis, vmin, vmax)
        return np.logical_and(axis >= vmin, axis <= vmax)

    def calc_horizontal_average(self, varname, tmin=np.nan, tmax=np.nan, rmin=np.nan, rmax=np.nan,
                                unit='rad'):  # :513-535
        from ..calc.averages import calc_Raverage, calc_RZaverage
        from ..exceptions import DimensionError, UnitError
        if unit == 'deg':
            tmin, tmax = np.deg2rad(tmin), np.deg2rad(tmax)
        elif unit == 'rad':
            pass
        elif unit == 'index':
            tmin, tmax = self.theta[tmin], self.theta[tmax]
        else:
            raise UnitError('invalid unit: use deg, rad or index')
        ndim = len(getattr(self, varname).shape)
        if ndim == 3:
            t_sel = self._select(self.theta, tmin, tmax)
            with np.errstate(all='ignore'):
                tmp = self._wsum3(varname, 1, t_sel.astype(np.float64), 0, float(np.sum(t_sel)))   # (Nz, Nr)
            setattr(self, f'havg_{varname}', calc_Raverage(tmp, self.r, 1, rmin, rmax))
        elif ndim == 2:   # (theta, r)
            var = self.device(varname) if varname in self._fields else getattr(self, varname)
            setattr(self, f'havg_{varname}', calc_RZaverage(var, self.theta, self.r, rmin, rmax, tmin, tmax))
        else:
            raise DimensionError('varname is neither 3-D (z, theta, r) nor 2-D (theta, r)')

    def calc_vertical_average(self, varname, zmin=np.nan, zmax=np.nan):  # :537-546
        from ..calc.averages import calc_Zaverage
        from ..exceptions import DimensionError
        ndim = len(getattr(self, varname).shape)
        if ndim == 3:
            z_sel = self._select(self.z, zmin, zmax)
            out = self._wsum3(varname, 0, z_sel.astype(np.float64), 0, float(np.sum(z_sel)))      # (Ntheta, Nr)
            setattr(self, f'zavg_{varname}', D.to_host(out))
        elif ndim == 1:
            setattr(self, f'zavg_{varname}', calc_Zaverage(getattr(self, varname), self.z, 0, zmin, zmax))
        else:
            raise DimensionError('varname is neither 3-D (z, theta, r) nor 1-D (z)')

    def create_filter_axis(self, axis, vmin, vmax):  # :548-552
        faxis = np.ones(axis.shape)
        faxis = np.where(axis >= vmin, faxis, 0)
        faxis = np.where(axis <= vmax, faxis, 0)
        return faxis

    def _nansum_weighted(self, a, b, g, wr=None, wz=None, fr=None, w2d=None):
        scratch, out = D.empty((_lib.MT_NANSUM_SCRATCH,)), D.empty((1,))
        dev = [None if x is None else D.to_device(np.ascontiguousarray(x, dtype=np.float64))
               for x in (wr, wz, fr, w2d)]
        _lib.check(_lib.load().mt_nansum_weighted(D.ptr(a), D.ptr(b), g, D.ptr(dev[0]), D.ptr(dev[1]),
                                                  D.ptr(dev[2]), D.ptr(dev[3]), D.ptr(scratch), D.ptr(out),
                                                  D.stream()), 'mt_nansum_weighted')
        return float(D.to_host(out)[0])

    def calc_IKE(self, rmin=np.nan, rmax=np.nan, zmin=np.nan, zmax=np.nan):  # :554-575
        if not self._has('kinetic_energy'):
            self.calc_kinetic_energy()
        if not self._has('rho'):
            self.calc_rho()
        if np.isnan(rmin):
            rmin = self.r[0]
        if np.isnan(rmax):
            rmax = self.r[-1]
        if np.isnan(zmin):
            zmin = self.z[0]
        if np.isnan(zmax):
            zmax = self.z[-1]
        filter_r = self.create_filter_axis(self.r, rmin, rmax)
        filter_z = self.create_filter_axis(self.z, zmin, zmax)
        wr = self.r * self.dr * self.dtheta * self.dz        # r3d*dr*dtheta*dz (:571), per radius
        self.IKE = self._nansum_weighted(self.device('rho'), self.device('kinetic_energy'), self._grid_struct(),
                                         wr=wr, wz=filter_z, fr=filter_r)

    def calc_IKE10(self, rmax=np.nan, min_wspd10=0., rho=np.nan):  # :577-600
        if not self._has('kinetic_energy10'):
            self.calc_kinetic_energy10()
        if not self._has('wind_speed10'):
            self.calc_wind_speed10()
        if not self._has('rho'):
            # as written in the reference: the level nearest 10 m is only looked up when rho had
            # to be computed here; otherwise the `rho` ARGUMENT (default NaN) is used as it is
            self.calc_rho()
            zidx = int(np.nanargmin(np.abs(self.z - 10.)))
            rho3 = self.device('rho')   # (theta, r, z) memory order: level zidx as a dense (theta, r) copy
            rho2 = D.permute3(rho3.view(-1)[zidx:], (self.Ntheta, self.Nr, 1), (self.Nr * self.Nz, self.Nz, 1))
            if not np.isnan(rho):
                rho2 = D.to_device(D.to_host(rho2) * 0 + rho)      # rho[zidx]*0 + rho keeps the NaNs (2-D, small)
        else:
            rho2 = D.to_device(np.full((self.Ntheta, self.Nr), float(rho)))
        r2d = np.stack([self.r] * self.Ntheta)
        if not np.isnan(rmax):
            r2d = np.where(r2d <= rmax, r2d, 0.)
        r2d = np.where(self.wind_speed10 > min_wspd10, r2d, 0.)
        h = 1.
        volume = r2d * self.dr * self.dtheta * h
        g = self._grid_struct(nz=1)
        g.layout = _lib.MT_LAYOUT_ZTR
        self.IKE10 = self._nansum_weighted(rho2, self.device('kinetic_energy10'), g, w2d=volume)
