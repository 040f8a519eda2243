"""Cartesian (z, y, x) grid -- same methods and attribute names as the reference's
``meteotools/grids/cartesian_grid.py`` (:13-391); arrays live on the device in C order
("ZTR" with t = y, r = x), kernels are the Cartesian instantiation of ``mt_car_d``.

The odd vorticity components of the reference (``zeta_t = dv/dz - dw/dx``, :251) are reproduced
as written.  Smoothers (:142-229, mt_smooth_pass) and range averages (:370-391, mt_axis_wsum) are
the SURVEY 8(f) rows 2 and 1.
"""
import numpy as np

from .. import _device as D
from .. import _lib
from ..tools import timer
from .cylindrical_grid import cylindrical_grid, _HYDRO_SETS
from .grid_general import gridsystem
from .thermaldynamic_calculation import calc_thermaldynamic


class cartesian_grid(gridsystem, calc_thermaldynamic):
    _layout = _lib.MT_LAYOUT_ZTR
    _derived_caches = {'hinterp_target_xxloc': ('_plan', '_fused_items'),
                       'hinterp_target_yyloc': ('_plan', '_fused_items')}

    def init_grid(self):  # :17-25
        self.meta_convert_to_float('dx', 'dy', 'dz', 'z_start', 'SOR_parameter', 'iteration_tolerance')
        self.meta_convert_to_int('Nx', 'Ny', 'Nz', 'max_iteration_times')
        self.create_coord()
        self.varlist3D = []
        self.varlist2D = []

    def make_center_loc(self):  # :28-30
        self.center_xloc = (self.x[0] + self.x[-1]) / 2
        self.center_yloc = (self.y[0] + self.y[-1]) / 2

    def create_coord(self):  # :32-46
        self.x = np.arange(0., self.Nx * self.dx, self.dx)
        self.y = np.arange(0., self.Ny * self.dy, self.dy)
        self.z = np.arange(0, self.Nz * self.dz, self.dz) + self.z_start
        self.make_center_loc()
        xx, yy = np.meshgrid(self.x, self.y)
        object.__setattr__(self, 'xx', xx)   # host coordinate tables, not device fields
        object.__setattr__(self, 'yy', yy)
        object.__setattr__(self, 'dis_from_center',
                           (xx - self.center_xloc) ** 2 + (yy - self.center_yloc) ** 2)
        self.set_coord_meta('z', 'z', self.z, unit='m', description='Z coordinate')
        self.set_coord_meta('y', 'y', self.y, unit='m', description='Y coordinate')
        self.set_coord_meta('x', 'x', self.x, unit='m', description='X coordinate')

    def _grid_shape3(self):
        return (self.Nz, self.Ny, self.Nx)

    # the cylindrical grid's set_data machinery works on (Ntheta, Nr) == (Ny, Nx) targets
    Ntheta = property(lambda self: self.Ny)
    Nr = property(lambda self: self.Nx)

    def set_horizontal_location(self, offset_x, offset_y):  # :48-59
        if not self._has('center_xloc') or not self._has('center_yloc'):
            self.make_center_loc()
        self.hinterp_target_xloc = self.x + offset_x - self.center_xloc
        self.hinterp_target_yloc = self.y + offset_y - self.center_yloc
        xx, yy = np.meshgrid(self.hinterp_target_xloc, self.hinterp_target_yloc)
        self.__dict__.pop('_plan', None)
        self.hinterp_target_xxloc, self.hinterp_target_yyloc = xx, yy

    xTR = property(lambda self: self.hinterp_target_xxloc)
    yTR = property(lambda self: self.hinterp_target_yyloc)

    def _targets_dev(self):
        return self.device('hinterp_target_xxloc'), self.device('hinterp_target_yyloc')

    _target_box = cylindrical_grid._target_box
    _interp2d_var = cylindrical_grid._interp2d_var
    _setdata3d = cylindrical_grid._setdata3d
    _plan_for = cylindrical_grid._plan_for
    _fused_items_for = cylindrical_grid._fused_items_for

    @timer
    def set_data(self, src_dx, src_dy, src_vertical, vertical_interp_order=1, inclusive=False, map_factors=None,
                 **vardict):
        """Interpolate source fields onto this grid (:76-125); terrain: z <= hgt -> NaN (:127-136).
        ``map_factors``: see ``cylindrical_grid.set_data`` (wrfout_grid.correct_map_scale on the device)."""
        map_factors = dict(map_factors or {})
        if vertical_interp_order not in (1, 2, 3):   # the reference leaves `tmp` undefined (:100-112)
            raise NameError(f"vertical_interp_order must be 1, 2 or 3, got {vertical_interp_order}")
        cur3d = []
        self.__dict__.pop('_masked_in_setdata', None)   # names masked inside this call's transpose only
        for varname, var in vardict.items():
            if len(var.shape) == 3:
                self.varlist3D.append(varname)
                cur3d.append(varname)
            elif len(var.shape) == 2:
                self._put(varname, self._interp2d_var(src_dx, src_dy, var, map_factors.get(varname)), ndim=2)
                self.varlist2D.append(varname)
        if cur3d:
            self._setdata3d(src_dx, src_dy, src_vertical, [vardict[n] for n in cur3d], cur3d, None, inclusive,
                            spline_order=vertical_interp_order, map_factors=map_factors)
        if 'hgt' in self.varlist2D:
            self.set_terrain()
            self.terrain_mask_vars(skip=self.__dict__.pop('_masked_in_setdata', ()))

    def _fused_terrain_mask(self, names):
        """The variables of this set_data call are masked (z <= hgt -> NaN, :127-136) inside the
        batched transpose of ``_setdata3d``; ``terrain_mask_vars`` then only visits the others."""
        if 'hgt' not in self.varlist2D:
            return None
        self.__dict__.setdefault('_masked_in_setdata', set()).update(names)
        return D.to_device(self.z), self.device('hgt')

    def set_terrain(self):  # :127-131 (the mask itself is applied by the kernel)
        pass

    @property
    def terrain_mask(self):
        return self.z.reshape(-1, 1, 1) <= np.asarray(self.hgt)[None]

    def terrain_mask_vars(self, skip=()):  # :134-136
        lib = _lib.load()
        n = self.Ny * self.Nx
        zd, hd = D.to_device(self.z), self.device('hgt')
        for varname in self.varlist3D:
            if varname in skip:      # masked already by the transpose of this set_data call
                continue
            t = self.device(varname)
            self._fields[varname].host = None
            _lib.check(lib.mt_terrain_mask_car(D.ptr(t), self.Nz, n, n, 1, D.ptr(zd), D.ptr(hd), D.stream()),
                       'mt_terrain_mask_car')

    # ------------------------------------------------------------------ fused dynamics (F3)
    def _grid_struct(self, nz=None, halo=(0, 0), ends=(True, True)):
        g = _lib.mt_grid_t()
        g.nz, g.nt, g.nr = (self.Nz if nz is None else nz), self.Ny, self.Nx
        g.layout = self._layout
        g.dz, g.dt, g.dr = self.dz, self.dy, self.dx
        g.halo_lo, g.halo_hi = halo
        g.is_bottom, g.is_top = int(ends[0]), int(ends[1])
        return g

    def _hydro_names(self):  # :285-303
        for names in _HYDRO_SETS:
            if self._has(*names):
                return names
        return ()

    def _run_car_d(self, outputs, stages=_lib.MT_STAGE_ALL, existing=()):
        lib = _lib.load()
        din = _lib.mt_dyn_in_t()
        keep = []   # broadcast temporaries of non-field attributes stay alive until the launch
        for k in ("u", "v", "w", "p", "T", "Th", "qv", "f"):
            if self._has(k):
                keep.append(self.device(k, ndim=2 if k == "f" else 3))
                setattr(din, k, keep[-1].data_ptr())
        hyd = self._hydro_names()
        if not hyd and ('density_factor' in outputs or 'Th_rho' in outputs):
            print('Warning: no liquid / solid hydrometeor mixing ratio set; Th_rho will equal Th_v')
        for i in range(6):
            din.q[i] = self.device(hyd[i]).data_ptr() if i < len(hyd) else None
        dout = _lib.mt_car_d_out_t()
        bufs = {}
        for k in outputs:
            bufs[k] = self._new()
            setattr(dout, k, bufs[k].data_ptr())
        for k in existing:
            setattr(dout, k, self.device(k).data_ptr())
        _lib.check(lib.mt_car_d(self._grid_struct(), din, dout, stages, D.stream()), 'mt_car_d')
        for k, b in bufs.items():
            self._put(k, b)

    def calc_diagnostics(self, outputs=None):
        """The whole car_d set (test_car_dynamics.py:144-202) in two HBM passes."""
        names = list(outputs) if outputs is not None else \
            [k for k in _lib.CAR_D_OUT if k != 'Th' or not self._has('Th')]
        if not self._has('u', 'v'):
            raise AttributeError('u and v must be set first')
        self._run_car_d(names)
        return self

    def _uvw(self, *need):
        if not self._has(*need):
            raise AttributeError(', '.join(need) + ' must be set first')

    def calc_wind_speed(self):  # :232-236
        self._uvw('u', 'v')
        self._run_car_d(['wind_speed'], _lib.MT_STAGE_A)

    def calc_kinetic_energy(self):  # :238-243
        self._uvw('u', 'v', 'w')
        self._run_car_d(['kinetic_energy'], _lib.MT_STAGE_A)

    def calc_vorticity_3D(self):  # :248-254
        self._uvw('u', 'v', 'w')
        self._run_car_d(['zeta_r', 'zeta_t', 'zeta_z'], _lib.MT_STAGE_B)

    def calc_abs_vorticity_3D(self):  # :256-262
        if not self._has('zeta_z'):
            self.calc_vorticity_3D()
        if self._has('f'):
            self._run_car_d(['abs_zeta_z'], _lib.MT_STAGE_B)
        else:
            raise AttributeError('f must be set first')

    def calc_divergence_hori(self):  # :264-269
        self._uvw('u', 'v')
        self._run_car_d(['div_hori'], _lib.MT_STAGE_B)

    def calc_divergence(self):  # :271-277
        self._uvw('u', 'v', 'w')
        self._run_car_d(['div'], _lib.MT_STAGE_B)

    def calc_density_factor(self):  # :279-303
        if not self._has('qv'):
            raise AttributeError('qv must be set first')
        self._run_car_d(['density_factor'], _lib.MT_STAGE_A)

    def calc_density_potential_temperature(self):  # :305-309
        self.calc_density_factor()
        if not self._has('Th'):
            self.calc_theta()
        self._run_car_d(['Th_rho'], _lib.MT_STAGE_A)

    def calc_PV(self):  # :311-322
        if not self._has('Th_rho'):
            self.calc_density_potential_temperature()
        if not self._has('abs_zeta_z'):
            self.calc_abs_vorticity_3D()
        if not self._has('rho'):
            self.calc_rho()
        self._run_car_d(['PV'], _lib.MT_STAGE_B, existing=('Th_rho', 'rho'))

    def calc_vorticity_z(self):  # :328-332
        self._uvw('u', 'v')
        self._run_car_d(['zeta_z'], _lib.MT_STAGE_B)

    def calc_abs_vorticity_z(self):  # :334-340
        if not self._has('zeta_z'):
            self.calc_vorticity_z()
        if self._has('f'):
            self._run_car_d(['abs_zeta_z'], _lib.MT_STAGE_B)
        else:
            raise AttributeError('f must be set first')

    def calc_deformation(self):  # :342-349
        self._uvw('u', 'v')
        self._run_car_d(['shear_deform', 'stretch_deform'], _lib.MT_STAGE_B)

    def calc_filamentation(self):  # :351-364
        if not self._has('shear_deform') or not self._has('stretch_deform'):
            self.calc_deformation()
        if not self._has('zeta_z'):
            self.calc_vorticity_z()
        self._run_car_d(['filamentation_time'], _lib.MT_STAGE_B)

    # ------------------------------------------------------------------ smoothers (8f-2)
    def _smooth(self, varname, naxes, ax_a, ax_b, center_weight, passes, dims, var):
        """`passes` Jacobi sweeps of mt_smooth_pass over a dense C-order (n0,n1,n2) device array
        (ping-pong buffers; the ghost cells are the ORIGINAL edge values, :159-164)."""
        lib = _lib.load()
        v0 = D.to_device(var).contiguous().view(*dims)
        cur, nxt = v0, None
        denom = float(2 * naxes + center_weight)
        for _ in range(int(passes)):
            nxt = D.empty(dims) if nxt is None or nxt is v0 else nxt
            _lib.check(lib.mt_smooth_pass(D.ptr(v0), D.ptr(cur), D.ptr(nxt), dims[0], dims[1], dims[2], naxes,
                                          ax_a, ax_b, float(center_weight), denom, D.stream()), 'mt_smooth_pass')
            cur, nxt = nxt, (cur if cur is not v0 else None)
        return cur

    def _smooth_source(self, varname):
        if varname in self._fields and self._fields[varname].ndim == 3:
            return self.device(varname)           # (Nz, Ny, Nx) C order on the device already
        return np.asarray(getattr(self, varname), dtype=np.float64)

    def _smooth_store(self, name, result, shape):
        result = result.view(*shape)
        if self._is_field_shape(shape):
            self._put(name, result, len(shape))
        else:
            object.__setattr__(self, name, D.to_host(result))

    def smooth1d(self, varname, axis=0, center_weight=2, passes=1):  # :142-167
        """1-c-1 smoothing along `axis`; result in smooth1d_<var>."""
        var = self._smooth_source(varname)
        shape = tuple(int(x) for x in var.shape)
        axis %= len(shape)
        outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
        inner = int(np.prod(shape[axis + 1:], dtype=np.int64)) if axis < len(shape) - 1 else 1
        out = self._smooth(varname, 1, 1, 0, center_weight, passes, (outer, shape[axis], inner), var)
        self._smooth_store(f'smooth1d_{varname}', out, shape)

    def smooth2d(self, varname, axis=(1, 2), center_weight=2, passes=1):  # :169-199
        """Five-point smoothing over the two axes `axis`; result in smooth2d_<var> (and, as in the
        reference, smooth1d_<var> is reset to zeros of the axis-moved shape, :183)."""
        var = self._smooth_source(varname)
        shape = tuple(int(x) for x in var.shape)
        a0, a1 = int(axis[0]), int(axis[1])
        moved = (shape[a0], shape[a1]) + tuple(n for i, n in enumerate(shape) if i not in (a0, a1))
        object.__setattr__(self, f'smooth1d_{varname}', np.zeros(moved))
        if len(shape) == 3:
            dims, ax = shape, (a0, a1)
        elif len(shape) == 2:
            dims, ax = (1,) + shape, (a0 + 1, a1 + 1)
        else:
            raise ValueError('smooth2d supports 2-D and 3-D variables')
        out = self._smooth(varname, 2, ax[0], ax[1], center_weight, passes, dims, var)
        self._smooth_store(f'smooth2d_{varname}', out, shape)

    def smooth3d(self, varname, center_weight=2, passes=1):  # :201-229
        """Seven-point smoothing; result in smooth3d_<var> (smooth1d_<var> is reset to zeros, :213)."""
        var = self._smooth_source(varname)
        shape = tuple(int(x) for x in var.shape)
        if len(shape) != 3:
            raise ValueError('smooth3d needs a 3-D variable')
        self.__dict__.pop(f'smooth1d_{varname}', None)
        setattr(self, f'smooth1d_{varname}', np.zeros(shape))
        out = self._smooth(varname, 3, 0, 1, center_weight, passes, shape, var)
        self._smooth_store(f'smooth3d_{varname}', out, shape)

    # ------------------------------------------------------------------ range averages (8f-1)
    _wsum3 = cylindrical_grid._wsum3
    _select = staticmethod(cylindrical_grid._select)

    def calc_horizontal_average(self, varname, xmin=np.nan, xmax=np.nan, ymin=np.nan, ymax=np.nan):  # :370-381
        from ..calc.averages import calc_Zaverage
        from ..exceptions import DimensionError
        ndim = len(getattr(self, varname).shape)
        if ndim == 3:
            x_sel = self._select(self.x, xmin, xmax)
            tmp = self._wsum3(varname, 2, x_sel.astype(np.float64), 0, float(np.sum(x_sel)))       # (Nz, Ny)
            setattr(self, f'havg_{varname}', calc_Zaverage(tmp, self.y, 1, ymin, ymax))
        elif ndim == 2:   # (y, x)
            var = self.device(varname) if varname in self._fields else getattr(self, varname)
            tmp = calc_Zaverage(var, self.x, 1, xmin, xmax)
            setattr(self, f'havg_{varname}', calc_Zaverage(tmp, self.y, 0, ymin, ymax))
        else:
            raise DimensionError('varname is neither 3-D (z, y, x) nor 2-D (y, x)')

    def calc_vertical_average(self, varname, zmin=np.nan, zmax=np.nan):  # :383-391
        from ..calc.averages import calc_Zaverage
        from ..exceptions import DimensionError
        ndim = len(getattr(self, varname).shape)
        if ndim == 3:
            z_sel = self._select(self.z, zmin, zmax)
            out = self._wsum3(varname, 0, z_sel.astype(np.float64), 0, float(np.sum(z_sel)))       # (Ny, Nx)
            setattr(self, f'zavg_{varname}', D.to_host(out))
        elif ndim == 1:
            setattr(self, f'zavg_{varname}', calc_Zaverage(getattr(self, varname), self.z, 0, zmin, zmax))
        else:
            raise DimensionError('varname is neither 3-D (z, y, x) nor 1-D (z)')
