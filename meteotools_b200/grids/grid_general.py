"""Base class of the grid objects: settings -> attributes (meteotools/grids/grid_general.py:8-75)
plus the DEVICE-RESIDENT FIELD STORE that replaces the reference's "every field is a NumPy
attribute" model.

Field store
-----------
Grid-shaped floating-point arrays assigned to a grid (``cyl.u = array``) or produced by
``set_data`` / ``calc_*`` live in HBM as float64 tensors in the grid's canonical layout and are
handed to the CUDA library by pointer.  The reference's attribute protocol keeps working:

* ``cyl.u``            -> NumPy array of shape (Nz, Ntheta, Nr) (lazy D2H, cached, READ-ONLY:
                          modify by assigning a new array, never in place);
* ``cyl.u = a``        -> replaces the field (uploaded on first use);
* ``del cyl.u``        -> removes it; ``'u' in dir(cyl)`` is what every ``calc_*`` tests, exactly
                          as the reference does;
* ``cyl.device('u')``  -> the CUDA tensor itself (no copy), for chaining device-side work.

netCDF I/O (create_ncfile / ncfile_addvar / ncfile_readvars, grid_general.py:107-156) is out of
scope: those methods raise NotImplementedError.
"""
import numpy as np

from .. import _device as D
from .. import _lib


class _Field:
    __slots__ = ("host", "dev", "ndim")

    def __init__(self, host=None, dev=None, ndim=3):
        self.host, self.dev, self.ndim = host, dev, ndim


class gridsystem:
    #: memory layout of 3-D fields on the device (set by the subclasses)
    _layout = _lib.MT_LAYOUT_ZTR

    def set_wrfpython_cpus(self, cpus):  # grid_general.py:14-15 (wrf-python's OpenMP: not used)
        self._cpus = cpus

    def __init__(self, settings=None, wrf_prefix='wrfout_', interp_cpus=1):
        object.__setattr__(self, "_fields", {})
        self.wrf_prefix = wrf_prefix
        self.interp_cpus = interp_cpus
        self.set_wrfpython_cpus(self.interp_cpus)
        self.meta_list = []
        self.var_meta = dict()
        self.coords = dict()
        self.coord_ncname = dict()
        self.ncname_convert = dict(vertical='z', tangential='theta', radial='r')
        if settings is not None:
            if isinstance(settings, dict):
                self.set_meta(settings)
                self.init_grid()
            elif isinstance(settings, str):
                self.open_ncfile(settings)

    # ---- settings -> attributes (grid_general.py:51-75), without exec ----------------------
    def set_meta(self, settings):
        for name, value in settings.items():
            self.meta_list.append(name)
            setattr(self, name, value)

    def init_grid(self):
        pass

    def meta_convert_to_float(self, *metalist):
        for meta in metalist:
            setattr(self, meta, float(getattr(self, meta)))

    def meta_convert_to_int(self, *metalist):
        for meta in metalist:
            v = getattr(self, meta)
            if v != int(v):
                print(f'Warning: wrfout setting "{meta}" is not a integer and will ignore the decimal.')
            setattr(self, meta, int(v))

    def set_var_meta(self, varname, **metadata):
        self.var_meta.setdefault(varname, dict()).update(metadata)

    def set_coord_meta(self, name, ncname, coord, **metadata):
        self.coords[name] = coord
        self.coord_ncname[name] = ncname
        self.set_var_meta(name, **metadata)

    def open_ncfile(self, filename):
        raise NotImplementedError("netCDF I/O is outside the accelerated hot path (SURVEY 2 #12)")

    create_ncfile = ncfile_addvar = ncfile_readvar = ncfile_readvars = ncfile_addcoord = open_ncfile

    # ---- field store -------------------------------------------------------------------------
    def _grid_shape3(self):
        raise NotImplementedError

    def _is_field_shape(self, shape):
        s3 = self._grid_shape3()
        return tuple(shape) == s3 or tuple(shape) == s3[1:]

    #: attribute name -> cached objects (in ``__dict__``) that are derived from it and must be rebuilt
    #: when the user assigns it directly, as the reference permits (``cyl.xTR = ...``)
    _derived_caches = {}

    def __setattr__(self, name, value):
        fields = self.__dict__.get("_fields")
        for cache in self._derived_caches.get(name, ()):
            self.__dict__.pop(cache, None)
        if fields is not None and not name.startswith("_"):
            is_np = isinstance(value, np.ndarray) and value.dtype.kind == "f"
            is_t = D.is_tensor(value) and value.is_floating_point()
            if (is_np or is_t) and self._has_grid() and self._is_field_shape(value.shape):
                self.__dict__.pop(name, None)
                if is_np:
                    view = value.view()
                    view.flags.writeable = False
                    fields[name] = _Field(host=view, ndim=value.ndim)
                else:
                    fields[name] = _Field(dev=self._canonical(D.to_device(value)), ndim=value.dim())
                return
            fields.pop(name, None)
        object.__setattr__(self, name, value)

    def _has_grid(self):
        try:
            self._grid_shape3()
            return True
        except (AttributeError, NotImplementedError):
            return False

    def __getattr__(self, name):  # only reached when normal lookup fails
        fields = self.__dict__.get("_fields")
        if fields is not None and name in fields:
            return self._host_view(name)
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def __delattr__(self, name):
        fields = self.__dict__.get("_fields")
        if fields is not None and name in fields:
            del fields[name]
            return
        object.__delattr__(self, name)

    def __dir__(self):
        return list(super().__dir__()) + list(self.__dict__.get("_fields", {}))

    def _has(self, *names):
        """``name in dir(self)`` of the reference, without building the dir() list."""
        f, d = self._fields, self.__dict__
        return all((n in f) or (n in d) or hasattr(type(self), n) for n in names)

    # canonical device layout <-> user-visible (z, t, r) C-order -----------------------------
    def _canonical(self, t):
        """(Nz,Nt,Nr) C-order device tensor -> canonical layout (a permuted copy for TRZ)."""
        if t.dim() == 3 and self._layout == _lib.MT_LAYOUT_TRZ:
            nz, nt, nr = t.shape
            # dst (nt, nr, nz): dst[i,j,k] = src[k*nt*nr + i*nr + j]
            return D.permute3(t, (nt, nr, nz), (nr, 1, nt * nr))
        return t

    def device(self, name, ndim=3):
        """CUDA tensor of a field in the canonical layout ((Nt,Nr,Nz) memory order for TRZ grids,
        (Nz,Nt,Nr) for ZTR grids; (Nt,Nr) for 2-D fields), uploading it if it is host-only.

        An attribute that is not a stored field -- a scalar (``cyl.f = 5e-5``), an integer array, a
        1-D profile -- is broadcast to the grid shape (``ndim`` = 3: (Nz,Nt,Nr), 2: (Nt,Nr)) as
        float64, the way the reference's NumPy expressions broadcast it (``self.f*self.vt``)."""
        f = self._fields.get(name)
        if f is None:
            if name not in self.__dict__:
                raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")
            val = np.asarray(self.__dict__[name])
            shape = self._grid_shape3() if ndim == 3 else self._grid_shape3()[1:]
            if val.dtype.kind not in "fiub":
                raise TypeError(f"attribute {name!r} ({val.dtype}) is not numeric: expected an array "
                                f"broadcastable to {shape}")
            try:
                b = np.broadcast_to(val, shape)
            except ValueError:
                raise TypeError(f"attribute {name!r} of shape {val.shape} cannot be broadcast to the grid "
                                f"shape {shape}") from None
            return self._canonical(D.to_device(np.ascontiguousarray(b, dtype=np.float64)))
        if f.ndim == 3 and ndim == 2:
            raise TypeError(f"field {name!r} is 3-D where a 2-D (Ntheta, Nr) / (Ny, Nx) field is expected")
        if f.dev is None:
            f.dev = self._canonical(D.to_device(f.host))
        return f.dev

    def _put(self, name, tensor, ndim=3):
        """Store a device tensor (already in canonical layout) as field `name`."""
        self.__dict__.pop(name, None)
        self._fields[name] = _Field(dev=tensor, ndim=ndim)

    def _new(self, ndim=3, dtype=None):
        nz, nt, nr = self._grid_shape3()
        if ndim == 2:
            return D.empty((nt, nr), dtype)
        return D.empty((nt, nr, nz) if self._layout == _lib.MT_LAYOUT_TRZ else (nz, nt, nr), dtype)

    def _host_view(self, name):
        f = self._fields[name]
        if f.host is None:
            a = D.to_host(f.dev)
            if f.ndim == 3 and self._layout == _lib.MT_LAYOUT_TRZ:
                a = a.transpose(2, 0, 1)  # (Nt,Nr,Nz) memory -> (Nz,Nt,Nr) view, like the reference's
            if a.dtype == np.uint8:
                a = a.astype(bool)
            a.flags.writeable = False
            f.host = a
        return f.host

    def to_host(self, *names):
        """Materialise the given fields (all if none given) as NumPy arrays: {name: array}."""
        return {n: self._host_view(n) for n in (names or list(self._fields))}
