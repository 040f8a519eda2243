"""Same export list as meteotools/interpolation/__init__.py:1-4.  Importing this package loads
lib/fastcompute.so (as the reference does at import, interp_2d.py:13-16) and raises
``BackendError`` if it has not been built."""
from .. import _lib as _backend

_backend.load()

from .interp_2d import interp2D_fast, interp2D_fast_layers  # noqa: E402,F401
from .interp_3d import interp3D_fast, interp3D_fast_layers  # noqa: E402,F401
from .interp_1d import interp1D_fast, interp1D_nonequal, \
    interp1D_fast_layers, interp1D_nonequal_layers  # noqa: E402,F401
from .interplevel import interplevel  # noqa: E402,F401  (wrf.interplevel semantics)
