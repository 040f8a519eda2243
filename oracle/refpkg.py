"""The UNMODIFIED reference Python package as a CPU baseline / checker -- TEST INFRASTRUCTURE ONLY.

``install()`` (called by ``__graft_entry__.build()`` in the build container, where
``/root/reference`` exists) copies ``/root/reference/meteotools`` -- byte for byte, minus the
colour tables and ``old_code`` -- into the git-ignored ``baseline/_ref/meteotools`` and puts the
oracle's C restatement of ``fastcompute.f90`` there as ``lib/fastcompute.so`` (gfortran is absent;
``oracle/_ref/fastcompute.so`` is used instead when it exists).  ``baseline/_ref`` is NOT
gpurun-ignored, so the copy travels to the GPU box, where ``/root/reference`` does not exist.

``import_reference()`` imports that copy with stub ``wrf`` / ``netCDF4`` modules
(``wrf.interplevel`` = the oracle's DINTERP3DZ restatement, parity unpinned).  Everything else --
``meteotools.interpolation`` (the ``asfortranarray`` copies, the checks), ``meteotools.calc`` and
``meteotools.grids`` (``cylindrical_grid.set_data``, ``calc_*``) -- is the reference's own code.

Nothing under ``meteotools_b200/`` may import this module.
"""
from __future__ import annotations

import os
import shutil
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
REF_SRC = "/root/reference/meteotools"
REF_DST = os.path.join(ROOT, "baseline", "_ref", "meteotools")


def install(force: bool = False) -> str | None:
    """Copy the reference package to baseline/_ref (build container only).  Returns the
    destination or None when the reference tree is absent."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if available() else None
    if force and os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    if not os.path.exists(os.path.join(REF_DST, "__init__.py")):
        os.makedirs(os.path.dirname(REF_DST), exist_ok=True)
        shutil.copytree(REF_SRC, REF_DST,
                        ignore=shutil.ignore_patterns("colors", "old_code", "*.png", "*.html", "__pycache__"))
    _install_lib()
    return REF_DST


def _install_lib():
    from . import _REF_SO, build
    so = _REF_SO if os.path.exists(_REF_SO) else build()
    lib_dir = os.path.join(REF_DST, "lib")
    os.makedirs(lib_dir, exist_ok=True)
    dst = os.path.join(lib_dir, "fastcompute.so")
    if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(so):
        shutil.copy(so, dst)


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, "__init__.py"))


def import_reference():
    """The reference package (module object) or None when baseline/_ref holds no copy."""
    if not available():
        return None
    if "meteotools" in sys.modules and getattr(sys.modules["meteotools"], "__file__", "").startswith(REF_DST):
        return sys.modules["meteotools"]
    import oracle
    _install_lib()
    wrf = types.ModuleType("wrf")
    wrf.omp_set_num_threads = lambda n: None
    wrf.interplevel = lambda var, vert, lev: oracle.interplevel(var, vert, lev)
    nc = types.ModuleType("netCDF4")
    nc.Dataset = object
    sys.modules.setdefault("wrf", wrf)
    sys.modules.setdefault("netCDF4", nc)
    parent = os.path.dirname(REF_DST)
    if parent not in sys.path:
        sys.path.insert(0, parent)
    import meteotools  # noqa: F401
    import meteotools.calc  # noqa: F401
    import meteotools.grids  # noqa: F401
    import meteotools.interpolation  # noqa: F401
    return sys.modules["meteotools"]
