"""meteotools_b200 -- B200 (sm_100a) implementation of the data-parallel hot path of meteotools.

Same Python surface as the reference's ``meteotools.interpolation``, ``meteotools.calc`` and
``meteotools.grids`` (function names, argument meaning, exceptions); the gfortran/OpenMP
``fastcompute`` backend is replaced by a C-ABI CUDA library built by ``build.py`` into
``lib/fastcompute.so`` (see include/meteotools_b200.h and INTEGRATION.md).

There is NO CPU fallback: importing the compute sub-packages without the built library, or
calling them without a CUDA device, raises.
"""
import os
import platform

# meteotools/__init__.py:5-8
path_of_meteotools = os.path.dirname(os.path.abspath(__file__))
platform_name = platform.system()

__version__ = "0.1.0"
