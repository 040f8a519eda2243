"""Export list of meteotools/grids/__init__.py:1-3.  ``wrfout_grid`` (netCDF / wrf-python file
reading, wrfout_grid.py) is I/O outside the hot path and is not provided."""
from .cylindrical_grid import cylindrical_grid  # noqa: F401
from .cartesian_grid import cartesian_grid  # noqa: F401
