"""Export list of meteotools/calc/__init__.py:1-12 for the rows on the hot path:
coordinates (a6), finite differences (a8, a9), thermodynamics (a12) and the range averages
(SURVEY 8f row 1).  FFT differences, quadratic roots and wind-direction conversion are out of
scope."""
from .coordinates import Make_cyclinder_coord, cartesian2cylindrical, \
    Make_radial_axis, Make_tangential_axis, Make_vertical_axis  # noqa: F401
from .differential import FD2, FD2_back, FD2_front, FD2_2, FD2_2_back, \
    FD2_2_front, FD4  # noqa: F401
from .thermaldynamic import calc_H, P_to_Z, Z_to_P, calc_theta, calc_T, \
    calc_saturated_vapor, calc_vapor, \
    calc_saturated_qv, qv_to_vapor, calc_theta_es, \
    calc_Td, calc_qv, calc_theta_v, calc_Tv, \
    calc_rho, calc_Tc, calc_theta_e, calc_dTdz, calc_RH  # noqa: F401
from .averages import calc_Raverage, calc_Zaverage, calc_RZaverage  # noqa: F401
