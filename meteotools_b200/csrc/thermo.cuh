// Point-wise moist thermodynamics of calc/thermaldynamic.py, evaluated in the reference's
// (NumPy left-to-right) order so that every +,-,*,/ rounds identically; exp/log/pow are CUDA's
// double-precision functions (<= 1-2 ulp from NumPy's: the 1e-10 tolerance of the north star).
#pragma once
#include "common.cuh"

namespace mt {
namespace th {

// calc/thermaldynamic.py:4-13
constexpr double Rd = 287.0, Cp = 1004.0, g = 9.8;
constexpr double kappa = Rd / Cp;
constexpr double Lv = 2.5e6, A = 2.53e11, B = 5420.0;
constexpr double CpRd = Cp / Rd;
constexpr double A622 = A * 0.622;
constexpr double Lv2 = Lv * Lv;
constexpr double one_m_eps = 1 - 0.622;

// x**k for the positive bases that occur here, as exp(k*log(x)): same special values as pow for
// x <= 0 / NaN / inf with a non-integer constant k, relative error ~ k*|log x| * 1e-16 (measured
// <= 4e-16 against NumPy's pow on the S3 fields), ~2x cheaper than CUDA's pow on the FP64 pipe.
__device__ __forceinline__ double powk(double x, double k) { return exp(k * log(x)); }

__device__ __forceinline__ double theta(double T, double P) { return T * powk(100000 / P, kappa); }          // :84
__device__ __forceinline__ double T_from_theta(double th, double P) { return th * powk(P / 100000, kappa); }  // :101
__device__ __forceinline__ double sat_vapor(double T)                                                          // :116
{
    return 611.2 * exp(17.67 * (T - 273.15) / (T - 273.15 + 243.5));
}
__device__ __forceinline__ double sat_qv_from_es(double es, double P) { return 0.622 * es / (P - es); }      // :151
__device__ __forceinline__ double qv_to_vapor(double P, double qv) { return P * qv / (0.622 + qv); }          // :168
__device__ __forceinline__ double Td_from_vapor(double es)                                                     // :224-225
{
    const double lnes = log(es / 611.2);
    return 243.5 * lnes / (17.67 - lnes) + 273.15;
}
__device__ __forceinline__ double theta_v(double th, double qv, double ql) { return th * (1 + qv * 0.61 - ql); }  // :309
__device__ __forceinline__ double Tv_from_qv(double T, double qv) { return T * (1 + (qv / 0.622)) / (1 + qv); }  // :335
__device__ __forceinline__ double Tv_from_vapor(double T, double vap, double P)                                // :333
{
    return T / (1 - (vap / P) * one_m_eps);
}
__device__ __forceinline__ double rho(double P, double Tv) { return P / (Rd * Tv); }                           // :364
// Condensation temperature, :396:  Tc = B / log(A*0.622/P/qv * (T/Tc2)**(Cp/Rd)).
// log(c * x**a) is evaluated as  g - a*log(Tc2)  with the sweep-invariant
//     g = log(A*0.622/P/qv) + a*log(T)
// computed once per point: one log + one divide per sweep instead of a pow, a log and four
// divides (the sweeps are FP64-pipe bound).  Differs from the reference's rounding by <= 1e-15
// relative (measured on the S3 fields), far inside the 1e-10 tolerance of exp/log/pow chains.
__device__ __forceinline__ double Tc_g(double T, double P, double qv) { return log(A622 / P / qv) + CpRd * log(T); }
__device__ __forceinline__ double Tc_step_g(double g, double Tc2) { return B / (g - CpRd * log(Tc2)); }
__device__ __forceinline__ double exp_factor(double th, double qv, double T)  // :189 / :456
{
    return th * exp(Lv * qv / Cp / T);
}
// same with one division, (Lv/Cp)*qv/T: <= 2 ulp of the exponent (tolerance class), used where the
// kernel is FP64-issue bound
__device__ __forceinline__ double exp_factor1(double th, double qv, double T) { return th * exp((Lv / Cp) * qv / T); }
__device__ __forceinline__ double moist_dTdz(double T, double qvs)                                              // :486
{
    return -g * ((1 + Lv * qvs / Rd / T) / (Cp + Lv2 * qvs * 0.622 / Rd / (T * T)));
}

// density factor of cylindrical_grid.py:423-447 / cartesian_grid.py:279-303:
// a = 1 + qv/0.622 ; b = 1 + qv + qc + qr + ... (left to right) ; a/b
struct Hydro {
    const double *q[6];
    int n;
};
__device__ __forceinline__ double density_factor(double qv, const Hydro &h, int64_t i)
{
    const double a = 1 + qv / 0.622;
    double b = 1 + qv;
    for (int k = 0; k < h.n; ++k) b = b + __ldg(h.q[k] + i);
    return a / b;
}

}  // namespace th
}  // namespace mt
