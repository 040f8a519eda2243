// Finite-difference formulas of calc/differential.py, written once and shared by the generic
// mt_fd kernel and the fused diagnostic kernels.  `v(j)` returns element j along the
// differentiated axis; the evaluation order is NumPy's left-to-right order so that every
// rounding matches (e.g. (-3*v0 + 4*v1 - v2)/2/delta; note "/2/delta", not "/(2*delta)").
#pragma once
#include "common.cuh"

namespace mt {

// FD2, differential.py:30-34.  The axis segment [lo, hi) is what the caller may READ (halo
// levels of a level shard included); `first`/`last` tell whether index 0 / n-1 of the segment
// are the GLOBAL ends (one-sided formulas) or interior points of a shard (centred).
template <class V>
__device__ __forceinline__ double fd2(V v, int i, int n, double delta)
{
    if (i == 0) return (-3 * v(0) + 4 * v(1) - v(2)) / 2 / delta;
    if (i == n - 1) return (3 * v(n - 1) - 4 * v(n - 2) + v(n - 3)) / 2 / delta;
    return (v(i + 1) - v(i - 1)) / 2 / delta;
}

// level-shard aware variant: i is a local index in an array of n levels that holds halo levels;
// bottom/top say whether local 0 / n-1 are the global first / last level.
template <class V>
__device__ __forceinline__ double fd2_shard(V v, int i, int n, double delta, bool bottom, bool top)
{
    if (i == 0 && bottom) return (-3 * v(0) + 4 * v(1) - v(2)) / 2 / delta;
    if (i == n - 1 && top) return (3 * v(n - 1) - 4 * v(n - 2) + v(n - 3)) / 2 / delta;
    return (v(i + 1) - v(i - 1)) / 2 / delta;
}

template <class V>
__device__ __forceinline__ double fd4(V v, int i, int n, double delta)  // differential.py:66-72
{
    if (i == 0) return (-3 * v(0) + 4 * v(1) - v(2)) / 2 / delta;
    if (i == 1) return (v(2) - v(0)) / 2 / delta;
    if (i == n - 1) return (3 * v(n - 1) - 4 * v(n - 2) + v(n - 3)) / 2 / delta;
    if (i == n - 2) return (v(n - 1) - v(n - 3)) / 2 / delta;
    return (-v(i + 2) + 8 * v(i + 1) - 8 * v(i - 1) + v(i - 2)) / 12 / delta;
}

template <class V>
__device__ __forceinline__ double fd2_front(V v, int i, int n, double delta)  // :92-94
{
    if (i >= n - 2) return nan_f64();
    return (-3 * v(i) + 4 * v(i + 1) - v(i + 2)) / 2 / delta;
}

template <class V>
__device__ __forceinline__ double fd2_back(V v, int i, int n, double delta)  // :113-115
{
    if (i < 2) return nan_f64();
    return (3 * v(i) - 4 * v(i - 1) + v(i - 2)) / 2 / delta;
}

template <class V>
__device__ __forceinline__ double fd2_2(V v, int i, int n, double delta)  // :134-138
{
    if (i == 0) return (2 * v(0) - 5 * v(1) + 4 * v(2) - v(3)) / delta / delta;
    if (i == n - 1) return (2 * v(n - 1) - 5 * v(n - 2) + 4 * v(n - 3) - v(n - 4)) / delta / delta;
    return (v(i + 1) - 2 * v(i) + v(i - 1)) / delta / delta;
}

template <class V>
__device__ __forceinline__ double fd2_2_front(V v, int i, int n, double delta)  // :157-159
{
    if (i >= n - 3) return nan_f64();
    return (2 * v(i) - 5 * v(i + 1) + 4 * v(i + 2) - v(i + 3)) / delta / delta;
}

template <class V>
__device__ __forceinline__ double fd2_2_back(V v, int i, int n, double delta)  // :178-180
{
    if (i < 3) return nan_f64();
    return (2 * v(i) - 5 * v(i - 1) + 4 * v(i - 2) - v(i - 3)) / delta / delta;
}

}  // namespace mt
