// cutoffs2.cuh — the two-point cutoffs CubicSplineCutoff and PolynomialCutoff.
// Reference: src/cutoffs.jl:23-29, :39-45 (unchanged up to dist_activation, switched on (r_a, r_c], zero beyond),
// :192-215 (cubic Hermite spline), :217-253 (5th-degree switching polynomial, OpenMM's switching function).
// __host__ __device__ so that a host harness can check the arithmetic without a GPU (tests/host/cutoffs_host.cu,
// tests/test_cutoffs_host.py). Used by pair.cuh's CUTM_TWO_POINT instantiations (lj_term / coul_term); the plain and
// shifted kernel variants never see this code.
#pragma once
#include "common.cuh"

namespace mb {

#ifndef MB_HD
#define MB_HD __host__ __device__ __forceinline__
#endif

enum { CUT_CUBIC_SPLINE = 4, CUT_POLYNOMIAL = 5 };

// In: fr = F(r)/r and e = V(r) of the unmodified interaction at distance r; v_act = V(r_a), f_act = F(r_a).
// Out (only for r > r_a): the switched F/r and V. The caller applies the r <= r_c test.
template <typename T>
MB_HD void cut2_apply(int kind, T ra, T rc, T r, T v_act, T f_act, T& fr, T& e) {
    if (r <= ra) return;
    const T span = rc - ra;
    const T t = (r - ra) / span;
    if (kind == CUT_CUBIC_SPLINE) {
        const T dpe = -f_act;  // dV/dr at r_a
        const T t2 = t * t, t3 = t2 * t;
        e = ((T)2 * t3 - (T)3 * t2 + (T)1) * v_act + (t3 - (T)2 * t2 + t) * span * dpe;
        const T F = -((T)6 * t2 - (T)6 * t) * v_act / span - ((T)3 * t2 - (T)4 * t + (T)1) * dpe;
        fr = F / r;
    } else {
        const T t2 = t * t, t3 = t2 * t;
        const T S = (T)1 - (T)6 * t3 * t2 + (T)15 * t2 * t2 - (T)10 * t3;
        const T dS = ((T)-30 * t2 * t2 + (T)60 * t3 - (T)30 * t2) / span;
        const T e0 = e;
        e = S * e0;
        fr = S * fr - dS * e0 / r;  // F_c = S F - S' V, divided by r
    }
}

// Lennard-Jones under a two-point cutoff: F/r and V at squared distance r2 for (sigma^2, eps)
template <typename T>
MB_HD void lj_cut2(int kind, T ra, T rc, T sig2, T eps, T r2, T& fr, T& e) {
    const T inv_r2 = (T)1 / r2;
    const T s2 = sig2 * inv_r2, s6 = s2 * s2 * s2;
    fr = (T)24 * eps * ((T)2 * s6 * s6 - s6) * inv_r2;
    e = (T)4 * eps * (s6 * s6 - s6);
    const T a2 = sig2 / (ra * ra), a6 = a2 * a2 * a2;
    const T v_act = (T)4 * eps * (a6 * a6 - a6);
    const T f_act = (T)24 * eps * ((T)2 * a6 * a6 - a6) / ra;
    cut2_apply<T>(kind, ra, rc, (T)sqrt((double)r2), v_act, f_act, fr, e);
}

// plain Coulomb (kqq = k_e q_i q_j) under a two-point cutoff
template <typename T>
MB_HD void coul_cut2(int kind, T ra, T rc, T kqq, T r2, T& fr, T& e) {
    const T r = (T)sqrt((double)r2);
    const T inv_r = (T)1 / r;
    fr = kqq * inv_r * inv_r * inv_r;
    e = kqq * inv_r;
    cut2_apply<T>(kind, ra, rc, r, kqq / ra, kqq / (ra * ra), fr, e);
}

}  // namespace mb
