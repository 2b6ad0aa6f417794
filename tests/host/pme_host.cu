// Test harness (NOT product code): the __host__ __device__ per-item functions of csrc/pme.cuh compiled for the HOST and
// driven by plain loops, so that tests/test_pme_host.py can check them against oracle/pme.py and the OpenMM goldens
// without a GPU. The FFTs in between are numpy's. T = double.
#include "../../molly.jl_b200/csrc/pme.cuh"

using namespace mb;
using T = double;
using T4 = VT<T>::T4;
using T2 = VT<T>::T2;

extern "C" {
// pos4: n x 4 (x, y, z, q); grid: K0*K1*K2 complex (re, im interleaved), zeroed by the caller
void pmeh_spread(int n, const int* K, const double* L, const double* pos4, double* grid) {
    PmeGeom g{{K[0], K[1], K[2]}, {L[0], L[1], L[2]}};
    for (int s = 0; s < n; s++) pme_spread_atom<T>(s, g, reinterpret_cast<const T4*>(pos4), reinterpret_cast<T2*>(grid));
}
double pmeh_conv(const int* K, const double* L, double f_div_eps, double alpha, const double* bx, const double* by,
                 const double* bz, double* grid) {
    PmeGeom g{{K[0], K[1], K[2]}, {L[0], L[1], L[2]}};
    const double pi = 3.14159265358979323846;
    const double factor = pi * pi / (alpha * alpha), boxfactor = pi * L[0] * L[1] * L[2];
    const size_t total = (size_t)K[0] * K[1] * K[2];
    double e = 0;
    for (size_t idx = 0; idx < total; idx++) e += pme_conv_point<T>(idx, g, f_div_eps, factor, boxfactor, bx, by, bz, reinterpret_cast<T2*>(grid));
    return 0.5 * e;
}
void pmeh_interp(int n, const int* K, const double* L, const double* pos4, const double* grid, double* f4) {
    PmeGeom g{{K[0], K[1], K[2]}, {L[0], L[1], L[2]}};
    for (int s = 0; s < n; s++) pme_interp_atom<T>(s, g, reinterpret_cast<const T4*>(pos4), reinterpret_cast<const T2*>(grid), reinterpret_cast<T4*>(f4));
}
double pmeh_exclusion(int n_pairs, const int* pairs, const double* L, const double* pos4, double* f4, double alpha, double f_div_eps) {
    double e = 0;
    for (int t = 0; t < n_pairs; t++) e += ewald_exclusion_pair<T>(t, pairs, nullptr, reinterpret_cast<const T4*>(pos4), reinterpret_cast<T4*>(f4), L, alpha, f_div_eps);
    return e;
}
}
