// Test harness (NOT product code): csrc/cutoffs2.cuh compiled for the host.
#include "../../molly.jl_b200/csrc/cutoffs2.cuh"
extern "C" {
void cut2h_lj(int kind, double ra, double rc, double sigma, double eps, double r, double* fr, double* e) {
    mb::lj_cut2<double>(kind, ra, rc, sigma * sigma, eps, r * r, *fr, *e);
}
void cut2h_lj_f32(int kind, float ra, float rc, float sigma, float eps, float r, float* fr, float* e) {
    mb::lj_cut2<float>(kind, ra, rc, sigma * sigma, eps, r * r, *fr, *e);
}
void cut2h_coul(int kind, double ra, double rc, double kqq, double r, double* fr, double* e) {
    mb::coul_cut2<double>(kind, ra, rc, kqq, r * r, *fr, *e);
}
}
