/*
 * molly_oracle.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 * Instantiates molly_oracle_impl.h for Float64 and Float32 and dispatches on
 * orc_system_t.dtype.
 */
#include "molly_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL double
#define SUF _f64
#define REAL_IS_FLOAT 0
#include "molly_oracle_impl.h"
#undef REAL
#undef SUF
#undef REAL_IS_FLOAT

#define REAL float
#define SUF _f32
#define REAL_IS_FLOAT 1
#include "molly_oracle_impl.h"
#undef REAL
#undef SUF
#undef REAL_IS_FLOAT

int orc_forces_allpairs(const orc_system_t *s, const void *coords, void *fs, double *pe, double *virial,
                        int n_threads) {
    return s->dtype == 32 ? orc_forces_allpairs_f32(s, coords, fs, pe, virial, n_threads)
                          : orc_forces_allpairs_f64(s, coords, fs, pe, virial, n_threads);
}
int64_t orc_neighbor_list(const orc_system_t *s, const void *coords, double r_list, orc_nl_entry_t **out) {
    return s->dtype == 32 ? orc_neighbor_list_f32(s, coords, r_list, out)
                          : orc_neighbor_list_f64(s, coords, r_list, out);
}
int orc_forces_nl(const orc_system_t *s, const void *coords, const orc_nl_entry_t *list, int64_t n_list,
                  void *fs, double *pe, double *virial, int n_threads) {
    return s->dtype == 32 ? orc_forces_nl_f32(s, coords, list, n_list, fs, pe, virial, n_threads, NULL)
                          : orc_forces_nl_f64(s, coords, list, n_list, fs, pe, virial, n_threads, NULL);
}
int orc_simulate_vv(const orc_system_t *s, void *coords, void *vel, double dt, int64_t n_steps,
                    int remove_cm_every, double r_list, int nl_every, int n_threads, double *pe_final) {
    return s->dtype == 32
               ? orc_simulate_vv_f32(s, coords, vel, dt, n_steps, remove_cm_every, r_list, nl_every, n_threads,
                                     pe_final)
               : orc_simulate_vv_f64(s, coords, vel, dt, n_steps, remove_cm_every, r_list, nl_every, n_threads,
                                     pe_final);
}
void orc_remove_cm(const orc_system_t *s, void *vel) {
    if (s->dtype == 32) orc_remove_cm_f32(s, vel); else orc_remove_cm_f64(s, vel);
}
double orc_vector_1D(double c1, double c2, double side) { return vector_1D_f64(c1, c2, side); }
double orc_wrap_coord_1D(double c, double side) { return wrap_coord_1D_f64(c, side); }
void orc_free(void *p) { free(p); }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
