/*
 * molly_oracle.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 * C restatement of Molly.jl's pairwise non-bonded + VelocityVerlet path.
 * See molly_oracle_impl.h for the reference file:line citations.
 */
#ifndef MOLLY_ORACLE_H
#define MOLLY_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_LJ = 0, ORC_COULOMB = 1, ORC_CRF = 2, ORC_EWALD_REAL = 3 };
enum { ORC_CUT_NONE = 0, ORC_CUT_DISTANCE = 1, ORC_CUT_SHIFTED_POTENTIAL = 2, ORC_CUT_SHIFTED_FORCE = 3,
       ORC_CUT_CUBIC_SPLINE = 4, ORC_CUT_POLYNOMIAL = 5 /* oracle only so far */ };
enum { ORC_MIX_LORENTZ = 0, ORC_MIX_GEOMETRIC = 1 };

/* Same field order as mb_inter_t in include/mollyb200.h so one ctypes struct serves both. */
typedef struct {
    int32_t kind;
    int32_t cutoff_kind;
    double r_cut;
    double r_act;
    double weight_special;
    double coulomb_const;
    double solvent_dielectric;
    double ewald_alpha;
    int32_t sigma_mix;
    int32_t eps_mix;
    int32_t approx_erfc;
    int32_t use_neighbors;
} orc_inter_t;

typedef struct {
    int32_t i, j, special;
} orc_nl_entry_t; /* types.jl:611-654: (i::Int32, j::Int32, special::Bool) padded to 12 B */

typedef struct {
    int64_t n_atoms;
    int32_t dtype; /* 32 or 64: element type of the void* arrays */
    int32_t n_inters;
    double box[3];
    const void *mass, *charge, *sigma, *eps, *lambda; /* lambda may be NULL (all 1) */
    const orc_inter_t *inters;
    /* CSR (both directions, sorted) of excluded / special partners; ptr NULL = none */
    const int64_t *excl_ptr;
    const int32_t *excl_idx;
    const int64_t *spec_ptr;
    const int32_t *spec_idx;
} orc_system_t;

int orc_forces_allpairs(const orc_system_t *s, const void *coords, void *fs, double *pe, double *virial,
                        int n_threads);
int64_t orc_neighbor_list(const orc_system_t *s, const void *coords, double r_list, orc_nl_entry_t **out);
int orc_forces_nl(const orc_system_t *s, const void *coords, const orc_nl_entry_t *list, int64_t n_list,
                  void *fs, double *pe, double *virial, int n_threads);
int orc_simulate_vv(const orc_system_t *s, void *coords, void *vel, double dt, int64_t n_steps,
                    int remove_cm_every, double r_list, int nl_every, int n_threads, double *pe_final);
void orc_remove_cm(const orc_system_t *s, void *vel);
double orc_vector_1D(double c1, double c2, double side);
double orc_wrap_coord_1D(double c, double side);
void orc_free(void *p);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
