/*
 * mollyb200.h — C ABI of libmollyb200.so, the B200-native (sm_100a) engine for
 * Molly.jl's pairwise non-bonded + VelocityVerlet hot path.
 *
 * Every entry point is what a Julia `ccall` (or Python ctypes) binds; no C++ or
 * torch types cross the boundary. Each function cites the reference interface it
 * replaces (paths relative to the Molly.jl v0.23.3 tree). See INTEGRATION.md
 * for the Julia-side shim.
 *
 * Conventions
 *  - All functions return an int32 status: 0 = OK, <0 = error; the message is
 *    available through mb_last_error(). Nothing throws across the boundary
 *    (the reference raises Julia `error(...)`, e.g. ext/MollyCUDAExt.jl:733-739;
 *    the shim turns a non-zero status into the same exception).
 *  - Array arguments may be DEVICE pointers (the Julia shim passes CuPtr) or HOST
 *    pointers (the Python harness, the e2e benchmark): the library detects which
 *    with cudaPointerGetAttributes and stages host buffers itself.
 *  - Element type of coordinate / velocity / force / atom arrays is the
 *    context's dtype (32 -> float, 64 -> double), matching System{D,AT,T}.
 *  - Units are Molly's: nm, ps, g/mol, kJ/mol (force kJ mol^-1 nm^-1); the
 *    accel conversion factor is exactly 1.
 *  - There is no CPU fallback: without a CUDA device every call fails loudly.
 */
#ifndef MOLLYB200_H
#define MOLLYB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mb_ctx mb_ctx;

/* interaction kinds: src/interactions/lennard_jones.jl:28-35, coulomb.jl:32-70, :698-747, :1320-1394 */
enum { MB_LJ = 0, MB_COULOMB = 1, MB_CRF = 2, MB_EWALD_REAL = 3 };
/* cutoffs: src/cutoffs.jl:47-253. The two-point cutoffs (CubicSplineCutoff :174-215, PolynomialCutoff :217-253) take
 * dist_activation in r_act and dist_cutoff in r_cut and apply to MB_LJ and MB_COULOMB. */
enum { MB_CUT_NONE = 0, MB_CUT_DISTANCE = 1, MB_CUT_SHIFTED_POTENTIAL = 2, MB_CUT_SHIFTED_FORCE = 3,
       MB_CUT_CUBIC_SPLINE = 4, MB_CUT_POLYNOMIAL = 5 };
/* mixing rules: src/mixing.jl:20-38 */
enum { MB_MIX_LORENTZ = 0, MB_MIX_GEOMETRIC = 1 };

enum {
    MB_OK = 0,
    MB_ERR_INVALID = -1,     /* bad argument / unsupported combination */
    MB_ERR_CUDA = -2,        /* CUDA runtime error */
    MB_ERR_CAPACITY = -3,    /* neighbour/halo capacity overflow (reference: tile overflow error, ext:733-739) */
    MB_ERR_STATE = -4,       /* call order (e.g. forces before atoms were set) */
    MB_ERR_NOGPU = -5        /* no CUDA device: there is no CPU fallback */
};

/* POD descriptor of one PairwiseInteraction (fields of the Julia structs). */
typedef struct {
    int32_t kind;             /* MB_LJ | MB_COULOMB | MB_CRF | MB_EWALD_REAL */
    int32_t cutoff_kind;      /* MB_CUT_* (CRF / Ewald carry their own dist_cutoff in r_cut) */
    double r_cut;             /* inter.cutoff.dist_cutoff or inter.dist_cutoff */
    double r_act;             /* inter.cutoff.dist_activation (MB_CUT_CUBIC_SPLINE / MB_CUT_POLYNOMIAL), else ignored */
    double weight_special;    /* inter.weight_special */
    double coulomb_const;     /* inter.coulomb_const (coulomb.jl:16) */
    double solvent_dielectric;/* CRF (coulomb.jl:676); +inf = conducting */
    double ewald_alpha;       /* CoulombEwald */
    int32_t sigma_mix;        /* MB_MIX_* for sigma (default Lorentz) */
    int32_t eps_mix;          /* MB_MIX_* for epsilon (default geometric) */
    int32_t approx_erfc;      /* CoulombEwald.approximate_erfc (coulomb.jl:1331, default true in the reference): erfc by
                               * calc_erfc's 5-term polynomial (:1384-1393) instead of the exact function */
    int32_t use_neighbors;    /* inter.use_neighbors */
} mb_inter_t;

typedef struct {
    int64_t n_atoms;
    int64_t n_rebuilds;        /* neighbour-structure rebuilds since context creation */
    int64_t n_force_evals;
    int64_t n_steps;           /* VelocityVerlet steps executed */
    int64_t n_list_entries;    /* full-shell list entries (incl. padding) of the last build */
    int64_t n_pairs_in_list;   /* full-shell list entries excluding padding */
    int32_t n_bricks, n_cells[3], brick_dims[3];
    int32_t halo_capacity, list_stride, max_neighbors, max_halo;
    int32_t path;              /* 0 = all-pairs kernel, 1 = cell/brick neighbour-list kernel */
    int32_t violations;        /* fixed-interval policy: steps where an atom moved > skin/2 */
    double r_list;
    int64_t kernel_launches;   /* kernels launched by this context since creation */
    /* device time per category measured with CUDA events on the context's stream while
     * mb_set_profiling(ctx, 1) is active: force kernel, VelocityVerlet kernels, rebuild pipeline */
    double force_ms, vv_ms, rebuild_ms;
    int64_t force_launches, vv_launches, rebuild_launches;
    int32_t graph_mode;        /* last mb_simulate_vv: 1 = CUDA-graph step with conditional rebuild node,
                                * 0 = stream launches, -1 = graph construction failed (stream launches) */
    int32_t n_prunes;          /* always 0 (field kept for ABI stability: the dual-list experiment of round 1 was removed) */
    int32_t peer_transport;    /* decomposed runs: 1 = halo exchange and sum(m v) over NVLink peer memory (IPC-mapped
                                * stores fused into the drift / kick kernels), 0 = NCCL send/recv + all-reduce */
    int32_t reserved_;         /* decomposed runs: the rebuild interval the next call will use (adapted from displacements) */
} mb_stats_t;

const char* mb_last_error(void);
int mb_device_count(void);

/* Context = what BuffersGPU + GPUNeighborFinder hold in the reference
 * (src/force.jl:485-522, src/neighbors.jl:104-115). dtype 32|64. cuda_stream may be NULL. */
int mb_ctx_create(int device, int dtype, void* cuda_stream, mb_ctx** out);
void mb_ctx_destroy(mb_ctx* ctx);

/* Atoms in Molly's bits layout Atom{Int32,T,T,T,T,T} (src/types.jl:466-475):
 * {int32 index; int32 atom_type; T mass; T charge; T sigma; T eps; T lambda; int32 alch_role} =
 * 32 B (f32) / 56 B (f64). Host or device pointer. */
int mb_set_atoms(mb_ctx* ctx, int64_t n, const void* atoms_aos);
/* Same information as plain arrays (what the oracle/tests hold). */
int mb_set_atoms_soa(mb_ctx* ctx, int64_t n, const void* mass, const void* charge, const void* sigma,
                     const void* eps);
/* CubicBoundary side lengths (src/spatial.jl:40). */
int mb_set_box(mb_ctx* ctx, const double side[3]);
/* TriclinicBoundary(bv1, bv2, bv3) (src/spatial.jl:151-215): three basis vectors, row-major (bv1 = basis_vectors[0..2] along
 * x; bv2 in the xy plane; bv3 with a positive z component), approx_images = true. Minimum image as vector() :528-534, wrap
 * as wrap_coords :584-600. Served by the no-list kernel (the reference's triclinic tests are small systems:
 * test/gpu_consistency.jl:287-337); specific interaction lists, PME and decomposed runs are refused for such a box. */
int mb_set_box_triclinic(mb_ctx* ctx, const double basis_vectors[9]);
/* sys.pairwise_inters translated to descriptors (dispatch by type in the reference). */
int mb_set_inters(mb_ctx* ctx, int n_inters, const mb_inter_t* inters);
/* GPUNeighborFinder sparse metadata (src/neighbors.jl:104-115, :171-195): 1-based pairs, any order,
 * duplicates allowed; excluded pairs are skipped by every interaction, special pairs are evaluated
 * with special=true; excluded wins over special. Host pointers. */
int mb_set_exceptions(mb_ctx* ctx, int64_t n_excl, const int32_t* excl_i, const int32_t* excl_j,
                      int64_t n_spec, const int32_t* spec_i, const int32_t* spec_j);
/* Neighbour policy: r_list = finder dist_cutoff (+ buffer); rebuild_every = n_steps of the finder
 * (src/neighbors.jl:327, :671); 0 = displacement-triggered (exact: rebuild when an atom moved
 * more than (r_list - max r_cut)/2 since the last build). */
int mb_set_neighbor_policy(mb_ctx* ctx, double r_list, int rebuild_every);

/* pairwise_forces_loop_gpu! (ext/MollyCUDAExt.jl:845; caller src/force.jl:1228): ADD the pairwise
 * forces for coords (n x 3, xyz packed) into fs_mat (3 x n column-major == n x 3 packed, original
 * atom order) and, if non-NULL, dr (x) f into virial (3x3, column-major, type T). */
int mb_forces(mb_ctx* ctx, const void* coords, void* fs_mat, void* virial, int64_t step_n);
/* pairwise_pe_loop_gpu! (ext/MollyCUDAExt.jl:936; caller src/energy.jl:427): ADD sum of pair
 * energies into pe[0] (type T). */
int mb_energy(mb_ctx* ctx, const void* coords, void* pe, int64_t step_n);
/* forces + energy in one traversal (TotalEnergyLogger-style callers). Either output may be NULL. */
int mb_forces_energy(mb_ctx* ctx, const void* coords, void* fs_mat, void* pe, void* virial,
                     int64_t step_n);

/* Specific (bonded) interaction lists, SURVEY.md §8(f)-1: InteractionList{2,3,4}Atoms (src/types.jl:89-157) with
 * HarmonicBond (kind 0; params k, r0), HarmonicAngle (kind 1; k, theta0), PeriodicTorsion (kind 2; one
 * (periodicity, phase, k) term per entry — a torsion with several terms is listed several times; impropers are the
 * same struct, src/interactions/periodic_torsion.jl:17-142). atom_idx: n_terms x (kind + 2), 1-based; params: double,
 * n_terms x 2 or 3. Host pointers. They are evaluated inside mb_simulate_vv (specific_forces_gpu!, src/force.jl:1231)
 * and by mb_forces_energy_all; mb_forces / mb_energy stay pairwise-only (the pairwise_*_loop_gpu! seam). */
int mb_set_specific(mb_ctx* ctx, int kind, int64_t n_terms, const int32_t* atom_idx, const double* params);
/* forces(sys) / potential_energy(sys) of pairwise + specific + general interactions in one call (ADD semantics). */
int mb_forces_energy_all(mb_ctx* ctx, const void* coords, void* fs_mat, void* pe, int64_t step_n);

/* LJDispersionCorrection general interaction (src/interactions/lennard_jones.jl:163-275; added by setup.jl:2000-2004
 * for cutoff systems): E = (factor_6 + factor_12) / V with the means of eps sigma^6, eps sigma^12 over all i <= j atom
 * pairs (Lorentz sigma, geometric eps), no force, isotropic virial 2 U6 + 4 U12 on the diagonal. The factors are
 * computed once on the host from the atoms (grouped by distinct (sigma, eps), double). dist_cutoff <= 0 switches it
 * off. Added by mb_forces_energy_all (energy) — it is part of OpenMM's lj_only / all_cut energies. */
int mb_set_lj_dispersion_correction(mb_ctx* ctx, double dist_cutoff);

/* Particle-mesh Ewald, SURVEY.md §8(f)-3. GPU parity: tests/test_zz_gpu_pme.py (OpenMM forces_all_pme_exact at the
 * reference's 1e-7 kJ/mol/nm / 1e-5 kJ/mol in f64); the per-item arithmetic is also checked on the host
 * (tests/test_pme_host.py), the plan by mb_pme_plan's test. Replaces the `PME`
 * general interaction (src/interactions/ewald.jl:363-958; constructor PME(dist_cutoff, atoms, boundary; error_tol,
 * order=5, eps_r)) and the `EwaldExclusion` specific interaction list (:979-1055) that src/setup.jl:1903-1912 builds
 * from find_excluded_pairs(eligible, special): pairs = excluded OR special, 1-based. Use together with an
 * MB_EWALD_REAL pairwise interaction of the same r_cut / error_tol (ewald_alpha = sqrt(-ln(2 error_tol)) / r_cut).
 * Once set, mb_forces_energy_all and mb_simulate_vv add the reciprocal-space + exclusion forces (and energies incl.
 * the self and neutralising-background terms) after the pair kernel. order = 0 switches it off; only order 5 exists. */
int mb_set_pme(mb_ctx* ctx, double r_cut, double error_tol, int order, double eps_r, int64_t n_pairs,
               const int32_t* pair_i, const int32_t* pair_j);
/* The host-side PME plan (no GPU needed; what the PME constructor computes, ewald.jl:373, :484-487, :311-361): Ewald
 * alpha, mesh dimensions, and (if moduli_out != NULL, capacity >= K0 + K1 + K2) the B-spline moduli of the three
 * dimensions back to back. */
int mb_pme_plan(const double box[3], double r_cut, double error_tol, int order, double* alpha_out, int32_t mesh_out[3],
                double* moduli_out, int capacity);

/* simulate!(sys, VelocityVerlet(dt, coupling, remove_CM_motion), n_steps) hot loop
 * (src/simulators.jl:547-668): wrap, [CM removal when init_step==0], neighbours, F0, then n_steps of
 * kick / drift / wrap / forces / kick / CM removal (every remove_cm_every steps; 0 = never) /
 * Andersen coupling (prob = dt/tau per atom per step; kT<=0 disables) / neighbour policy.
 * coords, vels: n x 3, updated in place (coords returned wrapped into [0,L)). */
typedef struct {
    double dt;
    int64_t n_steps;
    int64_t init_step;        /* simulate!'s init_step; CM motion is removed up front when 0 */
    int32_t remove_cm_every;  /* VelocityVerlet.remove_CM_motion (default 1) */
    double andersen_kT;       /* k*T in kJ/mol, <= 0: no thermostat */
    double andersen_prob;     /* dt / coupling_const */
    uint64_t rng_ctr1, rng_key; /* the two rand(rng, UInt64) of src/coupling.jl:197-212 */
} mb_vv_params_t;
int mb_simulate_vv(mb_ctx* ctx, void* coords, void* vels, const mb_vv_params_t* p);

/* remove_CM_motion! (ext/MollyCUDAExt.jl:2373; src/spatial.jl:901-929) on an n x 3 velocity array. */
int mb_remove_cm_motion(mb_ctx* ctx, void* vels);
/* kinetic_energy (src/energy.jl:56-70): writes 1/2 sum m v.v to *ke_host (double, host). */
int mb_kinetic_energy(mb_ctx* ctx, const void* vels, double* ke_host);

/* Kinetic energy tensor K = 1/2 sum m v (x) v (src/energy.jl:56-70; the reference copies masses and velocities to the
 * host for it, :58-59): 3x3 symmetric, row-major == column-major, host doubles. */
int mb_kinetic_energy_tensor(mb_ctx* ctx, const void* vels, double* ke_tensor9_host);
/* random_velocities!(sys, temp; rng) (src/spatial.jl:819-831, GPU kernel src/kernels.jl:688-703): fills vels (n x 3, host
 * or device) with Maxwell-Boltzmann velocities, sigma = sqrt(kT / m) per component, zero for massless atoms. Philox4x32-10
 * keyed by the caller's two rand(rng, UInt64); statistical parity with the reference (its uniform -> normal transform
 * lives in PhiloxRNG.jl, which is not vendored: SURVEY.md 8c). kT = k * temp in kJ/mol. */
int mb_random_velocities(mb_ctx* ctx, void* vels, double kT, uint64_t rng_ctr1, uint64_t rng_key);

/* find_neighbors(sys, nf, ..., force=true): force a rebuild from coords now (synchronous; also
 * re-derives capacities). */
int mb_rebuild_neighbors(mb_ctx* ctx, const void* coords);
int mb_stats(mb_ctx* ctx, mb_stats_t* host_out);
int mb_synchronize(mb_ctx* ctx);
/* Multiply the auto-derived halo/list capacities (after MB_ERR_CAPACITY). */
int mb_set_capacity_scale(mb_ctx* ctx, double scale);
/* Tuning overrides (CUDALaunchConfig analogue, src/cuda_config.jl:27-41): brick dims in cells
 * (0 = auto), lanes per i-atom (8, or 0 = default; other values are rejected). */
int mb_set_launch_config(mb_ctx* ctx, const int32_t brick_dims[3], int32_t lanes_per_atom);

/* Per-kernel-category CUDA-event timing (benchmark_gpu_tiles.jl-style stage timers,
 * benchmark/gpu_profile_utils.jl:12-18); results through mb_stats. Resets the accumulators. */
int mb_set_profiling(mb_ctx* ctx, int enable);

/* Spatial decomposition over ranks (one process per GPU; the reference has none, docs/src/documentation.md:1826).
 * The box is cut into z-slabs of whole cell layers; a rank integrates the atoms of its slab and evaluates forces for
 * its bricks. Per step: forward halo exchange of positions between neighbouring slabs (grouped ncclSend/ncclRecv of
 * contiguous slot ranges, 2 cell layers each way; full-shell lists need no reverse force exchange) and one 24-byte
 * all-reduce of sum(m v). At every rebuild (fixed interval, default 20 steps) positions and velocities are all-gathered
 * and every rank re-sorts the replicated system identically. mb_simulate_vv takes and returns the whole system on
 * every rank. rank 0 creates the id, the host runtime (torch.distributed / MPI) broadcasts its 128 bytes. */
int mb_comm_unique_id(void* out128);
int mb_comm_init(mb_ctx* ctx, const void* unique_id128, int rank, int nranks);
/* The host-side plan of the halo exchange (no GPU needed): layer_start = ncz + 1 slot offsets of the cell layers;
 * outputs are (peer, first slot, slot count) triples in the order the exchange posts them. */
int mb_decomp_plan(int ncz, int halo_layers, int nranks, int rank, const int32_t* layer_start, int32_t* send_out,
                   int32_t* n_send, int32_t* recv_out, int32_t* n_recv, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* MOLLYB200_H */
