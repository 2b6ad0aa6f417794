"""ctypes binding of libmollyb200.so (include/mollyb200.h). No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOLLYB200_LIB") or os.path.join(_HERE, "libmollyb200.so")  # override: tuning variants

MB_LJ, MB_COULOMB, MB_CRF, MB_EWALD_REAL = 0, 1, 2, 3
MB_CUT_NONE, MB_CUT_DISTANCE, MB_CUT_SHIFTED_POTENTIAL, MB_CUT_SHIFTED_FORCE = 0, 1, 2, 3
MB_CUT_CUBIC_SPLINE, MB_CUT_POLYNOMIAL = 4, 5
MB_MIX_LORENTZ, MB_MIX_GEOMETRIC = 0, 1
MB_OK, MB_ERR_INVALID, MB_ERR_CUDA, MB_ERR_CAPACITY, MB_ERR_STATE, MB_ERR_NOGPU = 0, -1, -2, -3, -4, -5

EXPORTED = [
    "mb_last_error", "mb_device_count", "mb_ctx_create", "mb_ctx_destroy", "mb_set_atoms", "mb_set_atoms_soa",
    "mb_set_box", "mb_set_inters", "mb_set_exceptions", "mb_set_neighbor_policy", "mb_forces", "mb_energy",
    "mb_forces_energy", "mb_simulate_vv", "mb_remove_cm_motion", "mb_kinetic_energy", "mb_rebuild_neighbors",
    "mb_stats", "mb_synchronize", "mb_set_capacity_scale", "mb_set_launch_config", "mb_comm_unique_id",
    "mb_comm_init", "mb_decomp_plan", "mb_set_profiling", "mb_set_specific", "mb_forces_energy_all", "mb_set_pme", "mb_pme_plan",
    "mb_set_lj_dispersion_correction", "mb_random_velocities", "mb_kinetic_energy_tensor", "mb_set_box_triclinic",
]


class MBInter(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("cutoff_kind", C.c_int32), ("r_cut", C.c_double), ("r_act", C.c_double),
        ("weight_special", C.c_double), ("coulomb_const", C.c_double), ("solvent_dielectric", C.c_double),
        ("ewald_alpha", C.c_double), ("sigma_mix", C.c_int32), ("eps_mix", C.c_int32), ("approx_erfc", C.c_int32),
        ("use_neighbors", C.c_int32),
    ]


class MBStats(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int64), ("n_rebuilds", C.c_int64), ("n_force_evals", C.c_int64), ("n_steps", C.c_int64),
        ("n_list_entries", C.c_int64), ("n_pairs_in_list", C.c_int64), ("n_bricks", C.c_int32),
        ("n_cells", C.c_int32 * 3), ("brick_dims", C.c_int32 * 3), ("halo_capacity", C.c_int32),
        ("list_stride", C.c_int32), ("max_neighbors", C.c_int32), ("max_halo", C.c_int32), ("path", C.c_int32),
        ("violations", C.c_int32), ("r_list", C.c_double), ("kernel_launches", C.c_int64),
        ("force_ms", C.c_double), ("vv_ms", C.c_double), ("rebuild_ms", C.c_double),
        ("force_launches", C.c_int64), ("vv_launches", C.c_int64), ("rebuild_launches", C.c_int64),
        ("graph_mode", C.c_int32), ("n_prunes", C.c_int32), ("peer_transport", C.c_int32), ("reserved_", C.c_int32),
    ]


class MBVVParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("n_steps", C.c_int64), ("init_step", C.c_int64), ("remove_cm_every", C.c_int32),
        ("andersen_kT", C.c_double), ("andersen_prob", C.c_double), ("rng_ctr1", C.c_uint64), ("rng_key", C.c_uint64),
    ]


class MollyB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmollyb200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libmollyb200.so. Fails loudly if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "mollyb200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.mb_last_error.restype = C.c_char_p
    L.mb_device_count.restype = C.c_int
    L.mb_ctx_create.argtypes = [C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.mb_ctx_destroy.argtypes = [vp]
    L.mb_ctx_destroy.restype = None
    L.mb_set_atoms.argtypes = [vp, i64, vp]
    L.mb_set_atoms_soa.argtypes = [vp, i64, vp, vp, vp, vp]
    L.mb_set_box.argtypes = [vp, C.POINTER(dbl)]
    L.mb_set_inters.argtypes = [vp, C.c_int, C.POINTER(MBInter)]
    L.mb_set_exceptions.argtypes = [vp, i64, vp, vp, i64, vp, vp]
    L.mb_set_neighbor_policy.argtypes = [vp, dbl, C.c_int]
    L.mb_forces.argtypes = [vp, vp, vp, vp, i64]
    L.mb_energy.argtypes = [vp, vp, vp, i64]
    L.mb_forces_energy.argtypes = [vp, vp, vp, vp, vp, i64]
    L.mb_simulate_vv.argtypes = [vp, vp, vp, C.POINTER(MBVVParams)]
    L.mb_remove_cm_motion.argtypes = [vp, vp]
    L.mb_kinetic_energy.argtypes = [vp, vp, C.POINTER(dbl)]
    L.mb_rebuild_neighbors.argtypes = [vp, vp]
    L.mb_stats.argtypes = [vp, C.POINTER(MBStats)]
    L.mb_synchronize.argtypes = [vp]
    L.mb_set_capacity_scale.argtypes = [vp, dbl]
    L.mb_set_launch_config.argtypes = [vp, C.POINTER(i32), i32]
    L.mb_comm_unique_id.argtypes = [vp]
    L.mb_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.mb_set_specific.argtypes = [vp, C.c_int, i64, vp, vp]
    L.mb_set_pme.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_double, i64, vp, vp]
    L.mb_pme_plan.argtypes = [vp, C.c_double, C.c_double, C.c_int, vp, vp, vp, C.c_int]
    L.mb_forces_energy_all.argtypes = [vp, vp, vp, vp, i64]
    L.mb_set_lj_dispersion_correction.argtypes = [vp, dbl]
    L.mb_random_velocities.argtypes = [vp, vp, dbl, C.c_uint64, C.c_uint64]
    L.mb_kinetic_energy_tensor.argtypes = [vp, vp, C.POINTER(dbl)]
    L.mb_decomp_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(i32), vp, C.POINTER(i32), C.c_int]
    L.mb_set_profiling.argtypes = [vp, C.c_int]
    for name in EXPORTED:
        fn = getattr(L, name)
        if name not in ("mb_last_error", "mb_ctx_destroy", "mb_device_count"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != MB_OK:
        raise MollyB200Error(rc, load().mb_last_error().decode())
