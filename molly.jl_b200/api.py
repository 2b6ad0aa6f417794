"""Host-side mirror of Molly.jl's API for the non-bonded + VelocityVerlet hot path.

Same names and argument meaning as the reference so the parity tests read like the reference's
own tests (Julia is not available in the build image; the Julia shim that binds the same C ABI is
julia/MollyB200Ext.jl):

    Atom, CubicBoundary, System                     src/types.jl:466-475, src/spatial.jl:40, src/types.jl:795-979
    NoCutoff, DistanceCutoff, Shifted*Cutoff        src/cutoffs.jl:47-190
    LennardJones, Coulomb, CoulombReactionField     src/interactions/lennard_jones.jl:28-35, coulomb.jl:32-70, :698-747
    GPUNeighborFinder (+ aliases)                   src/neighbors.jl:104-115
    VelocityVerlet, AndersenThermostat, simulate    src/simulators.jl:287-295, :547-668; src/coupling.jl:184-212
    forces, forces_virial, potential_energy         src/force.jl:678-720, src/energy.jl:202-248
    kinetic_energy, temperature, remove_CM_motion   src/energy.jl:44-175, src/spatial.jl:901-929

Everything numerical happens in libmollyb200.so on the GPU; this module only marshals arrays.
Units are Molly's (nm, ps, g/mol, kJ/mol) with the Unitful wrappers stripped.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi as capi
from ._capi import MollyB200Error  # noqa: F401

COULOMB_CONST = 138.93545764  # src/interactions/coulomb.jl:16
BOLTZMANN_K = 8.31446261815324e-3  # src/units.jl:186-198


# ------------------------------------------------------------------------------------------------
# types
# ------------------------------------------------------------------------------------------------
@dataclass
class Atom:
    index: int = 1
    atom_type: int = 1
    mass: float = 1.0
    charge: float = 0.0
    sigma: float = 0.0
    eps: float = 0.0
    lam: float = 1.0
    alch_role: int = 0


def atom_dtype(dtype) -> np.dtype:
    """numpy struct dtype with Molly's bits layout Atom{Int32,T,T,T,T,T} (32 B f32 / 56 B f64)."""
    f = np.dtype(dtype)
    return np.dtype([("index", np.int32), ("atom_type", np.int32), ("mass", f), ("charge", f), ("sigma", f),
                     ("eps", f), ("lam", f), ("alch_role", np.int32)], align=True)


def atoms_to_array(atoms: Sequence[Atom], dtype) -> np.ndarray:
    arr = np.zeros(len(atoms), atom_dtype(dtype))
    for k, a in enumerate(atoms):
        arr[k] = (a.index, a.atom_type, a.mass, a.charge, a.sigma, a.eps, a.lam, a.alch_role)
    return arr


def atoms_from_arrays(mass, charge, sigma, eps, dtype) -> np.ndarray:
    n = len(mass)
    arr = np.zeros(n, atom_dtype(dtype))
    arr["index"] = np.arange(1, n + 1)
    arr["atom_type"] = 1
    arr["mass"], arr["charge"], arr["sigma"], arr["eps"], arr["lam"] = mass, charge, sigma, eps, 1.0
    return arr


@dataclass
class CubicBoundary:
    x: float
    y: Optional[float] = None
    z: Optional[float] = None

    def __post_init__(self):
        if self.y is None:
            self.y = self.x
        if self.z is None:
            self.z = self.x

    @property
    def side_lengths(self):
        return np.array([self.x, self.y, self.z], np.float64)


class TriclinicBoundary:
    """TriclinicBoundary(bv1, bv2, bv3) — src/spatial.jl:151-215 (approx_images = true): lower-triangular basis vectors.
    Systems in such a box run on the no-list kernel."""

    def __init__(self, bv1, bv2, bv3):
        self.basis_vectors = np.array([bv1, bv2, bv3], np.float64).reshape(3, 3)
        bv = self.basis_vectors
        if not (bv[0, 0] > 0 and bv[0, 1] == 0 and bv[0, 2] == 0):
            raise ValueError("first basis vector must be along the x-axis with a positive x component")
        if not (bv[1, 1] > 0 and bv[1, 2] == 0):
            raise ValueError("second basis vector must be in the xy plane with a positive y component")
        if not bv[2, 2] > 0:
            raise ValueError("third basis vector must have a positive z component")

    @property
    def side_lengths(self):  # heights (what the engine's geometry code sees)
        return np.array([self.basis_vectors[0, 0], self.basis_vectors[1, 1], self.basis_vectors[2, 2]], np.float64)


@dataclass
class NoCutoff:
    pass


@dataclass
class DistanceCutoff:
    dist_cutoff: float


@dataclass
class ShiftedPotentialCutoff:
    dist_cutoff: float


@dataclass
class ShiftedForceCutoff:
    dist_cutoff: float


@dataclass
class CubicSplineCutoff:
    """CubicSplineCutoff(dist_activation, dist_cutoff) — src/cutoffs.jl:174-215."""
    dist_activation: float
    dist_cutoff: float

    def __post_init__(self):
        if self.dist_cutoff <= self.dist_activation:  # ArgumentError in the reference constructor (:181-187)
            raise ValueError(f"the cutoff radius {self.dist_cutoff} must be larger than the activation radius {self.dist_activation}")


@dataclass
class PolynomialCutoff:
    """PolynomialCutoff(dist_activation, dist_cutoff) — src/cutoffs.jl:217-253 (OpenMM's switching function)."""
    dist_activation: float
    dist_cutoff: float

    def __post_init__(self):
        if self.dist_cutoff <= self.dist_activation:
            raise ValueError(f"the cutoff radius {self.dist_cutoff} must be larger than the activation radius {self.dist_activation}")


def _cutoff_kind(c):
    """-> (kind, dist_cutoff, dist_activation)"""
    if isinstance(c, NoCutoff):
        return capi.MB_CUT_NONE, 0.0, 0.0
    if isinstance(c, DistanceCutoff):
        return capi.MB_CUT_DISTANCE, c.dist_cutoff, 0.0
    if isinstance(c, ShiftedPotentialCutoff):
        return capi.MB_CUT_SHIFTED_POTENTIAL, c.dist_cutoff, 0.0
    if isinstance(c, ShiftedForceCutoff):
        return capi.MB_CUT_SHIFTED_FORCE, c.dist_cutoff, 0.0
    if isinstance(c, CubicSplineCutoff):
        return capi.MB_CUT_CUBIC_SPLINE, c.dist_cutoff, c.dist_activation
    if isinstance(c, PolynomialCutoff):
        return capi.MB_CUT_POLYNOMIAL, c.dist_cutoff, c.dist_activation
    raise TypeError(f"unsupported cutoff {c!r}")


@dataclass
class LennardJones:
    cutoff: object = field(default_factory=NoCutoff)
    use_neighbors: bool = False
    weight_special: float = 1.0
    sigma_mixing: str = "lorentz"
    eps_mixing: str = "geometric"

    def descriptor(self):
        k, rc, ra = _cutoff_kind(self.cutoff)
        return capi.MBInter(capi.MB_LJ, k, rc, ra, self.weight_special, COULOMB_CONST, 1.0, 0.0,
                            capi.MB_MIX_GEOMETRIC if self.sigma_mixing == "geometric" else capi.MB_MIX_LORENTZ,
                            capi.MB_MIX_GEOMETRIC if self.eps_mixing == "geometric" else capi.MB_MIX_LORENTZ, 0,
                            int(self.use_neighbors))


@dataclass
class Coulomb:
    cutoff: object = field(default_factory=NoCutoff)
    use_neighbors: bool = False
    weight_special: float = 1.0
    coulomb_const: float = COULOMB_CONST

    def descriptor(self):
        k, rc, ra = _cutoff_kind(self.cutoff)
        return capi.MBInter(capi.MB_COULOMB, k, rc, ra, self.weight_special, self.coulomb_const, 1.0, 0.0, 0, 1, 0,
                            int(self.use_neighbors))


@dataclass
class CoulombReactionField:
    dist_cutoff: float
    solvent_dielectric: float = 78.3  # coulomb.jl:676
    use_neighbors: bool = False
    weight_special: float = 1.0
    coulomb_const: float = COULOMB_CONST

    def descriptor(self):
        return capi.MBInter(capi.MB_CRF, capi.MB_CUT_DISTANCE, self.dist_cutoff, 0.0, self.weight_special,
                            self.coulomb_const, self.solvent_dielectric, 0.0, 0, 1, 0, int(self.use_neighbors))


@dataclass
class CoulombEwald:
    """Real-space part of Ewald/PME (coulomb.jl:1320-1441); alpha = sqrt(-log(2 tol)) / dist_cutoff.
    approximate_erfc=True (the reference's default, :1331) evaluates erfc with calc_erfc's polynomial (:1384-1393)."""
    dist_cutoff: float
    error_tol: float = 5e-4
    use_neighbors: bool = False
    weight_special: float = 1.0
    coulomb_const: float = COULOMB_CONST
    approximate_erfc: bool = True

    def descriptor(self):
        alpha = np.sqrt(-np.log(2 * self.error_tol)) / self.dist_cutoff
        return capi.MBInter(capi.MB_EWALD_REAL, capi.MB_CUT_DISTANCE, self.dist_cutoff, 0.0, self.weight_special,
                            self.coulomb_const, 1.0, float(alpha), 0, 1, int(self.approximate_erfc), int(self.use_neighbors))


def _pairs_from(obj, n, want_true: bool):
    """Accept a dense (n,n) bool matrix or an (m,2) array of 1-based pairs."""
    if obj is None:
        return np.zeros((0, 2), np.int32)
    a = np.asarray(obj)
    if a.ndim == 2 and a.shape == (n, n) and a.dtype == np.bool_:
        m = a if want_true else ~a
        i, j = np.nonzero(np.triu(m, 1) | np.triu(m.T, 1))
        return np.stack([i + 1, j + 1], 1).astype(np.int32)
    return np.ascontiguousarray(a, np.int32).reshape(-1, 2)


@dataclass
class GPUNeighborFinder:
    """Device neighbour finder. `eligible`/`special` may be dense bool matrices (as in the reference
    constructor) or `excluded_pairs`/`special_pairs` sparse 1-based (m,2) arrays (neighbors.jl:104-115).
    n_steps: rebuild interval (reference default 10 CPU / 25 GPU); 0 = displacement-triggered (exact)."""
    dist_cutoff: float = 0.0
    eligible: object = None
    special: object = None
    excluded_pairs: object = None
    special_pairs: object = None
    n_steps: int = 0


# the reference's other finders build the same pair set; here they all map to the device cell list
DistanceNeighborFinder = GPUNeighborFinder
CellListMapNeighborFinder = GPUNeighborFinder
TreeNeighborFinder = GPUNeighborFinder


@dataclass
class PME:
    """PME(dist_cutoff, atoms, boundary; error_tol=0.0005, order=5, ϵr=1.0) general interaction
    (src/interactions/ewald.jl:363-421) together with the EwaldExclusion list that src/setup.jl:1903-1912 builds from
    find_excluded_pairs(eligible, special): `excluded_pairs` here = the excluded OR special pairs, 1-based (m,2).
    Use with CoulombEwald(dist_cutoff, error_tol) as the pairwise interaction. First implementation: see
    include/mollyb200.h (mb_set_pme) for its validation status."""
    dist_cutoff: float
    error_tol: float = 0.0005
    order: int = 5
    eps_r: float = 1.0
    excluded_pairs: object = None


@dataclass
class LJDispersionCorrection:
    """LJDispersionCorrection(atoms, dist_cutoff) general interaction (src/interactions/lennard_jones.jl:163-275): the
    factors are computed by the library from the System's atoms."""
    dist_cutoff: float


@dataclass
class InteractionList2Atoms:
    """InteractionList2Atoms of HarmonicBond (src/types.jl:89-157, interactions/harmonic_bond.jl): 1-based is/js,
    per-term k (kJ mol^-1 nm^-2) and r0 (nm)."""
    is_: object
    js: object
    k: object
    r0: object
    kind = 0

    def arrays(self):
        idx = np.stack([np.asarray(self.is_), np.asarray(self.js)], 1).astype(np.int32)
        par = np.stack([np.asarray(self.k, np.float64), np.asarray(self.r0, np.float64)], 1)
        return np.ascontiguousarray(idx), np.ascontiguousarray(par)


@dataclass
class InteractionList3Atoms:
    """InteractionList3Atoms of HarmonicAngle (interactions/harmonic_angle.jl): k (kJ mol^-1 rad^-2), theta0 (rad)."""
    is_: object
    js: object
    ks: object
    k: object
    theta0: object
    kind = 1

    def arrays(self):
        idx = np.stack([np.asarray(self.is_), np.asarray(self.js), np.asarray(self.ks)], 1).astype(np.int32)
        par = np.stack([np.asarray(self.k, np.float64), np.asarray(self.theta0, np.float64)], 1)
        return np.ascontiguousarray(idx), np.ascontiguousarray(par)


@dataclass
class InteractionList4Atoms:
    """InteractionList4Atoms of PeriodicTorsion (interactions/periodic_torsion.jl), flattened to one
    (periodicity, phase, k) term per entry; propers and impropers use the same list type."""
    is_: object
    js: object
    ks: object
    ls: object
    periodicity: object
    phase: object
    k: object
    kind = 2

    def arrays(self):
        idx = np.stack([np.asarray(self.is_), np.asarray(self.js), np.asarray(self.ks), np.asarray(self.ls)], 1).astype(np.int32)
        par = np.stack([np.asarray(self.periodicity, np.float64), np.asarray(self.phase, np.float64),
                        np.asarray(self.k, np.float64)], 1)
        return np.ascontiguousarray(idx), np.ascontiguousarray(par)


@dataclass
class AndersenThermostat:
    temperature: float
    coupling_const: float


@dataclass
class VelocityVerlet:
    dt: float
    coupling: object = None
    remove_CM_motion: int = 1


# ------------------------------------------------------------------------------------------------
# System
# ------------------------------------------------------------------------------------------------
class System:
    """System(atoms, coords, boundary, velocities, pairwise_inters, neighbor_finder) — src/types.jl:795-979.

    coords / velocities are numpy arrays (n,3) of `dtype` (host) or torch CUDA tensors (device); the
    engine accepts both through the same C entry points.
    """

    def __init__(self, atoms, coords, boundary, velocities=None, pairwise_inters=(), neighbor_finder=None,
                 dtype=np.float32, device: int = 0, k=BOLTZMANN_K, specific_inter_lists=(), general_inters=()):
        self.dtype = np.dtype(dtype)
        if isinstance(atoms, np.ndarray) and atoms.dtype.names:
            self.atoms = np.ascontiguousarray(atoms.astype(atom_dtype(self.dtype)))
        else:
            self.atoms = atoms_to_array(atoms, self.dtype)
        self.n = len(self.atoms)
        self.boundary = boundary
        self.coords = self._as_state(coords)
        self.velocities = self._as_state(velocities if velocities is not None else np.zeros((self.n, 3)))
        self.pairwise_inters = tuple(pairwise_inters)
        self.neighbor_finder = neighbor_finder
        self.specific_inter_lists = tuple(specific_inter_lists)
        self.general_inters = tuple(general_inters)
        self.device = device
        self.k = k
        self._ctx = None

    def _as_state(self, a):
        if hasattr(a, "data_ptr"):  # torch tensor (device resident)
            return a
        return np.ascontiguousarray(np.asarray(a, self.dtype).reshape(self.n, 3))

    @property
    def masses(self):
        return self.atoms["mass"].astype(np.float64)

    # ---- engine context ------------------------------------------------------------------------
    def engine(self):
        if self._ctx is None:
            L = capi.load()
            ctx = C.c_void_p()
            capi.check(L.mb_ctx_create(self.device, 32 if self.dtype == np.float32 else 64, None, C.byref(ctx)))
            self._ctx = ctx
            self._L = L
            self._configure()
        return self._ctx

    def _configure(self):
        L, ctx = self._L, self._ctx
        capi.check(L.mb_set_atoms(ctx, self.n, self.atoms.ctypes.data))
        if isinstance(self.boundary, TriclinicBoundary):
            capi.check(L.mb_set_box_triclinic(ctx, (C.c_double * 9)(*self.boundary.basis_vectors.ravel())))
        else:
            side = (C.c_double * 3)(*self.boundary.side_lengths)
            capi.check(L.mb_set_box(ctx, side))
        descs = [it.descriptor() for it in self.pairwise_inters]
        arr = (capi.MBInter * max(1, len(descs)))(*descs)
        capi.check(L.mb_set_inters(ctx, len(descs), arr))
        nf = self.neighbor_finder
        if nf is not None:
            if nf.excluded_pairs is not None:
                ex = _pairs_from(nf.excluded_pairs, self.n, want_true=True)
            else:
                ex = _pairs_from(nf.eligible, self.n, want_true=False)
            sp = _pairs_from(nf.special_pairs if nf.special_pairs is not None else nf.special, self.n, want_true=True)
            ei, ej = np.ascontiguousarray(ex[:, 0]), np.ascontiguousarray(ex[:, 1])
            si, sj = np.ascontiguousarray(sp[:, 0]), np.ascontiguousarray(sp[:, 1])
            self._keep = (ei, ej, si, sj)
            capi.check(L.mb_set_exceptions(ctx, len(ei), ei.ctypes.data, ej.ctypes.data, len(si), si.ctypes.data,
                                           sj.ctypes.data))
            capi.check(L.mb_set_neighbor_policy(ctx, float(nf.dist_cutoff), int(nf.n_steps)))
        # specific interaction lists: entries of the same kind are concatenated (e.g. propers + impropers)
        by_kind = {}
        for sil in self.specific_inter_lists:
            idx, par = sil.arrays()
            a, b = by_kind.get(sil.kind, (None, None))
            by_kind[sil.kind] = (idx if a is None else np.concatenate([a, idx]), par if b is None else np.concatenate([b, par]))
        for kind, (idx, par) in by_kind.items():
            idx, par = np.ascontiguousarray(idx), np.ascontiguousarray(par)
            capi.check(L.mb_set_specific(ctx, kind, len(idx), idx.ctypes.data, par.ctypes.data))

        for gi in self.general_inters:
            if isinstance(gi, LJDispersionCorrection):
                capi.check(L.mb_set_lj_dispersion_correction(ctx, float(gi.dist_cutoff)))
                continue
            if not isinstance(gi, PME):
                raise ValueError("only PME and LJDispersionCorrection are supported as general interactions")
            pairs = np.zeros((0, 2), np.int32) if gi.excluded_pairs is None else np.asarray(gi.excluded_pairs, np.int32).reshape(-1, 2)
            pi, pj = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
            self._keep_pme = (pi, pj)
            capi.check(L.mb_set_pme(ctx, float(gi.dist_cutoff), float(gi.error_tol), int(gi.order), float(gi.eps_r), len(pi),
                                    pi.ctypes.data, pj.ctypes.data))

    def close(self):
        if self._ctx is not None:
            self._L.mb_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self) -> dict:
        st = capi.MBStats()
        ctx = self.engine()
        capi.check(self._L.mb_stats(ctx, C.byref(st)))
        out = {}
        for name, _ in capi.MBStats._fields_:
            v = getattr(st, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out

    def set_profiling(self, enable: bool):
        ctx = self.engine()
        capi.check(self._L.mb_set_profiling(ctx, int(enable)))

    def set_launch_config(self, brick_dims=(0, 0, 0), lanes_per_atom=0):
        bd = (C.c_int32 * 3)(*brick_dims)
        ctx = self.engine()
        capi.check(self._L.mb_set_launch_config(ctx, bd, lanes_per_atom))


def _ptr(a):
    return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data


# ------------------------------------------------------------------------------------------------
# functions
# ------------------------------------------------------------------------------------------------
def forces(sys: System, neighbors=None, step_n: int = 0) -> np.ndarray:
    """forces(sys[, neighbors, step_n]) — src/force.jl:678-687. Returns (n,3) in kJ mol^-1 nm^-1."""
    ctx = sys.engine()
    fs = np.zeros((sys.n, 3), sys.dtype)
    if sys.specific_inter_lists or sys.general_inters:  # forces(sys) sums pairwise + specific + general interactions
        capi.check(sys._L.mb_forces_energy_all(ctx, _ptr(sys.coords), fs.ctypes.data, None, step_n))
    else:
        capi.check(sys._L.mb_forces(ctx, _ptr(sys.coords), fs.ctypes.data, None, step_n))
    return fs


def forces_virial(sys: System, neighbors=None, step_n: int = 0):
    """forces_virial — src/force.jl:703-720. Returns (forces, virial 3x3). Pairwise interactions only: a System with
    specific or general interactions is refused instead of silently dropping their virial."""
    if sys.specific_inter_lists or sys.general_inters:
        raise NotImplementedError("forces_virial covers the pairwise seam (pairwise_forces_loop_gpu!) only")
    ctx = sys.engine()
    fs = np.zeros((sys.n, 3), sys.dtype)
    vir = np.zeros(9, sys.dtype)
    capi.check(sys._L.mb_forces(ctx, _ptr(sys.coords), fs.ctypes.data, vir.ctypes.data, step_n))
    return fs, vir.reshape(3, 3).T.copy()


def potential_energy(sys: System, neighbors=None, step_n: int = 0) -> float:
    """potential_energy(sys[, neighbors, step_n]) — src/energy.jl:202-248 (pairwise part)."""
    ctx = sys.engine()
    pe = np.zeros(1, sys.dtype)
    if sys.specific_inter_lists or sys.general_inters:  # potential_energy(sys): pairwise + specific + general
        capi.check(sys._L.mb_forces_energy_all(ctx, _ptr(sys.coords), None, pe.ctypes.data, step_n))
    else:
        capi.check(sys._L.mb_energy(ctx, _ptr(sys.coords), pe.ctypes.data, step_n))
    return float(pe[0])


def forces_energy(sys: System, step_n: int = 0):
    """forces(sys) and potential_energy(sys) of all pairwise + specific interactions in one traversal."""
    ctx = sys.engine()
    fs = np.zeros((sys.n, 3), sys.dtype)
    pe = np.zeros(1, sys.dtype)
    if sys.specific_inter_lists or sys.general_inters:
        capi.check(sys._L.mb_forces_energy_all(ctx, _ptr(sys.coords), fs.ctypes.data, pe.ctypes.data, step_n))
    else:
        capi.check(sys._L.mb_forces_energy(ctx, _ptr(sys.coords), fs.ctypes.data, pe.ctypes.data, None, step_n))
    return fs, float(pe[0])


def find_neighbors(sys: System, *args, **kwargs):
    """find_neighbors(sys, nf::GPUNeighborFinder, ...) = nothing in the reference (neighbors.jl:364);
    here it forces a device rebuild and returns None."""
    ctx = sys.engine()
    capi.check(sys._L.mb_rebuild_neighbors(ctx, _ptr(sys.coords)))
    return None


def simulate(sys: System, sim: VelocityVerlet, n_steps: int, init_step: int = 0, rng=None, max_retries: int = 2):
    """simulate!(sys, sim::VelocityVerlet, n_steps) — src/simulators.jl:547-668. Mutates sys.coords / velocities."""
    ctx = sys.engine()
    p = capi.MBVVParams()
    p.dt = float(sim.dt)
    p.n_steps = int(n_steps)
    p.init_step = int(init_step)
    p.remove_cm_every = int(sim.remove_CM_motion)
    p.andersen_kT = 0.0
    p.andersen_prob = 0.0
    couplings = sim.coupling if isinstance(sim.coupling, (tuple, list)) else ((sim.coupling,) if sim.coupling else ())
    for c in couplings:
        if isinstance(c, AndersenThermostat):
            p.andersen_kT = sys.k * c.temperature
            p.andersen_prob = sim.dt / c.coupling_const
        else:
            raise TypeError(f"unsupported coupling {c!r} (the stock Molly path handles it)")
    rng = rng or np.random.default_rng()
    p.rng_ctr1 = int(rng.integers(0, 2 ** 63))
    p.rng_key = int(rng.integers(0, 2 ** 63))
    host = not hasattr(sys.coords, "data_ptr")
    backup = (sys.coords.copy(), sys.velocities.copy()) if host else None
    scale = 1.0
    for attempt in range(max_retries + 1):
        rc = sys._L.mb_simulate_vv(ctx, _ptr(sys.coords), _ptr(sys.velocities), C.byref(p))
        if rc == capi.MB_ERR_CAPACITY and backup is not None and attempt < max_retries:
            sys.coords[...], sys.velocities[...] = backup
            scale *= 2.0
            capi.check(sys._L.mb_set_capacity_scale(ctx, scale))
            continue
        capi.check(rc)
        break
    return sys


def kinetic_energy(sys: System) -> float:
    out = C.c_double(0.0)
    ctx = sys.engine()
    capi.check(sys._L.mb_kinetic_energy(ctx, _ptr(sys.velocities), C.byref(out)))
    return out.value


def temperature(sys: System) -> float:
    """src/energy.jl:158-175 with df = 3N - 3 for a periodic 3-D box."""
    df = 3 * sys.n - 3
    return 2.0 * kinetic_energy(sys) / (df * sys.k)


def remove_CM_motion(sys: System):
    ctx = sys.engine()
    capi.check(sys._L.mb_remove_cm_motion(ctx, _ptr(sys.velocities)))
    return sys


def random_velocities(sys: System, temp: float, rng=None) -> np.ndarray:
    """random_velocities(sys, temp; rng) — src/spatial.jl:803-831: Maxwell-Boltzmann velocities drawn on the device
    (Philox4x32-10 keyed by two 64-bit draws of `rng`, like the reference's GPU kernel src/kernels.jl:688-703)."""
    rng = rng or np.random.default_rng()
    ctx = sys.engine()
    out = np.zeros((sys.n, 3), sys.dtype)
    capi.check(sys._L.mb_random_velocities(ctx, out.ctypes.data, float(sys.k * temp), int(rng.integers(0, 2 ** 63)),
                                           int(rng.integers(0, 2 ** 63))))
    return out


def random_velocities_(sys: System, temp: float, rng=None) -> System:
    """random_velocities!(sys, temp): in place."""
    v = random_velocities(sys, temp, rng)
    if hasattr(sys.velocities, "data_ptr"):
        import torch
        sys.velocities.copy_(torch.from_numpy(v))
    else:
        sys.velocities[...] = v
    return sys


def kinetic_energy_tensor(sys: System) -> np.ndarray:
    """K = 1/2 sum m v (x) v — src/energy.jl:56-70 (3x3, kJ/mol)."""
    out = (C.c_double * 9)()
    ctx = sys.engine()
    capi.check(sys._L.mb_kinetic_energy_tensor(ctx, _ptr(sys.velocities), out))
    return np.array(out[:], np.float64).reshape(3, 3)


def wrap_coords(coords, boundary: CubicBoundary):
    """wrap_coords — src/spatial.jl:573-586 (host helper for test set-up)."""
    L = boundary.side_lengths.astype(coords.dtype)
    return coords - np.floor(coords / L) * L


def comm_unique_id() -> bytes:
    """ncclUniqueId (128 bytes) created on the calling rank; broadcast it to the other ranks with the host runtime."""
    buf = C.create_string_buffer(128)
    capi.check(capi.load().mb_comm_unique_id(buf))
    return buf.raw


def comm_init(sys: System, unique_id: bytes, rank: int, nranks: int):
    """Join the spatial decomposition: z-slabs of cell layers, NCCL halo exchange (see include/mollyb200.h)."""
    ctx = sys.engine()
    capi.check(sys._L.mb_comm_init(ctx, C.c_char_p(unique_id), int(rank), int(nranks)))


def decomp_plan(ncz: int, halo_layers: int, nranks: int, rank: int, layer_start):
    """Host-side halo-exchange plan: (send, recv) lists of (peer, first slot, slot count)."""
    ls = np.ascontiguousarray(layer_start, np.int32)
    cap = 4 * nranks + 8
    snd = np.zeros((cap, 3), np.int32)
    rcv = np.zeros((cap, 3), np.int32)
    ns, nr = C.c_int32(0), C.c_int32(0)
    capi.check(capi.load().mb_decomp_plan(ncz, halo_layers, nranks, rank, ls.ctypes.data, snd.ctypes.data, C.byref(ns),
                                          rcv.ctypes.data, C.byref(nr), cap))
    return snd[:ns.value].tolist(), rcv[:nr.value].tolist()


def device_count() -> int:
    return int(capi.load().mb_device_count())
