"""ctypes binding of the C oracle (oracle/molly_oracle.c) + small numpy helpers.

TEST INFRASTRUCTURE ONLY — see oracle/README.md. Reference citations live in
molly_oracle_impl.h next to each function.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmollyoracle.so")

LJ, COULOMB, CRF, EWALD_REAL = 0, 1, 2, 3
CUT_NONE, CUT_DISTANCE, CUT_SHIFTED_POTENTIAL, CUT_SHIFTED_FORCE = 0, 1, 2, 3
CUT_CUBIC_SPLINE, CUT_POLYNOMIAL = 4, 5  # two-point cutoffs (r_act, r_cut); oracle only so far
MIX_LORENTZ, MIX_GEOMETRIC = 0, 1
COULOMB_CONST = 138.93545764  # src/interactions/coulomb.jl:16
BOLTZMANN_K = 8.31446261815324e-3  # src/units.jl:186-198 (kJ mol^-1 K^-1)


class InterC(C.Structure):
    """orc_inter_t == mb_inter_t (include/mollyb200.h)."""

    _fields_ = [
        ("kind", C.c_int32),
        ("cutoff_kind", C.c_int32),
        ("r_cut", C.c_double),
        ("r_act", C.c_double),
        ("weight_special", C.c_double),
        ("coulomb_const", C.c_double),
        ("solvent_dielectric", C.c_double),
        ("ewald_alpha", C.c_double),
        ("sigma_mix", C.c_int32),
        ("eps_mix", C.c_int32),
        ("approx_erfc", C.c_int32),
        ("use_neighbors", C.c_int32),
    ]


class NLEntry(C.Structure):
    _fields_ = [("i", C.c_int32), ("j", C.c_int32), ("special", C.c_int32)]


class SystemC(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int64),
        ("dtype", C.c_int32),
        ("n_inters", C.c_int32),
        ("box", C.c_double * 3),
        ("mass", C.c_void_p),
        ("charge", C.c_void_p),
        ("sigma", C.c_void_p),
        ("eps", C.c_void_p),
        ("lam", C.c_void_p),
        ("inters", C.POINTER(InterC)),
        ("excl_ptr", C.c_void_p),
        ("excl_idx", C.c_void_p),
        ("spec_ptr", C.c_void_p),
        ("spec_idx", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc). Building the checker is not using it."""
    # make decides staleness (sources newer than the library); a missing make with a prebuilt library is fine
    if force and os.path.exists(_LIB_PATH):
        os.remove(_LIB_PATH)
    try:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    except (OSError, subprocess.CalledProcessError):
        if not os.path.exists(_LIB_PATH):
            raise
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_forces_allpairs.restype = C.c_int
        L.orc_forces_allpairs.argtypes = [C.POINTER(SystemC), C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.c_int]
        L.orc_neighbor_list.restype = C.c_int64
        L.orc_neighbor_list.argtypes = [C.POINTER(SystemC), C.c_void_p, C.c_double,
                                        C.POINTER(C.POINTER(NLEntry))]
        L.orc_forces_nl.restype = C.c_int
        L.orc_forces_nl.argtypes = [C.POINTER(SystemC), C.c_void_p, C.POINTER(NLEntry), C.c_int64, C.c_void_p,
                                    C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.orc_simulate_vv.restype = C.c_int
        L.orc_simulate_vv.argtypes = [C.POINTER(SystemC), C.c_void_p, C.c_void_p, C.c_double, C.c_int64,
                                      C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.orc_remove_cm.restype = None
        L.orc_remove_cm.argtypes = [C.POINTER(SystemC), C.c_void_p]
        L.orc_vector_1D.restype = C.c_double
        L.orc_vector_1D.argtypes = [C.c_double] * 3
        L.orc_wrap_coord_1D.restype = C.c_double
        L.orc_wrap_coord_1D.argtypes = [C.c_double] * 2
        L.orc_free.restype = None
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


DEFAULT_THREADS = 0  # 0 = whatever OpenMP reports; bench.py sets it (launchers like torchrun export OMP_NUM_THREADS=1)


def max_threads() -> int:
    return int(DEFAULT_THREADS) if DEFAULT_THREADS else int(lib().orc_max_threads())


@dataclass
class Inter:
    """One pairwise interaction (LennardJones / Coulomb / CoulombReactionField / CoulombEwald real part)."""

    kind: int
    cutoff_kind: int = CUT_NONE
    r_cut: float = 0.0
    r_act: float = 0.0  # dist_activation of the two-point cutoffs
    weight_special: float = 1.0
    coulomb_const: float = COULOMB_CONST
    solvent_dielectric: float = 78.3  # coulomb.jl:676
    ewald_alpha: float = 0.0
    sigma_mix: int = MIX_LORENTZ
    eps_mix: int = MIX_GEOMETRIC
    use_neighbors: bool = False
    approx_erfc: bool = False  # CoulombEwald(approximate_erfc=...), coulomb.jl:1331 (reference default: true)

    def to_c(self) -> InterC:
        return InterC(self.kind, self.cutoff_kind, self.r_cut, self.r_act, self.weight_special, self.coulomb_const,
                      self.solvent_dielectric, self.ewald_alpha, self.sigma_mix, self.eps_mix, int(self.approx_erfc),
                      int(self.use_neighbors))


def pairs_to_csr(n: int, pairs: np.ndarray):
    """Unordered 0-based pairs (m,2) -> symmetric CSR (ptr int64[n+1], idx int32 sorted)."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if len(pairs) == 0:
        return np.zeros(n + 1, np.int64), np.zeros(0, np.int32)
    a = np.concatenate([pairs[:, 0], pairs[:, 1]])
    b = np.concatenate([pairs[:, 1], pairs[:, 0]])
    key = np.unique(a * n + b)
    a, b = key // n, key % n
    ptr = np.zeros(n + 1, np.int64)
    np.add.at(ptr, a + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr, b.astype(np.int32)


@dataclass
class OracleSystem:
    """Plain-array description of a Molly System restricted to the hot path."""

    box: np.ndarray
    mass: np.ndarray
    charge: np.ndarray
    sigma: np.ndarray
    eps: np.ndarray
    inters: list
    excluded_pairs: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    special_pairs: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    dtype: type = np.float64

    def __post_init__(self):
        dt = np.dtype(self.dtype)
        self.n = len(self.mass)
        self.box = np.asarray(self.box, np.float64)
        self._mass = np.ascontiguousarray(self.mass, dt)
        self._charge = np.ascontiguousarray(self.charge, dt)
        self._sigma = np.ascontiguousarray(self.sigma, dt)
        self._eps = np.ascontiguousarray(self.eps, dt)
        self._inters = (InterC * max(1, len(self.inters)))(*[i.to_c() for i in self.inters])
        self._eptr, self._eidx = pairs_to_csr(self.n, self.excluded_pairs)
        self._sptr, self._sidx = pairs_to_csr(self.n, self.special_pairs)
        s = SystemC()
        s.n_atoms = self.n
        s.dtype = 32 if dt == np.float32 else 64
        s.n_inters = len(self.inters)
        s.box[0], s.box[1], s.box[2] = [float(x) for x in self.box]
        s.mass = self._mass.ctypes.data
        s.charge = self._charge.ctypes.data
        s.sigma = self._sigma.ctypes.data
        s.eps = self._eps.ctypes.data
        s.lam = None
        s.inters = self._inters
        s.excl_ptr = self._eptr.ctypes.data if len(self._eidx) else None
        s.excl_idx = self._eidx.ctypes.data if len(self._eidx) else None
        s.spec_ptr = self._sptr.ctypes.data if len(self._sidx) else None
        s.spec_idx = self._sidx.ctypes.data if len(self._sidx) else None
        self._c = s

    def _coords(self, coords):
        return np.ascontiguousarray(coords, self.dtype).reshape(self.n, 3)

    def forces_allpairs(self, coords, n_threads=0, energy=True, virial=False):
        """Brute-force forces (N,3), PE, virial(3,3) over all non-excluded i<j pairs."""
        x = self._coords(coords)
        f = np.zeros_like(x)
        pe = C.c_double(0.0)
        vir = (C.c_double * 9)(*([0.0] * 9))
        nt = n_threads or max_threads()
        rc = lib().orc_forces_allpairs(C.byref(self._c), x.ctypes.data, f.ctypes.data,
                                       C.byref(pe) if energy else None, vir if virial else None, nt)
        assert rc == 0
        return f, pe.value, np.array(list(vir)).reshape(3, 3)

    def neighbor_list(self, coords, r_list):
        x = self._coords(coords)
        out = C.POINTER(NLEntry)()
        n = lib().orc_neighbor_list(C.byref(self._c), x.ctypes.data, float(r_list), C.byref(out))
        arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(n, 3)).copy() if n else \
            np.zeros((0, 3), np.int32)
        lib().orc_free(out)
        return arr

    def forces_nl(self, coords, nl, n_threads=0, energy=True, virial=False):
        x = self._coords(coords)
        nl = np.ascontiguousarray(nl, np.int32)
        f = np.zeros_like(x)
        pe = C.c_double(0.0)
        vir = (C.c_double * 9)(*([0.0] * 9))
        nt = n_threads or max_threads()
        rc = lib().orc_forces_nl(C.byref(self._c), x.ctypes.data, C.cast(nl.ctypes.data, C.POINTER(NLEntry)),
                                 len(nl), f.ctypes.data, C.byref(pe) if energy else None,
                                 vir if virial else None, nt)
        assert rc == 0
        return f, pe.value, np.array(list(vir)).reshape(3, 3)

    def simulate_vv(self, coords, vel, dt, n_steps, remove_cm_every=1, r_list=0.0, nl_every=10, n_threads=0):
        """VelocityVerlet simulate! restatement. Returns (coords, vel, final PE)."""
        x = self._coords(coords).copy()
        v = np.ascontiguousarray(vel, self.dtype).reshape(self.n, 3).copy()
        pe = C.c_double(0.0)
        nt = n_threads or max_threads()
        rc = lib().orc_simulate_vv(C.byref(self._c), x.ctypes.data, v.ctypes.data, float(dt), int(n_steps),
                                   int(remove_cm_every), float(r_list), int(nl_every), nt, C.byref(pe))
        assert rc == 0
        return x, v, pe.value

    def remove_cm(self, vel):
        v = np.ascontiguousarray(vel, self.dtype).reshape(self.n, 3).copy()
        lib().orc_remove_cm(C.byref(self._c), v.ctypes.data)
        return v


# ---------------------------------------------------------------------------
# numpy helpers (scalars/general observables; tiny cases only)
# ---------------------------------------------------------------------------
def vector_1D(c1, c2, side):
    return lib().orc_vector_1D(float(c1), float(c2), float(side))


def wrap_coord_1D(c, side):
    return lib().orc_wrap_coord_1D(float(c), float(side))


def kinetic_energy(mass, vel):
    """src/energy.jl:56-70: K = 1/2 sum m v.v"""
    return 0.5 * float(np.sum(np.asarray(mass, np.float64)[:, None] * np.asarray(vel, np.float64) ** 2))


def temperature(mass, vel, n_constraints_df=0):
    """src/energy.jl:158-175 with df = 3N - 3 (fully periodic 3-D box)."""
    n = len(mass)
    df = 3 * n - 3 - n_constraints_df
    return 2.0 * kinetic_energy(mass, vel) / (df * BOLTZMANN_K)


def lj_dispersion_correction_energy(sigma, eps, box, r_cut):
    """LJDispersionCorrection energy (src/interactions/lennard_jones.jl:192-246).

    E = (factor_6 + factor_12)/V, means over all i<=j pairs (N(N+1)/2 terms), Lorentz sigma,
    geometric epsilon; grouped by distinct (sigma, eps) types for O(T^2).
    """
    sigma = np.asarray(sigma, np.float64)
    eps = np.asarray(eps, np.float64)
    n = len(sigma)
    types, counts = np.unique(np.stack([sigma, eps], 1), axis=0, return_counts=True)
    s6 = 0.0
    s12 = 0.0
    for a in range(len(types)):
        for b in range(a, len(types)):
            npairs = counts[a] * (counts[a] + 1) / 2 if a == b else counts[a] * counts[b]
            sig = (types[a, 0] + types[b, 0]) / 2
            e = np.sqrt(types[a, 1] * types[b, 1])
            s6 += npairs * e * sig ** 6
            s12 += npairs * e * sig ** 12
    n_pairs = n * (n + 1) / 2
    m6, m12 = s6 / n_pairs, s12 / n_pairs
    vol = float(np.prod(box))
    f6 = 8 * np.pi * n * n * (-m6 / (3 * r_cut ** 3))
    f12 = 8 * np.pi * n * n * (m12 / (9 * r_cut ** 9))
    return (f6 + f12) / vol
