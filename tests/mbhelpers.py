"""Deterministic synthetic systems shared by the tests, bench.py and smoke().

Generators follow the reference's benchmark scripts:
  * argon LJ fluid, rho = 1400 kg/m^3, sigma 0.34 nm, eps 0.997 kJ/mol, m 39.948
    (benchmark/benchmark_gpu_tiles.jl:13-56), FCC lattice + N(0, 0.01 nm) jitter for dynamics;
  * README example: 100 atoms, box 2.0 nm, sigma 0.3, eps 0.2, m 10 (README.md:72-95).
"""
from __future__ import annotations

import numpy as np

ARGON = dict(mass=39.948, sigma=0.34, eps=0.997)
ARGON_DENSITY = 1400.0 * 6.02214076e23 / (39.948e-3) / 1e27  # atoms / nm^3 = 21.105
K_B = 8.31446261815324e-3


def fcc_lattice(cells: int, a: float):
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    g = np.stack(np.meshgrid(np.arange(cells), np.arange(cells), np.arange(cells), indexing="ij"), -1).reshape(-1, 3)
    x = (g[:, None, :] + base[None, :, :]).reshape(-1, 3) * a
    return x, cells * a


def lj_fluid(cells: int, seed: int = 42, jitter: float = 0.01, temp: float = 90.0, dtype=np.float32):
    """4*cells^3 argon atoms on an FCC lattice at the reference density, jittered; MB velocities, CM removed."""
    a = (4.0 / ARGON_DENSITY) ** (1.0 / 3.0)
    x, L = fcc_lattice(cells, a)
    rng = np.random.default_rng(seed)
    x = x + rng.normal(0.0, jitter, x.shape) + 0.25 * a
    x = x - np.floor(x / L) * L
    n = len(x)
    v = rng.normal(0.0, np.sqrt(K_B * temp / ARGON["mass"]), (n, 3))
    v -= v.mean(0)
    return dict(n=n, box=np.array([L, L, L]), coords=x.astype(dtype), velocities=v.astype(dtype),
                mass=np.full(n, ARGON["mass"]), charge=np.zeros(n), sigma=np.full(n, ARGON["sigma"]),
                eps=np.full(n, ARGON["eps"]))


def readme_system(n: int = 100, box: float = 2.0, seed: int = 1, min_dist: float = 0.3, dtype=np.float64):
    """README.md:72-95: place_atoms-style rejection sampling (setup.jl:23-60), T = 298 K."""
    rng = np.random.default_rng(seed)
    pts = []
    while len(pts) < n:
        c = rng.random(3) * box
        ok = True
        for p in pts:
            d = c - p
            d -= box * np.round(d / box)
            if d @ d < min_dist * min_dist:
                ok = False
                break
        if ok:
            pts.append(c)
    x = np.array(pts)
    v = rng.normal(0.0, np.sqrt(K_B * 298.0 / 10.0), (n, 3))
    return dict(n=n, box=np.array([box] * 3), coords=x.astype(dtype), velocities=v.astype(dtype),
                mass=np.full(n, 10.0), charge=np.zeros(n), sigma=np.full(n, 0.3), eps=np.full(n, 0.2))


def molecular_system(n_mol: int, box, seed: int = 7, dtype=np.float64, stable: bool = False):
    """Small charged 4-site chain molecules A-B-C-D on a jittered grid: 1-2 and 1-3 pairs excluded,
    1-4 pairs special; three LJ types including a zero-epsilon one (TIP3P-hydrogen-like)."""
    rng = np.random.default_rng(seed)
    box = np.asarray(box, float)
    per_dim = int(np.ceil(n_mol ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(per_dim)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n_mol]
    centers = (grid + 0.5) / per_dim * box
    coords, q, sig, eps, mass = [], [], [], [], []
    excl, spec = [], []
    tq = [0.4, -0.4, 0.3, -0.3]
    ts = [0.32, 0.30, 0.25, 0.10]
    te = [0.6, 0.4, 0.2, 0.0]
    tm = [12.0, 14.0, 16.0, 1.008]
    if stable:  # for dynamics without bonded terms: every site keeps a repulsive core, charges are mild
        te = [0.6, 0.4, 0.2, 0.15]
        ts = [0.32, 0.30, 0.25, 0.22]
        tq = [0.2, -0.2, 0.15, -0.15]
        tm = [12.0, 14.0, 16.0, 4.0]
    for m, c in enumerate(centers):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        e = np.cross(d, rng.normal(size=3))
        e /= np.linalg.norm(e)
        pos = [c, c + 0.11 * d, c + 0.11 * d + 0.11 * e, c + 0.11 * e + 0.18 * d]
        base = 4 * m
        for k in range(4):
            coords.append(pos[k] + rng.normal(0, 0.005, 3))
            q.append(tq[k]); sig.append(ts[k]); eps.append(te[k]); mass.append(tm[k])
        excl += [(base, base + 1), (base + 1, base + 2), (base + 2, base + 3), (base, base + 2), (base + 1, base + 3)]
        spec += [(base, base + 3)]
    x = np.array(coords)
    x = x - np.floor(x / box) * box
    n = len(x)
    v = rng.normal(0.0, 0.3, (n, 3))
    return dict(n=n, box=box, coords=x.astype(dtype), velocities=v.astype(dtype), mass=np.array(mass),
                charge=np.array(q), sigma=np.array(sig), eps=np.array(eps),
                excluded=np.array(excl, np.int32), special=np.array(spec, np.int32))


def make_oracle(sysd, inters, dtype=np.float64):
    from oracle import oracle as o
    return o.OracleSystem(box=sysd["box"], mass=sysd["mass"], charge=sysd["charge"], sigma=sysd["sigma"],
                          eps=sysd["eps"], inters=inters, excluded_pairs=sysd.get("excluded", np.zeros((0, 2), np.int32)),
                          special_pairs=sysd.get("special", np.zeros((0, 2), np.int32)), dtype=dtype)


def make_system(sysd, inters, dtype, r_list=0.0, n_steps=0):
    """mollyb200.System for the same description (exception pairs become 1-based)."""
    import mollyb200 as mb
    atoms = mb.atoms_from_arrays(sysd["mass"], sysd["charge"], sysd["sigma"], sysd["eps"], dtype)
    nf = None
    if r_list > 0 or "excluded" in sysd:
        nf = mb.GPUNeighborFinder(dist_cutoff=r_list,
                                  excluded_pairs=sysd.get("excluded", np.zeros((0, 2), np.int32)) + 1,
                                  special_pairs=sysd.get("special", np.zeros((0, 2), np.int32)) + 1, n_steps=n_steps)
    return mb.System(atoms=atoms, coords=sysd["coords"].astype(dtype), boundary=mb.CubicBoundary(*sysd["box"]),
                     velocities=sysd["velocities"].astype(dtype), pairwise_inters=inters, neighbor_finder=nf, dtype=dtype)
