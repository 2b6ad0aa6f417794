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


# ---------------------------------------------------------------------------------------------------
# 6mrr (BASELINE config 3): full system from the golden fixture
# ---------------------------------------------------------------------------------------------------
def sixmrr_description(g):
    box = g["box"]
    x = g["coords"] - np.floor(g["coords"] / box) * box
    return dict(n=len(x), box=box, coords=x, velocities=g["velocities_300K"], mass=g["mass"], charge=g["charge"],
                sigma=g["sigma"], eps=g["eps"], excluded=g["excluded"], special=g["special"])


def sixmrr_specific_lists(g):
    import mollyb200 as mb
    b, a = g["bond_idx"] + 1, g["angle_idx"] + 1
    t = np.concatenate([g["proper_idx"], g["improper_idx"]]) + 1
    tp = np.concatenate([g["proper_par"], g["improper_par"]])
    return (mb.InteractionList2Atoms(b[:, 0], b[:, 1], g["bond_par"][:, 0], g["bond_par"][:, 1]),
            mb.InteractionList3Atoms(a[:, 0], a[:, 1], a[:, 2], g["angle_par"][:, 0], g["angle_par"][:, 1]),
            mb.InteractionList4Atoms(t[:, 0], t[:, 1], t[:, 2], t[:, 3], tp[:, 0], tp[:, 1], tp[:, 2]))


def sixmrr_system(g, dtype, r_list=1.2, n_steps=0, bonded=True, coords=None, velocities=None, device=0, dispersion=False):
    """System(6mrr_equil.pdb, ff99SBildn + tip3p; nonbonded_method=:cutoff) as benchmark/protein.jl:24-37 builds it:
    LJ(rc 1.0, w14 0.5) + CoulombReactionField(rc 1.0, eps 78.3, w14 0.8333) + bonds/angles/torsions."""
    import mollyb200 as mb
    sd = sixmrr_description(g)
    atoms = mb.atoms_from_arrays(sd["mass"], sd["charge"], sd["sigma"], sd["eps"], dtype)
    inters = (mb.LennardJones(cutoff=mb.DistanceCutoff(1.0), use_neighbors=True, weight_special=float(g["lj14scale"])),
              mb.CoulombReactionField(dist_cutoff=1.0, use_neighbors=True, weight_special=float(g["coulomb14scale"])))
    nf = mb.GPUNeighborFinder(dist_cutoff=r_list, excluded_pairs=g["excluded"] + 1, special_pairs=g["special"] + 1, n_steps=n_steps)
    x = sd["coords"] if coords is None else coords
    v = sd["velocities"] if velocities is None else velocities
    return mb.System(atoms=atoms, coords=np.asarray(x).astype(dtype), boundary=mb.CubicBoundary(*sd["box"]),
                     velocities=np.asarray(v).astype(dtype), pairwise_inters=inters, neighbor_finder=nf, dtype=dtype,
                     specific_inter_lists=sixmrr_specific_lists(g) if bonded else (), device=device,
                     general_inters=(mb.LJDispersionCorrection(1.0),) if dispersion else ())  # setup.jl:2000-2004


def sixmrr_oracle(g, dtype=np.float64):
    from oracle import oracle as o
    sd = sixmrr_description(g)
    inters = [o.Inter(o.LJ, o.CUT_DISTANCE, 1.0, weight_special=float(g["lj14scale"]), use_neighbors=True),
              o.Inter(o.CRF, o.CUT_DISTANCE, 1.0, weight_special=float(g["coulomb14scale"]), use_neighbors=True)]
    return make_oracle(sd, inters, dtype=dtype), sd


def bonded_forces_oracle(g, x):
    from oracle import bonded as bd
    box = g["box"]
    f = np.zeros_like(x, dtype=np.float64)
    e = 0.0
    for fn, idx, par in ((bd.bond_forces, "bond_idx", "bond_par"), (bd.angle_forces, "angle_idx", "angle_par"),
                         (bd.torsion_forces, "proper_idx", "proper_par"), (bd.torsion_forces, "improper_idx", "improper_par")):
        ff, ee = fn(np.asarray(x, np.float64), box, g[idx], g[par])
        f += ff
        e += ee
    return f, e


def oracle_vv_with_bonded(g, x, v, dt, n_steps, r_list=1.2, nl_every=10):
    """VelocityVerlet simulate! (src/simulators.jl:547-668) with pairwise (C oracle, neighbour list) + bonded (numpy)
    forces, f64, remove_CM_motion = 1."""
    orc, sd = sixmrr_oracle(g)
    box, m = sd["box"], sd["mass"]
    x = x - np.floor(x / box) * box
    v = orc.remove_cm(v)

    def forces(xx, nl):
        f, _, _ = orc.forces_nl(xx, nl, energy=False)
        return f + bonded_forces_oracle(g, xx)[0]
    nl = orc.neighbor_list(x, r_list)
    f = forces(x, nl)
    for step in range(1, n_steps + 1):
        v = v + f / m[:, None] * (dt / 2)
        x = x + v * dt
        x = x - np.floor(x / box) * box
        f = forces(x, nl)
        v = v + f / m[:, None] * (dt / 2)
        v = orc.remove_cm(v)
        if step % nl_every == 0:
            nl = orc.neighbor_list(x, r_list)
    return x, v


def oracle_vv_pme(g, x, v, dt, n_steps, r_list=1.2, nl_every=10):
    """simulate!(sys_pme_exact, VelocityVerlet(dt), n) of test/protein.jl:277-299 restated with the oracle's pieces (f64):
    LJ + CoulombEwald real space (C oracle over the neighbour list) + bonded + EwaldExclusion + PME reciprocal space
    (oracle/pme.py, numpy), remove_CM_motion = 1."""
    from oracle import oracle as o, pme
    sd = sixmrr_description(g)
    box, m = sd["box"], sd["mass"]
    alpha = pme.pme_alpha(1.0)
    inters = [o.Inter(o.LJ, o.CUT_DISTANCE, 1.0, weight_special=float(g["lj14scale"]), use_neighbors=True),
              o.Inter(o.EWALD_REAL, o.CUT_DISTANCE, 1.0, weight_special=float(g["coulomb14scale"]), ewald_alpha=alpha,
                      use_neighbors=True)]
    orc = make_oracle(sd, inters)
    excl = np.concatenate([g["excluded"], g["special"]])
    x = x - np.floor(x / box) * box
    v = orc.remove_cm(v)

    def forces(xx, nl):
        f, _, _ = orc.forces_nl(xx, nl, energy=False)
        fr, _, _ = pme.pme_reciprocal(xx, g["charge"], box, r_cut=1.0, error_tol=0.0005, order=5)
        fx, _ = pme.ewald_exclusion(xx, g["charge"], box, excl)
        return f + fr + fx + bonded_forces_oracle(g, xx)[0]
    nl = orc.neighbor_list(x, r_list)
    f = forces(x, nl)
    for step in range(1, n_steps + 1):
        v = v + f / m[:, None] * (dt / 2)
        x = x + v * dt
        x = x - np.floor(x / box) * box
        f = forces(x, nl)
        v = v + f / m[:, None] * (dt / 2)
        v = orc.remove_cm(v)
        if step % nl_every == 0:
            nl = orc.neighbor_list(x, r_list)
    return x, v


def sixmrr_pme_system(g, dtype, r_list=1.2, exact=True, velocities=None):
    """System(6mrr; nonbonded_method=:pme, approximate_pme=!exact) of test/protein.jl:76-85 / setup.jl:1894-1927: LJ +
    CoulombEwald + bonded lists + PME + EwaldExclusion(excluded or special) + LJDispersionCorrection."""
    import mollyb200 as mb
    sd = sixmrr_description(g)
    atoms = mb.atoms_from_arrays(sd["mass"], sd["charge"], sd["sigma"], sd["eps"], dtype)
    inters = (mb.LennardJones(cutoff=mb.DistanceCutoff(1.0), use_neighbors=True, weight_special=float(g["lj14scale"])),
              mb.CoulombEwald(dist_cutoff=1.0, error_tol=0.0005, use_neighbors=True, weight_special=float(g["coulomb14scale"]),
                              approximate_erfc=not exact))
    nf = mb.GPUNeighborFinder(dist_cutoff=r_list, excluded_pairs=g["excluded"] + 1, special_pairs=g["special"] + 1)
    pme = mb.PME(dist_cutoff=1.0, error_tol=0.0005, excluded_pairs=np.concatenate([g["excluded"], g["special"]]) + 1)
    v = sd["velocities"] if velocities is None else velocities
    return mb.System(atoms=atoms, coords=sd["coords"].astype(dtype), boundary=mb.CubicBoundary(*sd["box"]),
                     velocities=np.asarray(v).astype(dtype), pairwise_inters=inters, neighbor_finder=nf, dtype=dtype,
                     specific_inter_lists=sixmrr_specific_lists(g), general_inters=(pme, mb.LJDispersionCorrection(1.0)))
