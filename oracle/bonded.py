"""numpy restatement of Molly's bonded terms (TEST INFRASTRUCTURE): HarmonicBond, HarmonicAngle, PeriodicTorsion.

Formulas: SURVEY.md Appendix B.1 — src/interactions/harmonic_bond.jl:13-54, harmonic_angle.jl:45-67,
periodic_torsion.jl:17-142 (dihedral by atan2, spatial.jl:882-894). All displacements are minimum-image
(`vector`, src/spatial.jl:491-519). Returns forces (n,3) and the energy."""
import numpy as np


def _mic(d, box):
    return d - box * np.round(d / box)


def bond_forces(x, box, idx, par):
    f = np.zeros_like(x)
    ab = _mic(x[idx[:, 1]] - x[idx[:, 0]], box)
    r = np.linalg.norm(ab, axis=1)
    k, r0 = par[:, 0], par[:, 1]
    c = k * (r - r0)
    fi = (c / r)[:, None] * ab
    np.add.at(f, idx[:, 0], fi)
    np.add.at(f, idx[:, 1], -fi)
    return f, float(np.sum(0.5 * k * (r - r0) ** 2))


def angle_forces(x, box, idx, par):
    f = np.zeros_like(x)
    ba = _mic(x[idx[:, 0]] - x[idx[:, 1]], box)
    bc = _mic(x[idx[:, 2]] - x[idx[:, 1]], box)
    nba, nbc = np.linalg.norm(ba, axis=1), np.linalg.norm(bc, axis=1)
    cos = np.clip(np.sum(ba * bc, 1) / (nba * nbc), -1.0, 1.0)
    th = np.arccos(cos)
    k, th0 = par[:, 0], par[:, 1]
    n = np.cross(ba, bc)
    pa = np.cross(ba, n)
    pc = np.cross(-bc, n)
    pa /= np.linalg.norm(pa, axis=1)[:, None]
    pc /= np.linalg.norm(pc, axis=1)[:, None]
    t = -k * (th - th0)
    fa = (t / nba)[:, None] * pa
    fc = (t / nbc)[:, None] * pc
    np.add.at(f, idx[:, 0], fa)
    np.add.at(f, idx[:, 2], fc)
    np.add.at(f, idx[:, 1], -fa - fc)
    return f, float(np.sum(0.5 * k * (th - th0) ** 2))


def torsion_forces(x, box, idx, par):
    f = np.zeros_like(x)
    if len(idx) == 0:
        return f, 0.0
    ab = _mic(x[idx[:, 1]] - x[idx[:, 0]], box)
    bc = _mic(x[idx[:, 2]] - x[idx[:, 1]], box)
    cd = _mic(x[idx[:, 3]] - x[idx[:, 2]], box)
    m = np.cross(ab, bc)
    n = np.cross(bc, cd)
    nbc = np.linalg.norm(bc, axis=1)
    th = np.arctan2(np.sum(np.cross(m, n) * bc, 1) / nbc, np.sum(m * n, 1))
    per, phase, k = par[:, 0], par[:, 1], par[:, 2]
    dedth = -k * per * np.sin(per * th - phase)
    m2, n2 = np.sum(m * m, 1), np.sum(n * n, 1)
    fi = (dedth * nbc / m2)[:, None] * m
    fl = (-dedth * nbc / n2)[:, None] * n
    v = ((-np.sum(ab * bc, 1)) / nbc ** 2)[:, None] * fi - ((-np.sum(cd * bc, 1)) / nbc ** 2)[:, None] * fl
    fj = v - fi
    fk = -v - fl
    for col, ff in zip(range(4), (fi, fj, fk, fl)):
        np.add.at(f, idx[:, col], ff)
    return f, float(np.sum(k * (1.0 + np.cos(per * th - phase))))
