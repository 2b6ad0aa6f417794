"""TEST INFRASTRUCTURE (CPU oracle), not product code: TriclinicBoundary geometry restated in numpy.

Reference (Molly.jl v0.23.3): constructor src/spatial.jl:165-215 (basis vectors from lengths + angles :113-143),
minimum image `vector` with approx_images=true :528-534 and the exact 27-image search :536-551, `wrap_coords` :584-600,
LennardJones / Coulomb with DistanceCutoff as in oracle/molly_oracle_impl.h. Pure numpy / Python loops: small cases only.
Parity status: PINNED on the reference's own checks - basis-vector literals (test/basic.jl:133-135), approximate image ==
exact image up to half the smallest height (:221-234), wrap_coords idempotent inside the box (:219) - tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np


def basis_from_lengths_angles(lengths, angles_rad):
    """TriclinicBoundary(bv_lengths, angles): src/spatial.jl:113-143 (alpha = angle(b, c), beta = angle(a, c), gamma = angle(a, b))."""
    a, b, c = [float(v) for v in lengths]
    al, be, ga = [float(v) for v in angles_rad]
    v1 = np.array([a, 0.0, 0.0])
    v2 = np.array([b * np.cos(ga), b * np.sin(ga), 0.0])
    cx = c * np.cos(be)
    cy = c * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
    cz = np.sqrt(c * c - cx * cx - cy * cy)
    return np.array([v1, v2, np.array([cx, cy, cz])])


class Triclinic:
    def __init__(self, basis):
        bv = np.asarray(basis, np.float64).reshape(3, 3)
        if not (bv[0, 0] > 0 and bv[0, 1] == 0 and bv[0, 2] == 0 and bv[1, 1] > 0 and bv[1, 2] == 0 and bv[2, 2] > 0):
            raise ValueError("basis vectors must be lower-triangular with a positive diagonal (src/spatial.jl:173-186)")
        self.bv = bv
        self.rs = np.array([1.0 / bv[0, 0], 1.0 / bv[1, 1], 1.0 / bv[2, 2]])  # reciprocal_size :187
        by, bz, cy, cz = bv[1, 1], bv[1, 2], bv[2, 1], bv[2, 2]
        self.cot_bprojyz_cprojyz = abs((by * cy + bz * cz) / (by * cz - bz * cy))  # :195-199
        self.cprojxy_x_over_z = bv[2, 0] / abs(bv[2, 2])  # :202-204
        self.cprojxy_y_over_z = bv[2, 1] / abs(bv[2, 2])
        self.cot_a_b = bv[1, 0] / bv[1, 1]  # :207-208

    def vector(self, c1, c2):
        """approx_images = true (the default): z, then y, then x (src/spatial.jl:528-534)."""
        dr = np.asarray(c2, np.float64) - np.asarray(c1, np.float64)
        dr = dr - self.bv[2] * np.floor(dr[2] * self.rs[2] + 0.5)
        dr = dr - self.bv[1] * np.floor(dr[1] * self.rs[1] + 0.5)
        dr = dr - self.bv[0] * np.floor(dr[0] * self.rs[0] + 0.5)
        return dr

    def vector_exact(self, c1, c2):
        """approx_images = false: the closest of the 27 images (src/spatial.jl:536-551)."""
        best, best_d = None, np.inf
        for ox in (-1, 0, 1):
            for oy in (-1, 0, 1):
                for oz in (-1, 0, 1):
                    dr = np.asarray(c2, np.float64) + ox * self.bv[0] + oy * self.bv[1] + oz * self.bv[2] - np.asarray(c1, np.float64)
                    d = dr @ dr
                    if d < best_d:
                        best, best_d = dr, d
        return best

    def wrap(self, v):
        """wrap_coords(v, ::TriclinicBoundary): src/spatial.jl:584-600."""
        w = np.asarray(v, np.float64).copy()
        w = w - self.bv[2] * np.floor(w[2] * self.rs[2])
        w = w - self.bv[1] * np.floor((w[1] - w[2] * self.cot_bprojyz_cprojyz) * self.rs[1])
        dx, dy = w[2] * self.cprojxy_x_over_z, w[2] * self.cprojxy_y_over_z
        w = w - self.bv[0] * np.floor((w[0] - dx - (w[1] - dy) * self.cot_a_b) * self.rs[0])
        return w


def forces_energy(tric: Triclinic, coords, sigma, eps, charge=None, r_cut=np.inf, coulomb_const=138.93545764, r_cut_coul=None):
    """All pairs i < j: LennardJones (Lorentz sigma, geometric eps, zero shortcut) [+ plain Coulomb], DistanceCutoff.
    Sign convention src/force.jl:869-874: dr = vector(c_i, c_j), f = (F/r) dr, fs[i] -= f, fs[j] += f; virial += dr (x) f."""
    x = np.asarray(coords, np.float64)
    n = len(x)
    f = np.zeros((n, 3))
    pe = 0.0
    vir = np.zeros((3, 3))
    rcc = r_cut if r_cut_coul is None else r_cut_coul
    for i in range(n):
        for j in range(i + 1, n):
            dr = tric.vector(x[i], x[j])
            r2 = dr @ dr
            r = np.sqrt(r2)
            F = 0.0
            if not (sigma[i] == 0 or sigma[j] == 0 or eps[i] == 0 or eps[j] == 0) and r <= r_cut:
                s = (sigma[i] + sigma[j]) / 2
                e = np.sqrt(eps[i] * eps[j])
                s6 = (s * s / r2) ** 3
                F += (24 * e / r) * (2 * s6 * s6 - s6)
                pe += 4 * e * (s6 * s6 - s6)
            if charge is not None and r <= rcc:
                kqq = coulomb_const * charge[i] * charge[j]
                F += kqq / r2
                pe += kqq / r
            fv = (F / r) * dr
            f[i] -= fv
            f[j] += fv
            vir += np.outer(dr, fv)
    return f, pe, vir
