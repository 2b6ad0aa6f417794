"""Particle-mesh Ewald reciprocal space + Ewald exclusion terms, numpy restatement (TEST INFRASTRUCTURE ONLY).

SURVEY.md §8(f)-3: the next row after the pairwise path. Restates, in float64, what the reference's `PME` general
interaction and `EwaldExclusion` specific interaction compute (src/interactions/ewald.jl, itself modelled on OpenMM's
Reference PME), so that the GPU implementation of the next round has a pinned checker:

  pme_alpha, pme_mesh_dims          ewald.jl:373 (alpha = sqrt(-ln 2 tol) / rc), :484-487 (ceil(2 alpha L / (3 tol^0.2)), >= 6)
  bspline_moduli                    ewald.jl:311-361
  grid_placement, bsplines          ewald.jl:489-498, :518-556 (order-5 cardinal B-splines and derivatives)
  spread -> FFT -> convolution -> backward FFT -> interpolate   ewald.jl:598-617, :676-732, :838-873, :904-958
  self / neutralising-background energy                          ewald.jl:947-956
  ewald_exclusion                   ewald.jl:1016-1055; pair set = excluded OR special pairs (find_excluded_pairs :960-977)

The real-space term (`CoulombEwald`, coulomb.jl:1395-1441) is in the C oracle (kind EWALD_REAL). Pinned by
tests/test_oracle.py against OpenMM's `forces_all_pme_exact` / `energy_all_pme_exact` for 6mrr (test/protein.jl:206-276).
Only tests/ may import this module.
"""
from math import erf as _erf

import numpy as np

COULOMB_CONST = 138.93545764  # kJ mol^-1 nm e^-2 (src/interactions/coulomb.jl:16)


def pme_alpha(r_cut: float, error_tol: float = 0.0005) -> float:
    return float(np.sqrt(-np.log(2.0 * error_tol)) / r_cut)


def pme_mesh_dims(box, alpha: float, error_tol: float = 0.0005):
    return tuple(max(int(np.ceil(2.0 * alpha * L / (3.0 * error_tol ** 0.2))), 6) for L in box)


def bspline_moduli(order: int, mesh_dims):
    """|DFT of the B-spline coefficients|^2 per dimension (ewald.jl:311-361)."""
    data = np.zeros(order)
    data[0] = 1.0
    for k in range(3, order):  # k = 3 .. order-1 (1-based recursion of the reference)
        d = 1.0 / (k - 1.0)
        data[k - 1] = 0.0
        for l in range(1, k - 1):
            data[k - l - 1] = d * (l * data[k - l - 2] + (k - l) * data[k - l - 1])
        data[0] *= d
    d = 1.0 / (order - 1.0)
    data[order - 1] = 0.0
    for l in range(1, order - 1):
        data[order - l - 1] = d * (l * data[order - l - 2] + (order - l) * data[order - l - 1])
    data[0] *= d
    out = []
    for n in mesh_dims:
        bs = np.zeros(max(mesh_dims))
        bs[1:order + 1] = data  # bsplines_data[i+1] = data[i]
        j = np.arange(n)
        i = np.arange(n)[:, None]
        arg = 2.0 * np.pi * i * j[None, :] / n
        sc = (bs[:n][None, :] * np.cos(arg)).sum(1)
        ss = (bs[:n][None, :] * np.sin(arg)).sum(1)
        m = sc ** 2 + ss ** 2
        fixed = m.copy()
        for q in range(n):  # sequential, like the reference (a repaired entry can feed its successor)
            if fixed[q] < 1e-7:
                fixed[q] = 0.5 * (fixed[(q - 1) % n] + fixed[(q + 1) % n])
        out.append(fixed)
    return out


def bsplines(frac: np.ndarray, order: int):
    """theta, dtheta (n, order) for grid fractions frac (n,) — ewald.jl:518-556."""
    n = len(frac)
    th = np.zeros((n, order))
    dth = np.zeros((n, order))
    dr = frac
    th[:, order - 1] = 0.0
    th[:, 1] = dr
    th[:, 0] = 1.0 - dr
    for k in range(3, order):
        d = 1.0 / (k - 1.0)
        th[:, k - 1] = d * dr * th[:, k - 2]
        for l in range(1, k - 1):
            th[:, k - l - 1] = d * ((dr + l) * th[:, k - l - 2] + (k - l - dr) * th[:, k - l - 1])
        th[:, 0] *= d * (1.0 - dr)
    dth[:, 0] = -th[:, 0]
    for k in range(1, order):
        dth[:, k] = th[:, k - 1] - th[:, k]
    d = 1.0 / (order - 1.0)
    th[:, order - 1] = d * dr * th[:, order - 2]
    for l in range(1, order - 1):
        th[:, order - l - 1] = d * ((dr + l) * th[:, order - l - 2] + (order - l - dr) * th[:, order - l - 1])
    th[:, 0] *= d * (1.0 - dr)
    return th, dth


def pme_reciprocal(coords, charges, box, r_cut=1.0, error_tol=0.0005, order=5, eps_r=1.0, ke=COULOMB_CONST):
    """Reciprocal-space forces (n,3) and energy incl. self and neutralising-background terms (ewald.jl:904-958).
    Cubic (orthorhombic) box: recip_box = diag(1/L)."""
    x = np.asarray(coords, np.float64)
    q = np.asarray(charges, np.float64)
    box = np.asarray(box, np.float64)
    n = len(x)
    alpha = pme_alpha(r_cut, error_tol)
    K = pme_mesh_dims(box, alpha, error_tol)
    bsm = bspline_moduli(order, K)
    V = float(np.prod(box))
    f_div = ke / eps_r
    # grid placement
    t = x / box
    t = (t - np.floor(t)) * np.array(K)
    ti = np.floor(t).astype(np.int64)
    frac = t - ti
    idx0 = ti % np.array(K)
    th = [None] * 3
    dth = [None] * 3
    for d in range(3):
        th[d], dth[d] = bsplines(frac[:, d], order)
    # spread
    grid = np.zeros(K, np.float64)  # [x, y, z]
    o = np.arange(order)
    ix = (idx0[:, 0, None] + o) % K[0]
    iy = (idx0[:, 1, None] + o) % K[1]
    iz = (idx0[:, 2, None] + o) % K[2]
    w = q[:, None, None, None] * th[0][:, :, None, None] * th[1][:, None, :, None] * th[2][:, None, None, :]
    IX = np.broadcast_to(ix[:, :, None, None], w.shape)
    IY = np.broadcast_to(iy[:, None, :, None], w.shape)
    IZ = np.broadcast_to(iz[:, None, None, :], w.shape)
    np.add.at(grid, (IX.ravel(), IY.ravel(), IZ.ravel()), w.ravel())
    # forward FFT, convolution
    S = np.fft.fftn(grid)
    kx = np.arange(K[0])
    ky = np.arange(K[1])
    kz = np.arange(K[2])
    mx = np.where(kx < 0.5 * (K[0] + 1), kx, kx - K[0]) / box[0]
    my = np.where(ky < 0.5 * (K[1] + 1), ky, ky - K[1]) / box[1]
    mz = np.where(kz < 0.5 * (K[2] + 1), kz, kz - K[2]) / box[2]
    m2 = mx[:, None, None] ** 2 + my[None, :, None] ** 2 + mz[None, None, :] ** 2
    factor = np.pi ** 2 / alpha ** 2
    denom = m2 * (np.pi * V) * bsm[0][:K[0], None, None] * bsm[1][None, :K[1], None] * bsm[2][None, None, :K[2]]
    with np.errstate(divide="ignore", invalid="ignore"):
        eterm = f_div * np.exp(-factor * m2) / denom
    eterm[0, 0, 0] = 0.0
    e_recip = 0.5 * float((eterm * (S.real ** 2 + S.imag ** 2)).sum())
    conv = S * eterm
    conv[0, 0, 0] = S[0, 0, 0]  # the reference leaves the k = 0 element untouched (no force: sum of dtheta is 0)
    pot = np.fft.ifftn(conv).real * np.prod(K)  # bfft = unnormalised backward transform
    # interpolate
    g = pot[IX, IY, IZ]  # (n, order, order, order)
    fx = (dth[0][:, :, None, None] * th[1][:, None, :, None] * th[2][:, None, None, :] * g).sum((1, 2, 3))
    fy = (th[0][:, :, None, None] * dth[1][:, None, :, None] * th[2][:, None, None, :] * g).sum((1, 2, 3))
    fz = (th[0][:, :, None, None] * th[1][:, None, :, None] * dth[2][:, None, None, :] * g).sum((1, 2, 3))
    F = -q[:, None] * np.stack([fx * K[0] / box[0], fy * K[1] / box[1], fz * K[2] / box[2]], 1)
    charge_e = -f_div * np.pi * q.sum() ** 2 / (2.0 * V * alpha ** 2)
    self_e = -f_div * (q ** 2).sum() * alpha / np.sqrt(np.pi) + charge_e
    return F, e_recip + self_e, dict(alpha=alpha, mesh_dims=K, e_recip=e_recip, e_self=self_e)


def ewald_exclusion(coords, charges, box, pairs, r_cut=1.0, error_tol=0.0005, eps_r=1.0, ke=COULOMB_CONST):
    """EwaldExclusion over `pairs` (m,2, 0-based; excluded OR special pairs): forces (n,3), energy (ewald.jl:1016-1055)."""
    x = np.asarray(coords, np.float64)
    q = np.asarray(charges, np.float64)
    box = np.asarray(box, np.float64)
    alpha = pme_alpha(r_cut, error_tol)
    f_div = ke / eps_r
    i, j = np.asarray(pairs)[:, 0], np.asarray(pairs)[:, 1]
    d = x[j] - x[i]
    d -= box * np.round(d / box)  # vector(c_i, c_j) minimum image
    r = np.sqrt((d * d).sum(1))
    ar = alpha * r
    erf_ar = np.array([_erf(v) for v in ar])
    qq = q[i] * q[j]
    big = erf_ar > 1e-6
    de_dr = np.where(big, f_div * qq / r ** 3 * (erf_ar - 2.0 * ar * np.exp(-ar * ar) / np.sqrt(np.pi)), 0.0)
    fvec = de_dr[:, None] * d  # force on i is +F, on j is -F (SpecificForce2Atoms(F, -F))
    F = np.zeros_like(x)
    np.add.at(F, i, fvec)
    np.add.at(F, j, -fvec)
    e = np.where(big, -f_div * qq / r * erf_ar, -alpha * 2.0 * f_div * qq / np.sqrt(np.pi))
    return F, float(e.sum())
