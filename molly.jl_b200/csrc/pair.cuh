// pair.cuh — per-pair physics of the non-bonded path (device side).
//
// Reference formulas (Molly.jl v0.23.3):
//   LennardJones force / energy      src/interactions/lennard_jones.jl:79-140
//   mixing + zero shortcut           src/mixing.jl:5-34
//   Coulomb                          src/interactions/coulomb.jl:71-120
//   CoulombReactionField             src/interactions/coulomb.jl:748-814
//   CoulombEwald real space          src/interactions/coulomb.jl:1395-1441
//   cutoffs                          src/cutoffs.jl:15-45, :99-190
// Sign convention (src/force.jl:869-874): dr = c_j - c_i, f = (F/r) dr, fs[i] -= f. The kernels work
// with d = c_i - c_j, so the force on i is +fr * d with fr = F/r.
#pragma once
#include "common.cuh"
#include "cutoffs2.cuh"

namespace mb {

enum { COUL_NONE = 0, COUL_PLAIN = 1, COUL_CRF = 2, COUL_EWALD = 3 };
enum { CUT_NONE = 0, CUT_DISTANCE = 1, CUT_SHIFTED_POTENTIAL = 2, CUT_SHIFTED_FORCE = 3 };  // CUT_CUBIC_SPLINE = 4, CUT_POLYNOMIAL = 5: cutoffs2.cuh
// Kernel variants by cutoff family (template parameter CUTM of the pair functions and kernels): the plain truncation is
// the fast path, the shifted and the two-point (dist_activation, dist_cutoff) families are separate instantiations so
// that their extra constants and the sqrt never touch the registers of the fast path.
enum { CUTM_PLAIN = 0, CUTM_SHIFTED = 1, CUTM_TWO_POINT = 2 };

// Interaction tuple digested on the host into kernel constants.
template <typename T>
struct PairParams {
    // Lennard-Jones
    int has_lj;
    int lj_cut_kind;
    int geo_sigma;  // 0: Lorentz sigma (stored half-sigma), 1: geometric sigma (stored sqrt(sigma))
    int uniform_lj; // every atom has the same non-zero (sigma, eps)
    T lj_rc2, lj_rc, lj_inv_rc, lj_inv_rc2;
    T lj_ra, lj_inv_ra2;  // two-point cutoffs: dist_activation
    T lj_w14;
    T uni_sig2, uni_eps;
    T uni_A, uni_B;  // 48 eps sigma^12, 24 eps sigma^6 (uniform LJ, plain cutoff fast path)
    // Coulomb family
    int coul_kind;
    int coul_cut_kind;
    T c_rc2, c_rc, c_inv_rc, c_inv_rc2;
    T c_ra;
    T ke, krf, crf, c_w14, alpha;
    int approx_erfc;  // CoulombEwald(approximate_erfc=true) is the reference's default (coulomb.jl:1331)
    // whether the interaction goes through the neighbour list: only then do exclusions / special flags apply
    // (the reference's use_neighbors=false loop visits every pair with special=false, src/force.jl:828-855)
    int lj_nl, c_nl;
};

// LJ term: returns F/r and energy for sigma^2, eps at squared distance r2 (inv_r2 = 1/r2).
template <typename T, int CUTM, bool ENERGY>
__device__ __forceinline__ void lj_term(const PairParams<T>& P, T sig2, T eps, T r2, T inv_r2, T& fr, T& e) {
    T s2 = sig2 * inv_r2;
    T s6 = s2 * s2 * s2;
    T eps24 = (T)24 * eps;
    fr = eps24 * ((T)2 * s6 * s6 - s6) * inv_r2;
    if (ENERGY || CUTM == CUTM_TWO_POINT) e = (T)4 * eps * (s6 * s6 - s6);
    if (CUTM == CUTM_TWO_POINT) {
        // CubicSplineCutoff / PolynomialCutoff (src/cutoffs.jl:174-253): unchanged up to dist_activation, switched beyond
        if (P.lj_cut_kind >= CUT_CUBIC_SPLINE && r2 > P.lj_ra * P.lj_ra) {
            const T a2 = sig2 * P.lj_inv_ra2, a6 = a2 * a2 * a2;
            const T v_act = (T)4 * eps * (a6 * a6 - a6);
            const T f_act = eps24 * ((T)2 * a6 * a6 - a6) / P.lj_ra;
            cut2_apply<T>(P.lj_cut_kind, P.lj_ra, P.lj_rc, fsqrt(r2), v_act, f_act, fr, e);
        }
    }
    if (CUTM != CUTM_PLAIN) {  // (the two-point instantiation also serves a shifted partner interaction)
        if (P.lj_cut_kind == CUT_SHIFTED_POTENTIAL || P.lj_cut_kind == CUT_SHIFTED_FORCE) {
            T c2 = sig2 * P.lj_inv_rc2;
            T c6 = c2 * c2 * c2;
            T ec = (T)4 * eps * (c6 * c6 - c6);
            if (P.lj_cut_kind == CUT_SHIFTED_FORCE) {
                T fc = eps24 * ((T)2 * c6 * c6 - c6) * P.lj_inv_rc;  // F(rc)
                T inv_r = frsqrt(r2);
                fr -= fc * inv_r;
                if (ENERGY) e += (r2 * inv_r - P.lj_rc) * fc - ec;
            } else {
                if (ENERGY) e -= ec;
            }
        }
    }
}

// Coulomb-family term. kqq = ke*qi*qj. inv_r = 1/r.
template <typename T, int COUL, int CUTM, bool ENERGY, bool SPECIAL>
__device__ __forceinline__ void coul_term(const PairParams<T>& P, T kqq, T r2, T inv_r, T inv_r2, T& fr, T& e) {
    if (COUL == COUL_PLAIN) {
        fr = kqq * inv_r * inv_r2;
        if (ENERGY || CUTM == CUTM_TWO_POINT) e = kqq * inv_r;
        if (CUTM == CUTM_TWO_POINT) {
            if (P.coul_cut_kind >= CUT_CUBIC_SPLINE && r2 > P.c_ra * P.c_ra)
                cut2_apply<T>(P.coul_cut_kind, P.c_ra, P.c_rc, fsqrt(r2), kqq / P.c_ra, kqq / (P.c_ra * P.c_ra), fr, e);
        }
        if (CUTM != CUTM_PLAIN) {
            if (P.coul_cut_kind == CUT_SHIFTED_FORCE) {
                fr = kqq * (inv_r2 - P.c_inv_rc2) * inv_r;
                if (ENERGY) e = kqq * (inv_r + (r2 * inv_r - P.c_rc) * P.c_inv_rc2 - P.c_inv_rc);
            } else if (P.coul_cut_kind == CUT_SHIFTED_POTENTIAL) {
                if (ENERGY) e -= kqq * P.c_inv_rc;
            }
        }
    } else if (COUL == COUL_CRF) {
        if (SPECIAL) {  // 1-4 pairs do not use the reaction field (coulomb.jl:760-763)
            fr = kqq * inv_r * inv_r2;
            if (ENERGY) e = kqq * inv_r;
        } else {
            fr = kqq * (inv_r * inv_r2 - (T)2 * P.krf);
            if (ENERGY) e = kqq * (inv_r + P.krf * r2 - P.crf);
        }
    } else if (COUL == COUL_EWALD) {
        if (SPECIAL) {  // plain weighted Coulomb for special pairs (coulomb.jl:1419-1423)
            fr = kqq * inv_r * inv_r2;
            if (ENERGY) e = kqq * inv_r;
        } else {
            T r = r2 * inv_r;
            T ar = P.alpha * r;
            T ex = exp(-ar * ar);
            T erfc_ar;
            if (P.approx_erfc) {  // calc_erfc, coulomb.jl:1384-1393 (Abramowitz & Stegun 7.1.26), evaluated in T like the reference
                const T t = (T)1 / ((T)1 + (T)0.3275911 * ar);
                erfc_ar = ((T)0.254829592 + ((T)-0.284496736 + ((T)1.421413741 + ((T)-1.453152027 + (T)1.061405429 * t) * t) * t) * t) * t * ex;
            } else {
                erfc_ar = erfc(ar);
            }
            fr = kqq * inv_r * inv_r2 * (erfc_ar + (T)1.1283791670955125739 * ar * ex);  // 2/sqrt(pi)
            if (ENERGY) e = kqq * erfc_ar * inv_r;
        }
    }
    if (SPECIAL) {
        fr *= P.c_w14;
        if (ENERGY) e *= P.c_w14;
    }
}

// Full pair: sums the LJ and Coulomb-family terms with their own cutoffs.
// ljp_i/ljp_j = (sigma-part, eps-part) per atom: Lorentz: sigma/2, geometric: sqrt(sigma); eps: sqrt(eps)
// (0 if the zero shortcut applies). kq_i = ke * q_i.
template <typename T, int COUL, bool UNIFORM, int CUTM, bool ENERGY, bool SPECIAL>
__device__ __forceinline__ void pair_eval(const PairParams<T>& P, T r2, T lj_s_i, T lj_e_i, T lj_s_j, T lj_e_j, T kq_i,
                                          T q_j, T& fr_out, T& e_out) {
    T inv_r, inv_r2;
    if (COUL != COUL_NONE) {
        inv_r = frsqrt(r2);
        inv_r2 = inv_r * inv_r;
    } else {
        inv_r = (T)0;
        inv_r2 = frcp(r2);
    }
    T sig2, eps;
    if (UNIFORM) {
        sig2 = P.uni_sig2;
        eps = P.uni_eps;
    } else {
        T s = P.geo_sigma ? lj_s_i * lj_s_j : lj_s_i + lj_s_j;
        sig2 = s * s;
        eps = lj_e_i * lj_e_j;
    }
    T flj, elj = (T)0;
    if (UNIFORM && CUTM == CUTM_PLAIN) {
        // F/r = i^4 (A i^3 - B), E = i^3 (A/12 i^3 - B/6) with i = 1/r^2: two multiplies fewer than the sigma^2 form
        const T i3 = inv_r2 * inv_r2 * inv_r2;
        flj = (P.uni_A * i3 - P.uni_B) * (i3 * inv_r2);
        if (ENERGY) elj = i3 * (P.uni_A * (T)(1.0 / 12.0) * i3 - P.uni_B * (T)(1.0 / 6.0));
    } else {
        lj_term<T, CUTM, ENERGY>(P, sig2, eps, r2, inv_r2, flj, elj);
    }
    if (SPECIAL) {
        flj *= P.lj_w14;
        if (ENERGY) elj *= P.lj_w14;
    }
    bool in_lj = r2 <= P.lj_rc2;
    T fr = in_lj ? flj : (T)0;
    T e = (ENERGY && in_lj) ? elj : (T)0;
    if (COUL != COUL_NONE) {
        T fc, ec = (T)0;
        coul_term<T, COUL, CUTM, ENERGY, SPECIAL>(P, kq_i * q_j, r2, inv_r, inv_r2, fc, ec);
        bool in_c = r2 <= P.c_rc2;
        fr += in_c ? fc : (T)0;
        if (ENERGY) e += in_c ? ec : (T)0;
    }
    fr_out = fr;
    e_out = e;
}

// Runtime-flag variant for the all-pairs kernel: exclusion and special flags per interaction.
template <typename T, int COUL, int CUTM, bool ENERGY>
__device__ __forceinline__ void pair_eval_rt(const PairParams<T>& P, T r2, T lj_s_i, T lj_e_i, T lj_s_j, T lj_e_j, T kq_i,
                                             T q_j, bool excluded, bool special, T& fr_out, T& e_out) {
    T inv_r = frsqrt(r2);
    T inv_r2 = inv_r * inv_r;
    T fr = (T)0, e = (T)0;
    if (P.has_lj && !(excluded && P.lj_nl)) {
        T s = P.geo_sigma ? lj_s_i * lj_s_j : lj_s_i + lj_s_j;
        T flj, elj = (T)0;
        lj_term<T, CUTM, ENERGY>(P, s * s, lj_e_i * lj_e_j, r2, inv_r2, flj, elj);
        if (special && P.lj_nl) { flj *= P.lj_w14; elj *= P.lj_w14; }
        if (r2 <= P.lj_rc2) { fr += flj; e += elj; }
    }
    if (COUL != COUL_NONE && !(excluded && P.c_nl)) {
        T fc, ec = (T)0;
        if (special && P.c_nl) coul_term<T, COUL, CUTM, ENERGY, true>(P, kq_i * q_j, r2, inv_r, inv_r2, fc, ec);
        else coul_term<T, COUL, CUTM, ENERGY, false>(P, kq_i * q_j, r2, inv_r, inv_r2, fc, ec);
        if (r2 <= P.c_rc2) { fr += fc; e += ec; }
    }
    fr_out = fr;
    e_out = e;
}

}  // namespace mb
