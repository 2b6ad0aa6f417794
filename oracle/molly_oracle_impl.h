/*
 * molly_oracle_impl.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of Molly.jl's pairwise non-bonded hot path, written from the
 * reference's documented behaviour (file:line citations below are relative to
 * /root/reference). It is included twice by molly_oracle.c, once with
 * REAL=double (suffix _f64) and once with REAL=float (suffix _f32), mirroring
 * the way the reference computes everything in the System's float type T
 * (SURVEY.md Appendix A.6).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may call into this. The product (libmollyb200.so) never links it.
 *
 * Parity status: PINNED — see oracle/README.md: checked against the
 * reference's per-pair known answers (test/interactions.jl:61-82, :374-395),
 * MIC/wrap known answers (test/basic.jl:2-38), the 4 602 420-pair count
 * (test/basic.jl:592) and the OpenMM 6mrr golden force/energy files
 * (test/protein.jl:206-276) by tests/test_oracle_*.py.
 */

#ifndef REAL
#error "define REAL and SUF before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* ---- src/spatial.jl:491-500 : vector_1D (select chain, no rounding call) ---- */
static inline REAL FN(vector_1D)(REAL c1, REAL c2, REAL side) {
    REAL v12 = c2 - c1;
    REAL v12_p_sl = v12 + side;
    REAL v12_m_sl = v12 - side;
    if (v12 > (REAL)0) {
        return (v12 < -v12_m_sl) ? v12 : v12_m_sl;
    } else {
        return (-v12 < v12_p_sl) ? v12 : v12_p_sl;
    }
}

/* ---- src/spatial.jl:573-579 : wrap_coord_1D ---- */
static inline REAL FN(wrap_coord_1D)(REAL c, REAL side) {
    if (isinf(side)) return c;
#if REAL_IS_FLOAT
    return c - floorf(c / side) * side;
#else
    return c - floor(c / side) * side;
#endif
}

static inline REAL FN(rsqrt_)(REAL x) {
#if REAL_IS_FLOAT
    return sqrtf(x);
#else
    return sqrt(x);
#endif
}

/* Per-pair parameter block handed to the pair function. */
typedef struct {
    REAL sig_i, eps_i, q_i, lam_i;
    REAL sig_j, eps_j, q_j, lam_j;
} FN(pairparm);

/*
 * One interaction evaluated on one pair. Returns F/r (so that f_vec = fr*dr,
 * dr = c_j - c_i, fs[i] -= f_vec, fs[j] += f_vec; src/force.jl:869-874) and
 * the pair energy through *pe.
 *
 * LJ:       src/interactions/lennard_jones.jl:79-140, mixing src/mixing.jl:5-34
 * Coulomb:  src/interactions/coulomb.jl:71-120
 * CRF:      src/interactions/coulomb.jl:748-814
 * Ewald-real: src/interactions/coulomb.jl:1395-1441; erfc exact or calc_erfc's polynomial (:1384-1393) per approx_erfc
 * cutoffs:  src/cutoffs.jl:15-45 (NoCutoff / DistanceCutoff),
 *           :99-141 (ShiftedPotential), :143-190 (ShiftedForce), :192-253 (CubicSpline, Polynomial; LJ only)
 */
static inline void FN(pair_eval)(const orc_inter_t *in, const FN(pairparm) * p, REAL r2, int special,
                                 REAL *fr_out, REAL *pe_out) {
    REAL r = FN(rsqrt_)(r2);
    REAL fr = 0, pe = 0;
    REAL w = special ? (REAL)in->weight_special : (REAL)1;
    REAL rc = (REAL)in->r_cut;
    if (in->kind == ORC_LJ) {
        /* LJZeroShortcut: mixing.jl:7-11 */
        if (p->eps_i == 0 || p->eps_j == 0 || p->sig_i == 0 || p->sig_j == 0 || p->lam_i == 0 ||
            p->lam_j == 0) {
            *fr_out = 0;
            *pe_out = 0;
            return;
        }
        REAL sig = (in->sigma_mix == ORC_MIX_GEOMETRIC) ? FN(rsqrt_)(p->sig_i * p->sig_j)
                                                         : (p->sig_i + p->sig_j) / 2;
        REAL eps = (in->eps_mix == ORC_MIX_LORENTZ) ? (p->eps_i + p->eps_j) / 2
                                                     : FN(rsqrt_)(p->eps_i * p->eps_j);
        REAL s2 = sig * sig;
        REAL six = (s2 / (r * r));
        six = six * six * six;
        REAL f = (24 * eps / r) * (2 * six * six - six); /* pairwise_force :106-109 */
        REAL e = 4 * eps * (six * six - six);            /* pairwise_pe :137-140 */
        if (in->cutoff_kind == ORC_CUT_NONE) {
            /* nothing */
        } else if (in->cutoff_kind == ORC_CUT_DISTANCE) {
            if (!(r <= rc)) { f = 0; e = 0; }
        } else if (in->cutoff_kind == ORC_CUT_SHIFTED_POTENTIAL) {
            /* cutoffs.jl ShiftedPotential: force unchanged, pe - pe(rc) */
            REAL sc = s2 / (rc * rc);
            sc = sc * sc * sc;
            REAL ec = 4 * eps * (sc * sc - sc);
            e = e - ec;
            if (!(r <= rc)) { f = 0; e = 0; }
        } else if (in->cutoff_kind == ORC_CUT_SHIFTED_FORCE) {
            /* cutoffs.jl ShiftedForce: f - f(rc); pe - (r-rc)*(-f(rc)) - pe(rc) */
            REAL sc = s2 / (rc * rc);
            sc = sc * sc * sc;
            REAL fc = (24 * eps / rc) * (2 * sc * sc - sc);
            REAL ec = 4 * eps * (sc * sc - sc);
            f = f - fc;
            e = e + (r - rc) * fc - ec;
            if (!(r <= rc)) { f = 0; e = 0; }
        } else if (in->cutoff_kind == ORC_CUT_CUBIC_SPLINE || in->cutoff_kind == ORC_CUT_POLYNOMIAL) {
            /* two-point cutoffs, cutoffs.jl:23-29, :39-45: unchanged up to r_act, switched on (r_act, rc], zero beyond */
            REAL ra = (REAL)in->r_act;
            if (!(r <= ra)) {
                REAL t = (r - ra) / (rc - ra);
                if (in->cutoff_kind == ORC_CUT_CUBIC_SPLINE) { /* cutoffs.jl:192-215 */
                    REAL sa = s2 / (ra * ra);
                    sa = sa * sa * sa;
                    REAL pe_act = 4 * eps * (sa * sa - sa);
                    REAL dpe_act = -((24 * eps / ra) * (2 * sa * sa - sa));
                    e = (2 * t * t * t - 3 * t * t + 1) * pe_act + (t * t * t - 2 * t * t + t) * (rc - ra) * dpe_act;
                    f = -(6 * t * t - 6 * t) * pe_act / (rc - ra) - (3 * t * t - 4 * t + 1) * dpe_act;
                } else { /* PolynomialCutoff, cutoffs.jl:241-253 */
                    REAL t2 = t * t, t3 = t2 * t;
                    REAL S = 1 - 6 * t3 * t2 + 15 * t2 * t2 - 10 * t3;
                    REAL dS = (-30 * t2 * t2 + 60 * t3 - 30 * t2) / (rc - ra);
                    REAL f0 = f, e0 = e;
                    e = S * e0;
                    f = S * f0 - dS * e0;
                }
                if (!(r <= rc)) { f = 0; e = 0; }
            }
        }
        fr = (f / r) * w;
        pe = e * w;
    } else if (in->kind == ORC_COULOMB) {
        REAL ke = (REAL)in->coulomb_const;
        REAL kqq = ke * p->q_i * p->q_j;
        REAL f = kqq / (r * r);
        REAL e = kqq * (1 / r);
        if (in->cutoff_kind == ORC_CUT_DISTANCE) {
            if (!(r <= rc)) { f = 0; e = 0; }
        } else if (in->cutoff_kind == ORC_CUT_SHIFTED_POTENTIAL) {
            e = e - kqq * (1 / rc);
            if (!(r <= rc)) { f = 0; e = 0; }
        } else if (in->cutoff_kind == ORC_CUT_SHIFTED_FORCE) {
            REAL fc = kqq / (rc * rc);
            REAL ec = kqq * (1 / rc);
            f = f - fc;
            e = e + (r - rc) * fc - ec;
            if (!(r <= rc)) { f = 0; e = 0; }
        } else if (in->cutoff_kind == ORC_CUT_CUBIC_SPLINE || in->cutoff_kind == ORC_CUT_POLYNOMIAL) {
            /* two-point cutoffs on V = kqq / r (generic force_cutoff / pe_cutoff, cutoffs.jl:23-29, :39-45) */
            REAL ra = (REAL)in->r_act;
            if (!(r <= ra)) {
                REAL t = (r - ra) / (rc - ra);
                if (in->cutoff_kind == ORC_CUT_CUBIC_SPLINE) { /* cutoffs.jl:192-215 */
                    REAL pe_act = kqq * (1 / ra);
                    REAL dpe_act = -(kqq / (ra * ra));
                    e = (2 * t * t * t - 3 * t * t + 1) * pe_act + (t * t * t - 2 * t * t + t) * (rc - ra) * dpe_act;
                    f = -(6 * t * t - 6 * t) * pe_act / (rc - ra) - (3 * t * t - 4 * t + 1) * dpe_act;
                } else { /* PolynomialCutoff, cutoffs.jl:241-253 */
                    REAL t2 = t * t, t3 = t2 * t;
                    REAL S = 1 - 6 * t3 * t2 + 15 * t2 * t2 - 10 * t3;
                    REAL dS = (-30 * t2 * t2 + 60 * t3 - 30 * t2) / (rc - ra);
                    REAL f0 = f, e0 = e;
                    e = S * e0;
                    f = S * f0 - dS * e0;
                }
                if (!(r <= rc)) { f = 0; e = 0; }
            }
        }
        fr = (f / r) * w;
        pe = e * w;
    } else if (in->kind == ORC_CRF) {
        REAL ke = (REAL)in->coulomb_const;
        REAL kqq = ke * p->q_i * p->q_j;
        REAL eps_s = (REAL)in->solvent_dielectric;
        REAL krf, crf;
        if (special) {
            krf = 0;
            crf = 0;
        } else if (isinf(in->solvent_dielectric)) {
            krf = 1 / (2 * rc * rc * rc);
            crf = 3 * (1 / (2 * rc));
        } else {
            krf = (1 / (rc * rc * rc)) * (eps_s - 1) / (2 * eps_s + 1);
            crf = (1 / rc) * (3 * eps_s) / (2 * eps_s + 1);
        }
        REAL f = kqq * (1 / r - 2 * krf * r2) * (1 / r2); /* already F/r */
        REAL e = kqq * (1 / r + krf * r2 - crf);
        if (!(r <= rc)) { f = 0; e = 0; }
        fr = f * w;
        pe = e * w;
    } else if (in->kind == ORC_EWALD_REAL) {
        REAL ke = (REAL)in->coulomb_const;
        REAL kqq = ke * p->q_i * p->q_j;
        if (special) {
            /* special pairs: plain weighted Coulomb, coulomb.jl:1419-1423 */
            REAL f = kqq / (r * r);
            REAL e = kqq / r;
            if (!(r <= rc)) { f = 0; e = 0; }
            fr = (f / r) * w;
            pe = e * w;
        } else {
            double a = in->ewald_alpha;
            double ar = a * (double)r;
            double ex = exp(-ar * ar);
            double erfc_ar;
            if (in->approx_erfc) {
                /* calc_erfc, coulomb.jl:1384-1393: Abramowitz & Stegun 7.1.26 (the reference's default) */
                double t = 1.0 / (1.0 + 0.3275911 * ar);
                erfc_ar = (0.254829592 + (-0.284496736 + (1.421413741 + (-1.453152027 + 1.061405429 * t) * t) * t) * t) * t * ex;
            } else {
                erfc_ar = erfc(ar);
            }
            REAL f = (REAL)((double)kqq * (erfc_ar + 2.0 * ar * ex / 1.7724538509055160273) /
                            ((double)r2 * (double)r));
            REAL e = (REAL)((double)kqq * erfc_ar / (double)r);
            if (!(r <= rc)) { f = 0; e = 0; }
            fr = f;
            pe = e;
        }
    }
    *fr_out = fr;
    *pe_out = pe;
}

/* exclusion / special lookup: CSR of partners per atom (both directions), sorted */
static inline int FN(csr_has)(const int64_t *ptr, const int32_t *idx, int32_t i, int32_t j) {
    if (!ptr) return 0;
    int64_t lo = ptr[i], hi = ptr[i + 1];
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int32_t v = idx[mid];
        if (v == j) return 1;
        if (v < j) lo = mid + 1; else hi = mid;
    }
    return 0;
}

/*
 * Evaluate all interactions of the tuple on one pair and accumulate
 * (src/force.jl:857-881; energy src/energy.jl:273-289).
 * which: 0 = every interaction, 1 = only use_neighbors==0, 2 = only use_neighbors==1
 */
static inline void FN(pair_accumulate)(const orc_system_t *s, const REAL *coords, int32_t i, int32_t j,
                                       int special, int excluded, int which, REAL *fs, double *pe_acc, double *vir) {
    const REAL *ci = coords + 3 * (size_t)i, *cj = coords + 3 * (size_t)j;
    REAL dr[3];
    for (int d = 0; d < 3; d++) dr[d] = FN(vector_1D)(ci[d], cj[d], (REAL)s->box[d]);
    REAL r2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
    FN(pairparm) p;
    const REAL *sig = (const REAL *)s->sigma, *eps = (const REAL *)s->eps, *q = (const REAL *)s->charge;
    const REAL *lam = (const REAL *)s->lambda;
    p.sig_i = sig[i]; p.eps_i = eps[i]; p.q_i = q[i]; p.lam_i = lam ? lam[i] : (REAL)1;
    p.sig_j = sig[j]; p.eps_j = eps[j]; p.q_j = q[j]; p.lam_j = lam ? lam[j] : (REAL)1;
    REAL frsum = 0;
    REAL pesum = 0;
    for (int k = 0; k < s->n_inters; k++) {
        const orc_inter_t *in = &s->inters[k];
        if (which == 1 && in->use_neighbors) continue;
        if (which == 2 && !in->use_neighbors) continue;
        /* eligibility / special flags only exist for interactions that go through the neighbour list
         * (src/force.jl:828-855: the use_neighbors=false loop visits every i<j with special=false) */
        if (in->use_neighbors && excluded) continue;
        REAL fr, pe;
        FN(pair_eval)(in, &p, r2, in->use_neighbors ? special : 0, &fr, &pe);
        frsum += fr;
        pesum += pe;
    }
    if (fs) {
        for (int d = 0; d < 3; d++) {
            REAL f = frsum * dr[d];
            fs[3 * (size_t)i + d] -= f;
            fs[3 * (size_t)j + d] += f;
        }
    }
    if (vir) {
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) vir[3 * a + b] += (double)(dr[a] * (frsum * dr[b]));
    }
    if (pe_acc) *pe_acc += (double)pesum;
}

/*
 * Brute-force O(N^2) evaluation over all i<j pairs. Interactions with
 * use_neighbors=true skip excluded pairs and see the special flag; interactions
 * with use_neighbors=false see every pair with special=false (src/force.jl:828-855).
 * This is the semantic definition of the path:
 * the neighbour structures only pre-filter. Threads own private force copies
 * that are reduced at the end (src/force.jl:886-969, :808-826).
 */
int FN(orc_forces_allpairs)(const orc_system_t *s, const void *coords_v, void *fs_v, double *pe_out,
                            double *virial_out /* 9 or NULL */, int n_threads) {
    const REAL *coords = (const REAL *)coords_v;
    REAL *fs = (REAL *)fs_v;
    int64_t n = s->n_atoms;
    if (n_threads < 1) n_threads = 1;
    REAL *scratch = NULL;
    if (fs) {
        scratch = (REAL *)calloc((size_t)n_threads * 3 * n, sizeof(REAL));
        if (!scratch) return -1;
    }
    double pe_total = 0;
    double vir_total[9] = {0};
#pragma omp parallel num_threads(n_threads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        REAL *myfs = scratch ? scratch + (size_t)tid * 3 * n : NULL;
        double mype = 0;
        double myvir[9] = {0};
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < n; i++) {
            for (int64_t j = i + 1; j < n; j++) {
                int excluded = FN(csr_has)(s->excl_ptr, s->excl_idx, (int32_t)i, (int32_t)j);
                int special = FN(csr_has)(s->spec_ptr, s->spec_idx, (int32_t)i, (int32_t)j);
                FN(pair_accumulate)(s, coords, (int32_t)i, (int32_t)j, special, excluded, 0, myfs,
                                    pe_out ? &mype : NULL, virial_out ? myvir : NULL);
            }
        }
#pragma omp critical
        {
            pe_total += mype;
            for (int k = 0; k < 9; k++) vir_total[k] += myvir[k];
        }
    }
    if (fs) {
        /* reduce_force_chunks!: src/force.jl:808-826 (adds into fs) */
        for (int t = 0; t < n_threads; t++)
            for (int64_t k = 0; k < 3 * n; k++) fs[k] += scratch[(size_t)t * 3 * n + k];
        free(scratch);
    }
    if (pe_out) *pe_out += pe_total;
    if (virial_out)
        for (int k = 0; k < 9; k++) virial_out[k] += vir_total[k];
    return 0;
}

/*
 * Neighbour list in the reference's format (types.jl:611-654): entries
 * (i, j, special) with i<j (0-based here), built by a cell list with radius
 * r_list, filtered by eligibility (src/neighbors.jl:609-623, :665-693).
 * Geometry follows CellListMap.jl 0.10 semantics as pinned by the reference's
 * tests (test/basic.jl:512-641): every unordered pair whose minimum-image
 * distance is <= r_list appears exactly once (the comparison is d2 <= cutoff2,
 * in double here as CellListMap computes in the coordinate type; ties at exactly
 * r_list are not exercised by the reference's tests).
 * Returns number of entries; *out is malloc'd (free with orc_free).
 */
int64_t FN(orc_neighbor_list_mt)(const orc_system_t *s, const void *coords_v, double r_list,
                                 orc_nl_entry_t **out, int n_threads_nl) {
    const REAL *coords = (const REAL *)coords_v;
    int64_t n = s->n_atoms;
    int nc[3];
    double cs[3];
    for (int d = 0; d < 3; d++) {
        nc[d] = (int)floor(s->box[d] / r_list);
        if (nc[d] < 1) nc[d] = 1;
        cs[d] = s->box[d] / nc[d];
    }
    int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * n);
    int64_t *cstart = (int64_t *)calloc(ncell + 1, sizeof(int64_t));
    int32_t *sorted = (int32_t *)malloc(sizeof(int32_t) * n);
    for (int64_t i = 0; i < n; i++) {
        int c[3];
        for (int d = 0; d < 3; d++) {
            double x = coords[3 * i + d];
            x -= floor(x / s->box[d]) * s->box[d];
            c[d] = (int)(x / cs[d]);
            if (c[d] >= nc[d]) c[d] = nc[d] - 1;
            if (c[d] < 0) c[d] = 0;
        }
        cell_of[i] = (c[2] * nc[1] + c[1]) * nc[0] + c[0];
        cstart[cell_of[i] + 1]++;
    }
    for (int64_t c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * ncell);
    memcpy(fill, cstart, sizeof(int64_t) * ncell);
    for (int64_t i = 0; i < n; i++) sorted[fill[cell_of[i]]++] = (int32_t)i;
    free(fill);

    /* The pair search is threaded like the reference's (CellListMap.map_pairwise! with parallel = n_threads > 1,
     * src/neighbors.jl:676-680): contiguous blocks of atoms are claimed by threads, every block fills its own buffer,
     * and the buffers are concatenated in block order, so the list is the same as a serial scan's whatever the
     * thread count. */
    double rl2 = r_list * r_list;
    int nthr = n_threads_nl > 0 ? n_threads_nl : omp_get_max_threads();
    if (nthr < 1) nthr = 1;
    int64_t nblk = (int64_t)nthr * 8;
    if (nblk > n) nblk = n > 0 ? n : 1;
    orc_nl_entry_t **bbuf = (orc_nl_entry_t **)calloc(nblk, sizeof(orc_nl_entry_t *));
    int64_t *bcnt = (int64_t *)calloc(nblk + 1, sizeof(int64_t));
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthr)
    for (int64_t blk = 0; blk < nblk; blk++) {
        int64_t i0 = n * blk / nblk, i1 = n * (blk + 1) / nblk;
        int64_t cap = (i1 - i0) * 64 + 1024, cnt = 0;
        orc_nl_entry_t *list = (orc_nl_entry_t *)malloc(sizeof(orc_nl_entry_t) * cap);
        /* distinct neighbour cells per dimension (handles nc < 3 without double counting) */
        for (int64_t i = i0; i < i1; i++) {
            int32_t ci = cell_of[i];
            int cx = ci % nc[0], cy = (ci / nc[0]) % nc[1], cz = ci / (nc[0] * nc[1]);
            int lst[3][3], ln[3];
            int cc[3] = {cx, cy, cz};
            for (int d = 0; d < 3; d++) {
                ln[d] = 0;
                for (int o = -1; o <= 1; o++) {
                    int v = (cc[d] + o + nc[d]) % nc[d];
                    int dup = 0;
                    for (int k = 0; k < ln[d]; k++) dup |= (lst[d][k] == v);
                    if (!dup) lst[d][ln[d]++] = v;
                }
            }
            for (int a = 0; a < ln[2]; a++)
                for (int b = 0; b < ln[1]; b++)
                    for (int c = 0; c < ln[0]; c++) {
                        int64_t cj = ((int64_t)lst[2][a] * nc[1] + lst[1][b]) * nc[0] + lst[0][c];
                        for (int64_t k = cstart[cj]; k < cstart[cj + 1]; k++) {
                            int32_t j = sorted[k];
                            if (j <= i) continue;
                            double d2 = 0;
                            for (int d = 0; d < 3; d++) {
                                double v = (double)FN(vector_1D)(coords[3 * i + d], coords[3 * (int64_t)j + d],
                                                                 (REAL)s->box[d]);
                                d2 += v * v;
                            }
                            if (d2 > rl2) continue;
                            if (FN(csr_has)(s->excl_ptr, s->excl_idx, (int32_t)i, j)) continue;
                            if (cnt == cap) {
                                cap *= 2;
                                list = (orc_nl_entry_t *)realloc(list, sizeof(orc_nl_entry_t) * cap);
                            }
                            list[cnt].i = (int32_t)i;
                            list[cnt].j = j;
                            list[cnt].special = FN(csr_has)(s->spec_ptr, s->spec_idx, (int32_t)i, j);
                            cnt++;
                        }
                    }
        }
        bbuf[blk] = list;
        bcnt[blk + 1] = cnt;
    }
    for (int64_t blk = 0; blk < nblk; blk++) bcnt[blk + 1] += bcnt[blk];
    int64_t cnt = bcnt[nblk];
    orc_nl_entry_t *list = (orc_nl_entry_t *)malloc(sizeof(orc_nl_entry_t) * (cnt > 0 ? cnt : 1));
#pragma omp parallel for schedule(static) num_threads(nthr)
    for (int64_t blk = 0; blk < nblk; blk++) {
        memcpy(list + bcnt[blk], bbuf[blk], sizeof(orc_nl_entry_t) * (bcnt[blk + 1] - bcnt[blk]));
        free(bbuf[blk]);
    }
    free(bbuf);
    free(bcnt);
    free(cell_of);
    free(cstart);
    free(sorted);
    *out = list;
    return cnt;
}

int64_t FN(orc_neighbor_list)(const orc_system_t *s, const void *coords_v, double r_list, orc_nl_entry_t **out) {
    return FN(orc_neighbor_list_mt)(s, coords_v, r_list, out, 0);
}

/*
 * Force/energy over a neighbour list, the reference's multi-threaded CPU
 * algorithm (src/force.jl:886-969): interactions with use_neighbors=false loop
 * over all i<j, the others over the list in blocks of 512 claimed by threads;
 * per-thread force copies are reduced afterwards (:808-826).
 */
int FN(orc_forces_nl)(const orc_system_t *s, const void *coords_v, const orc_nl_entry_t *list, int64_t n_list,
                      void *fs_v, double *pe_out, double *virial_out, int n_threads, void *scratch_v) {
    const REAL *coords = (const REAL *)coords_v;
    REAL *fs = (REAL *)fs_v;
    int64_t n = s->n_atoms;
    if (n_threads < 1) n_threads = 1;
    REAL *scratch = (REAL *)scratch_v;
    int own_scratch = 0;
    if (fs && !scratch) {
        scratch = (REAL *)malloc((size_t)n_threads * 3 * n * sizeof(REAL));
        own_scratch = 1;
    }
    int any_nonl = 0, any_nl = 0;
    for (int k = 0; k < s->n_inters; k++) {
        if (s->inters[k].use_neighbors) any_nl = 1; else any_nonl = 1;
    }
    double pe_total = 0;
    double vir_total[9] = {0};
    int64_t n_blocks = (n_list + 511) / 512;
#pragma omp parallel num_threads(n_threads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        REAL *myfs = fs ? scratch + (size_t)tid * 3 * n : NULL;
        if (myfs) memset(myfs, 0, sizeof(REAL) * 3 * n);
        double mype = 0;
        double myvir[9] = {0};
        if (any_nonl) {
#pragma omp for schedule(dynamic, 16)
            for (int64_t i = 0; i < n; i++)
                for (int64_t j = i + 1; j < n; j++)
                    FN(pair_accumulate)(s, coords, (int32_t)i, (int32_t)j, 0, 0, 1, myfs, pe_out ? &mype : NULL,
                                        virial_out ? myvir : NULL);
        }
        if (any_nl) {
#pragma omp for schedule(dynamic, 1)
            for (int64_t b = 0; b < n_blocks; b++) {
                int64_t lo = b * 512, hi = lo + 512 < n_list ? lo + 512 : n_list;
                for (int64_t k = lo; k < hi; k++)
                    FN(pair_accumulate)(s, coords, list[k].i, list[k].j, list[k].special, 0, 2, myfs,
                                        pe_out ? &mype : NULL, virial_out ? myvir : NULL);
            }
        }
#pragma omp critical
        {
            pe_total += mype;
            for (int k = 0; k < 9; k++) vir_total[k] += myvir[k];
        }
        if (fs) {
#pragma omp barrier
#pragma omp for schedule(static)
            for (int64_t k = 0; k < 3 * n; k++) {
                REAL acc = 0;
                for (int t = 0; t < n_threads; t++) acc += scratch[(size_t)t * 3 * n + k];
                fs[k] += acc;
            }
        }
    }
    if (own_scratch) free(scratch);
    if (pe_out) *pe_out += pe_total;
    if (virial_out)
        for (int k = 0; k < 9; k++) virial_out[k] += vir_total[k];
    return 0;
}

/* src/spatial.jl:901-916 : remove_CM_motion! (sequential sum, like the CPU path) */
void FN(orc_remove_cm)(const orc_system_t *s, void *vel_v) {
    REAL *v = (REAL *)vel_v;
    const REAL *m = (const REAL *)s->mass;
    int64_t n = s->n_atoms;
    REAL p[3] = {0, 0, 0};
    REAL mt = 0;
    for (int64_t i = 0; i < n; i++) {
        for (int d = 0; d < 3; d++) p[d] += v[3 * i + d] * m[i];
        mt += m[i];
    }
    for (int d = 0; d < 3; d++) p[d] /= mt;
    for (int64_t i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) v[3 * i + d] -= p[d];
}

/*
 * VelocityVerlet simulate! (src/simulators.jl:547-668), no coupling, no
 * constraints, no loggers: wrap -> [CM removal] -> neighbours -> F0; per step
 * kick, drift, wrap, forces, kick, CM removal (if step % remove_cm == 0),
 * find_neighbors when step % nl_every == 0 (src/neighbors.jl:671).
 * r_list <= 0 means "no neighbour list": all interactions brute force.
 */
int FN(orc_simulate_vv)(const orc_system_t *s, void *coords_v, void *vel_v, double dt_d, int64_t n_steps,
                        int remove_cm_every, double r_list, int nl_every, int n_threads, double *pe_final) {
    REAL *x = (REAL *)coords_v, *v = (REAL *)vel_v;
    const REAL *m = (const REAL *)s->mass;
    int64_t n = s->n_atoms;
    REAL dt = (REAL)dt_d;
    REAL dt2 = dt / 2;
    if (n_threads < 1) n_threads = 1;
    REAL *f = (REAL *)calloc(3 * n, sizeof(REAL));
    REAL *scratch = (REAL *)malloc((size_t)n_threads * 3 * n * sizeof(REAL));
    orc_nl_entry_t *list = NULL;
    int64_t n_list = 0;
    for (int64_t i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) x[3 * i + d] = FN(wrap_coord_1D)(x[3 * i + d], (REAL)s->box[d]);
    if (remove_cm_every) FN(orc_remove_cm)(s, v);
    if (r_list > 0) n_list = FN(orc_neighbor_list_mt)(s, x, r_list, &list, n_threads);
#define ORC_FORCE_EVAL()                                                                              \
    do {                                                                                              \
        memset(f, 0, sizeof(REAL) * 3 * n);                                                           \
        if (r_list > 0)                                                                               \
            FN(orc_forces_nl)(s, x, list, n_list, f, NULL, NULL, n_threads, scratch);                 \
        else                                                                                          \
            FN(orc_forces_allpairs)(s, x, f, NULL, NULL, n_threads);                                  \
    } while (0)
    ORC_FORCE_EVAL();
    for (int64_t step = 1; step <= n_steps; step++) {
        for (int64_t i = 0; i < n; i++) {
            for (int d = 0; d < 3; d++) {
                REAL a = (m[i] == 0) ? (REAL)0 : f[3 * i + d] / m[i]; /* calc_accels force.jl:17 */
                v[3 * i + d] += a * dt2;
                x[3 * i + d] += v[3 * i + d] * dt;
                x[3 * i + d] = FN(wrap_coord_1D)(x[3 * i + d], (REAL)s->box[d]);
            }
        }
        ORC_FORCE_EVAL();
        for (int64_t i = 0; i < n; i++)
            for (int d = 0; d < 3; d++) {
                REAL a = (m[i] == 0) ? (REAL)0 : f[3 * i + d] / m[i];
                v[3 * i + d] += a * dt2;
            }
        if (remove_cm_every && step % remove_cm_every == 0) FN(orc_remove_cm)(s, v);
        if (r_list > 0 && nl_every > 0 && step % nl_every == 0) {
            free(list);
            n_list = FN(orc_neighbor_list_mt)(s, x, r_list, &list, n_threads);
        }
    }
    if (pe_final) {
        double pe = 0;
        if (r_list > 0)
            FN(orc_forces_nl)(s, x, list, n_list, NULL, &pe, NULL, n_threads, NULL);
        else
            FN(orc_forces_allpairs)(s, x, NULL, &pe, NULL, n_threads);
        *pe_final = pe;
    }
#undef ORC_FORCE_EVAL
    free(list);
    free(f);
    free(scratch);
    return 0;
}

#undef FN
#undef CAT
#undef CAT_
