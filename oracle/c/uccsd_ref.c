/* ORACLE (test infrastructure, not product code) — C restatement of the UCCSD state-vector
 * path, matrix-free, OpenMP.  Same mathematics as oracle/statevector.py (validated against
 * it in tests/test_oracle_c.py), fast enough to check the GPU at 20-24 qubits and to serve
 * as the CPU baseline ("port") in bench.py.  PARITY UNPINNED by the reference for UCCSD
 * numbers (see oracle/statevector.py); conventions pinned through the numpy oracle.
 *
 * Follows (reference call sites): ansatz.state / energy / gradient,
 * mqc/solver/dmet_interface.py:279-302 and mqc/solver/dmet_solver_vqechem.py:71; state layout
 * qubit q <-> index bit N-1-q (test/testlog_vqe:56-94).
 *
 * Every sign is evaluated literally, ladder operator by ladder operator (JW parity =
 * number of occupied qubits below the one acted on), not through precomputed masks, so the
 * checker shares no mask algebra with the CUDA library.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;
typedef uint64_t u64;

void orc_set_num_threads(int n) {          /* torchrun exports OMP_NUM_THREADS=1: the CPU baseline undoes that */
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline int occ(u64 x, int n, int q) { return (int)((x >> (n - 1 - q)) & 1ull); }
/* number of occupied qubits with number < q */
static inline int below(u64 x, int n, int q) {
    if (q == 0) return 0;
    return __builtin_popcountll(x >> (n - q));
}
/* apply one ladder operator; returns 0 if annihilated, else +-1, updates *x */
static inline int ladder(u64* x, int n, int q, int create) {
    const int o = occ(*x, n, q);
    if (create == o) return 0;
    const int s = (below(*x, n, q) & 1) ? -1 : 1;
    *x ^= 1ull << (n - 1 - q);
    return s;
}
/* T = a_a^+ a_i  (kind 1, idx = i,a)   or   a_a^+ a_b^+ a_j a_i  (kind 2, idx = i,j,a,b) */
static inline int apply_T(u64* x, int n, int kind, const int* idx) {
    int s;
    if (kind == 1) {
        s = ladder(x, n, idx[0], 0); if (!s) return 0;
        int t = ladder(x, n, idx[1], 1); return s * t;
    }
    s = ladder(x, n, idx[0], 0); if (!s) return 0;
    int t = ladder(x, n, idx[1], 0); if (!t) return 0; s *= t;
    t = ladder(x, n, idx[3], 1); if (!t) return 0; s *= t;
    t = ladder(x, n, idx[2], 1); if (!t) return 0; return s * t;
}

/* psi <- exp(theta (T - T^+)) psi */
void orc_apply_excitation(cplx* psi, int n, int kind, const int* idx, double theta) {
    const u64 dim = 1ull << n;
    const double c = cos(theta), sn = sin(theta);
#pragma omp parallel for schedule(static)
    for (u64 x = 0; x < dim; ++x) {
        u64 y = x;
        const int s = apply_T(&y, n, kind, idx);
        if (!s) continue;
        const cplx a = psi[x], b = psi[y];
        psi[x] = c * a - s * sn * b;
        psi[y] = c * b + s * sn * a;
    }
}

/* returns 2 Re <lam| (T - T^+) |psi>, then un-applies the factor on both vectors */
double orc_adjoint_step(cplx* psi, cplx* lam, int n, int kind, const int* idx, double theta) {
    const u64 dim = 1ull << n;
    const double c = cos(theta), sn = -sin(theta);
    double g = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : g)
    for (u64 x = 0; x < dim; ++x) {
        u64 y = x;
        const int s = apply_T(&y, n, kind, idx);
        if (!s) continue;
        const cplx a = psi[x], b = psi[y], la = lam[x], lb = lam[y];
        /* (A psi)[y] = s psi[x], (A psi)[x] = -s psi[y] */
        g += creal(conj(lb) * (s * a) - conj(la) * (s * b));
        psi[x] = c * a - s * sn * b;  psi[y] = c * b + s * sn * a;
        lam[x] = c * la - s * sn * lb; lam[y] = c * lb + s * sn * la;
    }
    return 2.0 * g;
}

/* lam = sum_t c_t X^x Z^z psi, skipping zero amplitudes (the UCCSD state is sparse) */
void orc_happly(const cplx* psi, cplx* lam, int n, int64_t n_terms, const u64* xm, const u64* zm,
                const double* cre, const double* cim, int64_t t_begin, int64_t t_end) {
    const u64 dim = 1ull << n;
    u64 nnz = 0;
    for (u64 x = 0; x < dim; ++x) nnz += (psi[x] != 0.0);
    u64* nz = (u64*)malloc(sizeof(u64) * (nnz ? nnz : 1));
    u64 k = 0;
    for (u64 x = 0; x < dim; ++x) if (psi[x] != 0.0) nz[k++] = x;
    memset(lam, 0, sizeof(cplx) * dim);
    if (t_end > n_terms) t_end = n_terms;
    for (int64_t t = t_begin; t < t_end; ++t) {
        const cplx ct = cre[t] + I * cim[t];
        const u64 xt = xm[t], zt = zm[t];
#pragma omp parallel for schedule(static)
        for (u64 j = 0; j < nnz; ++j) {
            const u64 x = nz[j];
            const cplx v = (__builtin_popcountll(x & zt) & 1) ? -ct : ct;
            lam[x ^ xt] += v * psi[x];
        }
    }
    free(nz);
}

void orc_reference_state(cplx* psi, int n, u64 ref_index) {
    memset(psi, 0, sizeof(cplx) * (1ull << n));
    psi[ref_index] = 1.0;
}

/* full evaluation: energy and gradient (adjoint method).  psi, lam: caller-provided 2^n buffers */
double orc_energy_grad(cplx* psi, cplx* lam, int n, int n_factors, const int* kind, const int* idx4,
                       const int* theta_index, const double* theta, u64 ref_index, int64_t n_terms,
                       const u64* xm, const u64* zm, const double* cre, const double* cim, double* grad,
                       int n_theta, int want_grad) {
    orc_reference_state(psi, n, ref_index);
    for (int k = 0; k < n_factors; ++k)
        orc_apply_excitation(psi, n, kind[k], idx4 + 4 * k, theta[theta_index[k]]);
    orc_happly(psi, lam, n, n_terms, xm, zm, cre, cim, 0, n_terms);
    const u64 dim = 1ull << n;
    double e = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : e)
    for (u64 x = 0; x < dim; ++x) e += creal(conj(psi[x]) * lam[x]);
    if (want_grad) {
        for (int t = 0; t < n_theta; ++t) grad[t] = 0.0;
        for (int k = n_factors - 1; k >= 0; --k)
            grad[theta_index[k]] += orc_adjoint_step(psi, lam, n, kind[k], idx4 + 4 * k, theta[theta_index[k]]);
    }
    return e;
}

/* ------------------------------------------------------------------------------------------
 * Sector-aware variant.  The UCCSD factors conserve (N_alpha, N_beta) (alpha = even qubits,
 * beta = odd qubits), so only the determinants of one sector are ever non-zero.  These
 * routines sweep the ascending LIST of sector indices instead of all 2^n, and apply H one
 * X-mask group at a time, skipping images outside the sector (their contributions cancel
 * within a group of a number-conserving operator).  Same literal ladder signs as above.
 * psi / lam stay dense 2^n arrays indexed by the OpenFermion basis index.
 * ------------------------------------------------------------------------------------------ */
static void spin_masks(int n, u64* am, u64* bm) {
    u64 a = 0, b = 0;
    for (int q = 0; q < n; ++q) { if (q & 1) b |= 1ull << (n - 1 - q); else a |= 1ull << (n - 1 - q); }
    *am = a; *bm = b;
}

/* all x with popc(x & alpha) == na and popc(x & beta) == nb, ascending; out == NULL: count only */
int64_t orc_sector_list(int n, int na, int nb, u64* out) {
    u64 am, bm;
    spin_masks(n, &am, &bm);
    const u64 dim = 1ull << n;
    int64_t k = 0;
    for (u64 x = 0; x < dim; ++x)
        if (__builtin_popcountll(x & am) == na && __builtin_popcountll(x & bm) == nb) { if (out) out[k] = x; ++k; }
    return k;
}

void orc_apply_excitation_list(cplx* psi, int n, int kind, const int* idx, double theta, const u64* list, int64_t cnt) {
    const double c = cos(theta), sn = sin(theta);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < cnt; ++j) {
        const u64 x = list[j];
        u64 y = x;
        const int s = apply_T(&y, n, kind, idx);
        if (!s) continue;
        const cplx a = psi[x], b = psi[y];
        psi[x] = c * a - s * sn * b;
        psi[y] = c * b + s * sn * a;
    }
}

double orc_adjoint_step_list(cplx* psi, cplx* lam, int n, int kind, const int* idx, double theta, const u64* list, int64_t cnt) {
    const double c = cos(theta), sn = -sin(theta);
    double g = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : g)
    for (int64_t j = 0; j < cnt; ++j) {
        const u64 x = list[j];
        u64 y = x;
        const int s = apply_T(&y, n, kind, idx);
        if (!s) continue;
        const cplx a = psi[x], b = psi[y], la = lam[x], lb = lam[y];
        g += creal(conj(lb) * (s * a) - conj(la) * (s * b));
        psi[x] = c * a - s * sn * b;  psi[y] = c * b + s * sn * a;
        lam[x] = c * la - s * sn * lb; lam[y] = c * lb + s * sn * la;
    }
    return 2.0 * g;
}

static const u64* g_sort_key;
static int cmp_by_key(const void* a, const void* b) {
    const u64 ka = g_sort_key[*(const int64_t*)a], kb = g_sort_key[*(const int64_t*)b];
    if (ka != kb) return ka < kb ? -1 : 1;
    const int64_t ia = *(const int64_t*)a, ib = *(const int64_t*)b;
    return ia < ib ? -1 : (ia > ib);
}

/* lam[list] = (H psi)[list]; terms grouped by X mask; only sector -> sector images */
void orc_happly_sector(const cplx* psi, cplx* lam, int n, int na, int nb, int64_t n_terms, const u64* xm, const u64* zm,
                       const double* cre, const double* cim, const u64* list, int64_t cnt) {
    u64 am, bm;
    spin_masks(n, &am, &bm);
    int64_t* ord = (int64_t*)malloc(sizeof(int64_t) * (n_terms ? n_terms : 1));
    for (int64_t t = 0; t < n_terms; ++t) ord[t] = t;
    g_sort_key = xm;
    qsort(ord, (size_t)n_terms, sizeof(int64_t), cmp_by_key);
    /* group boundaries; groups whose flip changes N_alpha or N_beta never map the sector into itself */
    int64_t* gb = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_terms + 1));
    int64_t ng = 0, t0 = 0;
    while (t0 < n_terms) {
        int64_t t1 = t0;
        const u64 xg = xm[ord[t0]];
        while (t1 < n_terms && xm[ord[t1]] == xg) ++t1;
        const int da = __builtin_popcountll(xg & am), db = __builtin_popcountll(xg & bm);
        if (!((da | db) & 1)) gb[ng++] = t0;
        else { /* drop the group: compact ord[] so that kept groups stay contiguous */ }
        t0 = t1;
    }
    /* second pass: contiguous copy of the kept groups */
    int64_t* ord2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_terms ? n_terms : 1));
    int64_t* gb2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ng + 1));
    int64_t w = 0;
    for (int64_t g = 0; g < ng; ++g) {
        gb2[g] = w;
        const u64 xg = xm[ord[gb[g]]];
        for (int64_t k = gb[g]; k < n_terms && xm[ord[k]] == xg; ++k) ord2[w++] = ord[k];
    }
    gb2[ng] = w;
    /* gather form: (H psi)[y] = sum_g psi[y ^ x_g] * sum_{t in g} c_t (-1)^popc((y ^ x_g) & z_t),
     * evaluated in de-interleaved (alpha string, beta string) coordinates on a compact copy of the
     * sector (13.7 MB at 24 qubits instead of random reads into the 268 MB dense vector) */
    const int no = (n + 1) / 2;
    const size_t nstr = (size_t)1 << no;
    int32_t* rank_a = (int32_t*)malloc(sizeof(int32_t) * nstr);
    int32_t* rank_b = (int32_t*)malloc(sizeof(int32_t) * nstr);
    uint32_t* astr = (uint32_t*)malloc(sizeof(uint32_t) * nstr);
    uint32_t* bstr = (uint32_t*)malloc(sizeof(uint32_t) * nstr);
    int64_t n_a = 0, n_b = 0;
    for (size_t v = 0; v < nstr; ++v) {
        rank_a[v] = rank_b[v] = -1;
        if (__builtin_popcountll(v) == na && (v >> ((n + 1) / 2)) == 0) { rank_a[v] = (int32_t)n_a; astr[n_a++] = (uint32_t)v; }
        if (__builtin_popcountll(v) == nb && (v >> (n / 2)) == 0) { rank_b[v] = (int32_t)n_b; bstr[n_b++] = (uint32_t)v; }
    }
    /* de-interleave: alpha orbital p = qubit 2p = bit n-1-2p;  beta orbital p = qubit 2p+1 */
    uint32_t* gxa = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(ng + 1));
    uint32_t* gxb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(ng + 1));
    uint32_t* tza = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(w + 1));
    uint32_t* tzb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(w + 1));
    double* tcr = (double*)malloc(sizeof(double) * (size_t)(w + 1));
    double* tci = (double*)malloc(sizeof(double) * (size_t)(w + 1));
#define ORC_SPLIT(mask, outa, outb) do { uint32_t a_ = 0, b_ = 0; \
        for (int q_ = 0; q_ < n; ++q_) if (((mask) >> (n - 1 - q_)) & 1ull) { if (q_ & 1) b_ |= 1u << (q_ / 2); else a_ |= 1u << (q_ / 2); } \
        (outa) = a_; (outb) = b_; } while (0)
    for (int64_t g = 0; g < ng; ++g) ORC_SPLIT(xm[ord2[gb2[g]]], gxa[g], gxb[g]);
    for (int64_t k = 0; k < w; ++k) { ORC_SPLIT(zm[ord2[k]], tza[k], tzb[k]); tcr[k] = cre[ord2[k]]; tci[k] = cim[ord2[k]]; }
    cplx* pc = (cplx*)malloc(sizeof(cplx) * (size_t)((n_a * n_b) > 0 ? n_a * n_b : 1));
    u64* xof = (u64*)malloc(sizeof(u64) * (size_t)((n_a * n_b) > 0 ? n_a * n_b : 1));
#pragma omp parallel for schedule(static)
    for (int64_t ia = 0; ia < n_a; ++ia)
        for (int64_t ib = 0; ib < n_b; ++ib) {
            u64 x = 0;
            for (int p = 0; p < no; ++p) {
                if ((astr[ia] >> p) & 1u) x |= 1ull << (n - 1 - 2 * p);
                if ((bstr[ib] >> p) & 1u) x |= 1ull << (n - 2 - 2 * p);
            }
            xof[ia * n_b + ib] = x;
            pc[ia * n_b + ib] = psi[x];
        }
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ia = 0; ia < n_a; ++ia) {
        const uint32_t a = astr[ia];
        for (int64_t ib = 0; ib < n_b; ++ib) {
            const uint32_t b = bstr[ib];
            double ar = 0.0, ai = 0.0;
            for (int64_t g = 0; g < ng; ++g) {
                const uint32_t a2 = a ^ gxa[g], b2 = b ^ gxb[g];
                const int32_t ra = rank_a[a2], rb = rank_b[b2];
                if ((ra | rb) < 0) continue;
                const cplx p = pc[(int64_t)ra * n_b + rb];
                double wr = 0.0, wi = 0.0;
                for (int64_t k = gb2[g]; k < gb2[g + 1]; ++k) {
                    if ((__builtin_popcount(a2 & tza[k]) + __builtin_popcount(b2 & tzb[k])) & 1) { wr -= tcr[k]; wi -= tci[k]; }
                    else { wr += tcr[k]; wi += tci[k]; }
                }
                ar += wr * creal(p) - wi * cimag(p);
                ai += wr * cimag(p) + wi * creal(p);
            }
            lam[xof[ia * n_b + ib]] = ar + I * ai;
        }
    }
    free(rank_a); free(rank_b); free(astr); free(bstr); free(gxa); free(gxb); free(tza); free(tzb); free(tcr); free(tci);
    free(pc); free(xof);
    free(gb); free(gb2); free(ord2);
    free(ord);
}

/* full evaluation on the sector of the reference determinant */
double orc_energy_grad_sector(cplx* psi, cplx* lam, int n, int n_factors, const int* kind, const int* idx4,
                              const int* theta_index, const double* theta, u64 ref_index, int64_t n_terms,
                              const u64* xm, const u64* zm, const double* cre, const double* cim, double* grad,
                              int n_theta, int want_grad, const u64* list, int64_t cnt, double* phase_seconds) {
    u64 am, bm;
    spin_masks(n, &am, &bm);
    const int na = __builtin_popcountll(ref_index & am), nb = __builtin_popcountll(ref_index & bm);
#ifdef _OPENMP
    double t0 = omp_get_wtime();
#else
    double t0 = 0.0;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < cnt; ++j) { psi[list[j]] = 0.0; lam[list[j]] = 0.0; }
    psi[ref_index] = 1.0;
    for (int k = 0; k < n_factors; ++k)
        orc_apply_excitation_list(psi, n, kind[k], idx4 + 4 * k, theta[theta_index[k]], list, cnt);
#ifdef _OPENMP
    double t1 = omp_get_wtime();
#else
    double t1 = 0.0;
#endif
    orc_happly_sector(psi, lam, n, na, nb, n_terms, xm, zm, cre, cim, list, cnt);
    double e = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : e)
    for (int64_t j = 0; j < cnt; ++j) e += creal(conj(psi[list[j]]) * lam[list[j]]);
#ifdef _OPENMP
    double t2 = omp_get_wtime();
#else
    double t2 = 0.0;
#endif
    if (want_grad) {
        for (int t = 0; t < n_theta; ++t) grad[t] = 0.0;
        for (int k = n_factors - 1; k >= 0; --k)
            grad[theta_index[k]] += orc_adjoint_step_list(psi, lam, n, kind[k], idx4 + 4 * k, theta[theta_index[k]], list, cnt);
    }
#ifdef _OPENMP
    double t3 = omp_get_wtime();
#else
    double t3 = 0.0;
#endif
    if (phase_seconds) { phase_seconds[0] = t1 - t0; phase_seconds[1] = t2 - t1; phase_seconds[2] = t3 - t2; }
    return e;
}
