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
