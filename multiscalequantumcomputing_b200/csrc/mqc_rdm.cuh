// K4 on the compact sector matrix C[alpha string][beta string]: 1- and 2-RDM without atomics.
//
// Reference semantics: mqc/tools/rdm.py:131-138 (make_rdm1) and :157-172 (make_rdm12 +
// reorder_rdm), i.e. <p^+ q r^+ s> accumulated over all determinants and link pairs, then
// rdm2[:, k, k, :] -= rdm1.T.  With E_rs = sum_sigma a^+_{r sigma} a_{s sigma} and
//     D[rs][K] = <K| E_rs |psi>                                   (n^2 x N_det, 58 % dense at n = 12)
// the spin-summed two-body density is a Gram matrix over the determinants,
//     <E_pq E_rs> = sum_K conj(D[qp][K]) D[rs][K],      <E_ps> = sum_K conj(C[K]) D[ps][K],
//     dm2[p,q,r,s] = <E_pq E_rs> - delta_qr <E_ps>,     dm1[p,q] = <E_qp>            (pyscf order)
// so the N_det * links^2 scattered atomic updates of the literal loop (81 ms at 24 qubits, 50 s at
// 32) become a rank-N_det update of an n^2 x n^2 matrix: a CTA builds D for a block of 32
// consecutive determinants of one alpha row in shared memory (the alpha part reads 32 contiguous
// amplitudes of the source row, the beta part gathers within the own row) and adds its outer
// product to a T x T register tile per thread; CTAs keep private accumulators over all their
// blocks and write them out once (k_rdm_finish sums the partials).  FP64 FMA bound:
// 2 n^4 N_det FMA = 3.5e10 at 24 qubits.
#pragma once
#include <cuda_runtime.h>

#include "mqc_common.h"

#define MQC_RDM_KB 32          // determinants per block
#define MQC_RDM_THREADS 256    // 16 x 16 threads, T x T accumulators each

struct MqcRdmC {
    const void* C;             // compact state, canonical order: double2, or double with real amplitudes
    const unsigned* aorb;      // [nA] alpha strings (orbital space), ascending
    const unsigned* borb;      // [nB]
    const unsigned* rank_a;    // [2^n] string -> row (0xffffffff: not in the sector)
    const unsigned* rank_b;
    unsigned nA, nB;
    unsigned pitch;            // elements between consecutive rows of C
    int n;                     // spatial orbitals
    int n2;                    // n * n
    int nblk;                  // blocks of BS = 16 T rows the n^2 x n^2 matrix is cut into
    double* gpart;             // [gridDim.x][nblk * nblk][BS * BS]
    double* g1part;            // [gridDim.x][n2]
};

// <K| a^+_r a_s |J> for same-spin orbitals r != s of the spin species with string `own` (K's), the other
// species' string `other`; JW order: qubit 2 p (alpha), 2 p + 1 (beta).  Returns the sign bit.
__device__ __forceinline__ unsigned mqc_rdm_sign(const unsigned own, const unsigned other, const int r, const int s, const bool beta) {
    const int lo = r < s ? r : s, hi = r < s ? s : r;
    const unsigned strictly = ((1u << hi) - 1u) & ~((2u << lo) - 1u);          // lo < p < hi
    // alpha excitation: beta orbitals lo <= b < hi lie between; beta excitation: alpha orbitals lo < a <= hi
    const unsigned cross = beta ? (((2u << hi) - 1u) & ~((2u << lo) - 1u)) : (((1u << hi) - 1u) & ~((1u << lo) - 1u));
    return (unsigned)(__popc(own & strictly) + __popc(other & cross)) & 1u;
}

template <bool REAL> struct MqcRdmAmp { typedef double2 T; };
template <> struct MqcRdmAmp<true> { typedef double T; };
__device__ __forceinline__ double2 mqc_rdm_zero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double mqc_rdm_zero(double) { return 0.0; }
__device__ __forceinline__ void mqc_rdm_add(double2& v, const double2 c, const unsigned neg) {
    if (neg) { v.x -= c.x; v.y -= c.y; } else { v.x += c.x; v.y += c.y; }
}
__device__ __forceinline__ void mqc_rdm_add(double& v, const double c, const unsigned neg) { v += neg ? -c : c; }
// acc + Re(conj(a) b)
__device__ __forceinline__ double mqc_rdm_dot(const double2 a, const double2 b, const double acc) { return fma(a.x, b.x, fma(a.y, b.y, acc)); }
__device__ __forceinline__ double mqc_rdm_dot(const double a, const double b, const double acc) { return fma(a, b, acc); }

// REAL: real amplitudes (8-byte compact storage): D is real, half the shared memory and half the FMAs
template <int T, bool REAL>
__global__ void __launch_bounds__(MQC_RDM_THREADS, 1)
k_rdm_gram(const __grid_constant__ MqcRdmC R) {
    typedef typename MqcRdmAmp<REAL>::T AT;
    const AT* __restrict__ Cst = static_cast<const AT*>(R.C);
    constexpr int BS = 16 * T;
    extern __shared__ __align__(16) unsigned char rdm_smem[];
    AT* Dx = reinterpret_cast<AT*>(rdm_smem);                      // [KB][BS]  rows of block bx, transposed
    AT* Dy = Dx + MQC_RDM_KB * BS;                                       // [KB][BS]  rows of block by (== Dx when bx == by)
    __shared__ AT ck[MQC_RDM_KB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int bx = blockIdx.y / R.nblk, by = blockIdx.y - bx * R.nblk;
    const bool same = bx == by;
    if (same) Dy = Dx;
    const int n = R.n, n2 = R.n2;
    const unsigned nB = R.nB;
    const unsigned cols_blocks = (nB + MQC_RDM_KB - 1) / MQC_RDM_KB;
    const unsigned long long n_blocks = (unsigned long long)R.nA * cols_blocks;
    double acc[T][T];
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) acc[a][b] = 0.0;
    double acc1 = 0.0;                                                        // <E_rs> for rs = bx * BS + tid (by == 0 only)

    for (unsigned long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const unsigned ia = (unsigned)(blk / cols_blocks), ib0 = (unsigned)(blk - (unsigned long long)ia * cols_blocks) * MQC_RDM_KB;
        const unsigned Ka = __ldg(R.aorb + ia);
        const int kv = min((int)MQC_RDM_KB, (int)(nB - ib0));                  // valid determinants of this block
        const AT* __restrict__ Crow = Cst + (size_t)ia * R.pitch;
        __syncthreads();                                                      // previous block's Gram pass is done with Dx / Dy
        if (tid < MQC_RDM_KB) ck[tid] = tid < kv ? Crow[ib0 + tid] : mqc_rdm_zero(AT());
        for (int half = 0; half < (same ? 1 : 2); ++half) {
            AT* D = half ? Dy : Dx;
            const int r0 = (half ? by : bx) * BS;
            for (int e = tid; e < MQC_RDM_KB * BS; e += MQC_RDM_THREADS) {
                const int i = e % BS, k = e / BS;
                const int rs = r0 + i;
                AT val = mqc_rdm_zero(AT());
                if (rs < n2 && k < kv) {
                    const int r = rs / n, s = rs - r * n;
                    const unsigned Kb = __ldg(R.borb + ib0 + k);
                    // alpha: K has r_alpha, J = K with r -> s
                    if ((Ka >> r) & 1u) {
                        if (r == s) mqc_rdm_add(val, Crow[ib0 + k], 0u);
                        else if (!((Ka >> s) & 1u)) {
                            const unsigned Ja = Ka ^ (1u << r) ^ (1u << s);
                            mqc_rdm_add(val, Cst[(size_t)__ldg(R.rank_a + Ja) * R.pitch + ib0 + k], mqc_rdm_sign(Ka, Kb, r, s, false));
                        }
                    }
                    if ((Kb >> r) & 1u) {
                        if (r == s) mqc_rdm_add(val, Crow[ib0 + k], 0u);
                        else if (!((Kb >> s) & 1u)) {
                            const unsigned Jb = Kb ^ (1u << r) ^ (1u << s);
                            mqc_rdm_add(val, Crow[__ldg(R.rank_b + Jb)], mqc_rdm_sign(Kb, Ka, r, s, true));
                        }
                    }
                }
                D[k * BS + i] = val;
            }
        }
        __syncthreads();
        // 1-RDM row: <E_rs> += sum_k Re(conj(C_k) D[rs][k])
        if (by == 0 && tid < BS) {
            double a1 = 0.0;
#pragma unroll 8
            for (int k = 0; k < MQC_RDM_KB; ++k) a1 = mqc_rdm_dot(ck[k], Dx[k * BS + tid], a1);
            acc1 += a1;
        }
        // Gram update: acc[a][b] += sum_k Re(conj(Dx[ty + 16 a][k]) Dy[tx + 16 b][k])
#pragma unroll 1
        for (int k = 0; k < MQC_RDM_KB; ++k) {
            AT yb[T];
#pragma unroll
            for (int b = 0; b < T; ++b) yb[b] = Dy[k * BS + tx + 16 * b];
#pragma unroll
            for (int a = 0; a < T; ++a) {
                const AT xa = Dx[k * BS + ty + 16 * a];
#pragma unroll
                for (int b = 0; b < T; ++b) acc[a][b] = mqc_rdm_dot(xa, yb[b], acc[a][b]);
            }
        }
    }
    double* out = R.gpart + ((size_t)blockIdx.x * R.nblk * R.nblk + blockIdx.y) * BS * BS;
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) out[(size_t)(ty + 16 * a) * BS + tx + 16 * b] = acc[a][b];
    if (by == 0 && tid < BS && bx * BS + tid < n2) R.g1part[(size_t)blockIdx.x * n2 + bx * BS + tid] = acc1;
}

// dm2[p,q,r,s] = G[(q,p)][(r,s)] - delta_qr g1[(p,s)],  dm1[p,q] = g1[(q,p)];  G, g1 = sums of the partials
__global__ void k_rdm_finish(const double* __restrict__ gpart, const double* __restrict__ g1part, const int n_groups, const int n,
                             const int nblk, const int BS, double* __restrict__ dm1, double* __restrict__ dm2) {
    const int n2 = n * n;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)n2) {
        double v = 0.0;
        for (int g = 0; g < n_groups; ++g) v += g1part[(size_t)g * n2 + idx];
        const int q = (int)idx / n, p = (int)idx - q * n;                     // idx = (q, p)
        dm1[p * n + q] = v;
    }
    if (!dm2 || idx >= (size_t)n2 * n2) return;
    const int x = (int)(idx / n2), y = (int)(idx - (size_t)x * n2);            // x = (q, p), y = (r, s)
    const int q = x / n, p = x - q * n, r = y / n, s = y - r * n;
    const int bx = x / BS, by = y / BS;
    const size_t off = ((size_t)bx * nblk + by) * BS * BS + (size_t)(x - bx * BS) * BS + (y - by * BS);
    const size_t stride = (size_t)nblk * nblk * BS * BS;
    double v = 0.0;
    for (int g = 0; g < n_groups; ++g) v += gpart[(size_t)g * stride + off];
    if (q == r) {
        double e = 0.0;
        for (int g = 0; g < n_groups; ++g) e += g1part[(size_t)g * n2 + p * n + s];
        v -= e;
    }
    dm2[(((size_t)p * n + q) * n + r) * n + s] = v;
}
