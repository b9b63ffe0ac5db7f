// K2, alpha-beta double excitations in the sign-adjusted string basis ("sigma-3" of determinant CI).
//
// The opposite-spin two-body part of a number-conserving Hamiltonian,
//     H_ab = sum_{i != j, k != l} W[ij][kl] a^+_{i alpha} a_{j alpha} a^+_{k beta} a_{l beta},
// is 1 296 of the ~1 400 multiply-adds per determinant at 24 qubits (n^2 (n-1)^2 / ... grows as n^4).
// In the Jordan-Wigner (interleaved alpha0 beta0 alpha1 beta1 ...) determinant basis its matrix
// elements carry a sign that couples the two strings,
//     popc(a & betw(i,j)) + popc(b & betw(k,l)) + popc(a & cross_a(k,l)) + popc(b' & cross_b(i,j)),
// which costs 4-5 instructions per multiply-add in k_sigma.  With the diagonal basis change
//     C~[a][b] = (-1)^{theta(a,b)} C[a][b],   theta(a,b) = popc(a & G(b)),
//     G(b) bit p = parity of the beta orbitals q < p occupied in b
// (the permutation sign between interleaved and alpha-major operator order) the two cross terms
// cancel exactly and the sign factorises:  <a'b'|H_ab|ab>~ = s_a(a; i,j) s_b(b; k,l) W[ij][kl].
// The kernel stages LB source rows (already multiplied by theta) and the LB x n(n-1) link weights
// s_a W[ij][.] in shared memory; a thread owns output columns and walks their N_b (n - N_b) incoming
// beta excitations - no padding, two shared-memory loads and two FMAs per multiply-add, one sign
// per ELL entry.
//
// W is not handed over by the caller (the C ABI receives Pauli terms): it is read off the term table,
// one X-mask group (two alpha + two beta flipped qubits) at a time, by evaluating the group on the
// two-electron determinant {j_alpha, l_beta} for each of its four occupation patterns, after an EXACT
// structural check: every term of the group must carry, outside the flipped qubits, precisely the
// Jordan-Wigner string of the two excitations (qubits strictly between the two alpha qubits XOR qubits
// strictly between the two beta qubits) - then the group acts on the sector as W[pattern] times the
// fermionic sign for every spectator occupation.  A group that fails (an operator that is not a
// number-conserving two-body operator) is left to the generic kernels.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

#include "mqc_common.h"

#define MQC_AB_THREADS 256
#define MQC_AB_R 4             // output beta strings per thread
#define MQC_AB_LB 4            // alpha links per batch
#define MQC_AB_WSTRIDE 512     // weights per link in shared memory (n (n - 1) <= 306 for n <= 18)

struct MqcSigmaAB {            // host tables
    bool ok = false;
    int n_orb = 0, n_pair = 0;             // n (n - 1) ordered orbital pairs (i, j), i != j: index i * (n - 1) + (j - (j > i))
    unsigned n_a = 0, n_b = 0;
    int e_a = 0, e_b = 0;                   // links per row / column: N_a (n - N_a), N_b (n - N_b)
    std::vector<double> W;                  // [n_pair][n_pair]
    std::vector<unsigned> lnk_a;            // [e_a][n_a]  source row (16) | pair index (9) << 16 | sign << 31
    std::vector<unsigned> ell_b;            // [e_b][n_b]  source column (16) | pair index (9) << 16 | sign << 31
    std::vector<unsigned> gB;               // [n_b] G(b)
    std::vector<size_t> used_terms;         // indices of the Pauli terms this part covers
};

struct MqcSigmaABDev {
    const double* W;
    const unsigned* lnk_a;
    const unsigned* ell_b;
    const unsigned* gB;
    const unsigned* aorb;                   // [n_a] alpha strings (orbital space)
    unsigned n_a, n_b;
    int e_a, e_b, n_pair;
};

static inline int mqc_ab_pair(int i, int j, int n) { return i * (n - 1) + (j - (j > i ? 1 : 0)); }
static inline unsigned mqc_ab_between(int p, int q) {          // orbitals strictly between
    const int lo = p < q ? p : q, hi = p < q ? q : p;
    return ((1u << hi) - 1u) & ~((2u << lo) - 1u);
}

// xq / zq: qubit-space masks (bit q <-> qubit q = 2 p + spin), <x ^ xq| P |x> = (-1)^popc(x & zq)
static inline MqcSigmaAB mqc_sigma_ab_build(int n_orb, int na, int nb, const std::vector<u64>& xq, const std::vector<u64>& zq,
                                            const std::vector<double2>& coef, const std::vector<unsigned>& astr,
                                            const std::vector<unsigned>& bstr) {
    MqcSigmaAB T;
    const int n = n_orb;
    if (n < 2 || n > 18 || astr.size() > 65535 || bstr.size() > 65535) return T;
    T.n_orb = n; T.n_pair = n * (n - 1);
    T.n_a = (unsigned)astr.size(); T.n_b = (unsigned)bstr.size();
    T.e_a = na * (n - na); T.e_b = nb * (n - nb);
    if (T.e_a == 0 || T.e_b == 0 || T.n_pair > MQC_AB_WSTRIDE) return T;
    u64 amask = 0, bmask = 0;
    for (int p = 0; p < n; ++p) { amask |= 1ull << (2 * p); bmask |= 1ull << (2 * p + 1); }
    std::map<u64, std::vector<size_t>> groups;
    for (size_t t = 0; t < xq.size(); ++t)
        if (__builtin_popcountll(xq[t] & amask) == 2 && __builtin_popcountll(xq[t] & bmask) == 2) groups[xq[t]].push_back(t);
    if (groups.empty()) return T;
    T.W.assign((size_t)T.n_pair * T.n_pair, 0.0);
    auto element = [&](const std::vector<size_t>& g, u64 x, double& re, double& im) {
        re = im = 0.0;
        for (size_t t : g) { const double s = (__builtin_popcountll(x & zq[t]) & 1) ? -1.0 : 1.0; re += s * coef[t].x; im += s * coef[t].y; }
    };
    // sign of a^+_{i a} a_{j a} a^+_{k b} a_{l b} on the determinant (a, b) in JW order (beta operators act first)
    auto jw_sign = [&](unsigned a, unsigned b, int i, int j, int k, int l) -> int {
        const int lo_kl = std::min(k, l), hi_kl = std::max(k, l), lo_ij = std::min(i, j), hi_ij = std::max(i, j);
        const unsigned cross_a = ((2u << hi_kl) - 1u) & ~((2u << lo_kl) - 1u);        // alpha orbitals lo < p <= hi
        const unsigned cross_b = ((1u << hi_ij) - 1u) & ~((1u << lo_ij) - 1u);        // beta orbitals lo <= q < hi
        const unsigned b2 = (b & ~(1u << l)) | (1u << k);
        return (__builtin_popcount(b & mqc_ab_between(k, l)) + __builtin_popcount(a & cross_a) +
                __builtin_popcount(a & mqc_ab_between(i, j)) + __builtin_popcount(b2 & cross_b)) & 1;
    };
    auto spread = [&](unsigned a, unsigned b) { u64 x = 0; for (int p = 0; p < n; ++p) { if ((a >> p) & 1u) x |= 1ull << (2 * p); if ((b >> p) & 1u) x |= 1ull << (2 * p + 1); } return x; };
    double wmax = 0.0;
    for (auto& kv : groups) {
        const u64 xm = kv.first;
        int ao[2], bo[2], ca = 0, cb = 0;
        for (int p = 0; p < n; ++p) { if ((xm >> (2 * p)) & 1ull) ao[ca++] = p; if ((xm >> (2 * p + 1)) & 1ull) bo[cb++] = p; }
        bool good = true;
        double w4[4] = {0, 0, 0, 0};
        {
            auto betw_q = [](int p, int q) { const int lo = p < q ? p : q, hi = p < q ? q : p; return ((1ull << hi) - 1ull) & ~((2ull << lo) - 1ull); };
            const u64 zexp = (betw_q(2 * ao[0], 2 * ao[1]) ^ betw_q(2 * bo[0] + 1, 2 * bo[1] + 1)) & ~xm;
            for (size_t t : kv.second) if ((zq[t] & ~xm) != zexp) { good = false; break; }
        }
        for (int c = 0; c < 4 && good; ++c) {
            const int i = ao[c & 1], j = ao[1 - (c & 1)], k = bo[(c >> 1) & 1], l = bo[1 - ((c >> 1) & 1)];
            double re, im;
            element(kv.second, spread(1u << j, 1u << l), re, im);
            if (std::fabs(im) > 1e-14 * std::max(1.0, std::fabs(re))) { good = false; break; }
            const double w = jw_sign(1u << j, 1u << l, i, j, k, l) ? -re : re;
            w4[c] = w;
        }
        if (!good) continue;
        for (int c = 0; c < 4; ++c) {
            const int i = ao[c & 1], j = ao[1 - (c & 1)], k = bo[(c >> 1) & 1], l = bo[1 - ((c >> 1) & 1)];
            T.W[(size_t)mqc_ab_pair(i, j, n) * T.n_pair + mqc_ab_pair(k, l, n)] = w4[c];
            wmax = std::max(wmax, std::fabs(w4[c]));
        }
        for (size_t t : kv.second) T.used_terms.push_back(t);
    }
    if (T.used_terms.empty()) return T;
    std::sort(T.used_terms.begin(), T.used_terms.end());
    // rank tables
    std::vector<int> rank_a((size_t)1 << n, -1), rank_b((size_t)1 << n, -1);
    for (size_t r = 0; r < astr.size(); ++r) rank_a[astr[r]] = (int)r;
    for (size_t r = 0; r < bstr.size(); ++r) rank_b[bstr[r]] = (int)r;
    // alpha links INTO row a': a' has i, lacks j; source a = a' - i + j; s_a = popc(a & betw(i,j))
    T.lnk_a.assign((size_t)T.e_a * T.n_a, 0u);
    for (unsigned r = 0; r < T.n_a; ++r) {
        const unsigned ap = astr[r];
        int e = 0;
        for (int i = 0; i < n; ++i) if ((ap >> i) & 1u) for (int j = 0; j < n; ++j) if (!((ap >> j) & 1u)) {
            const unsigned a = (ap & ~(1u << i)) | (1u << j);
            const unsigned sg = (unsigned)__builtin_popcount(a & mqc_ab_between(i, j)) & 1u;
            T.lnk_a[(size_t)e * T.n_a + r] = (unsigned)rank_a[a] | ((unsigned)mqc_ab_pair(i, j, n) << 16) | (sg << 31);
            ++e;
        }
    }
    T.ell_b.assign((size_t)T.e_b * T.n_b, 0u);
    T.gB.assign(T.n_b, 0u);
    for (unsigned c = 0; c < T.n_b; ++c) {
        const unsigned bp = bstr[c];
        unsigned g = 0;
        for (int p = 0; p < n; ++p) if (__builtin_popcount(bp & ((1u << p) - 1u)) & 1) g |= 1u << p;
        T.gB[c] = g;
        int e = 0;
        for (int k = 0; k < n; ++k) if ((bp >> k) & 1u) for (int l = 0; l < n; ++l) if (!((bp >> l) & 1u)) {
            const unsigned b = (bp & ~(1u << k)) | (1u << l);
            const unsigned sg = (unsigned)__builtin_popcount(b & mqc_ab_between(k, l)) & 1u;
            T.ell_b[(size_t)e * T.n_b + c] = (unsigned)rank_b[b] | ((unsigned)mqc_ab_pair(k, l, n) << 16) | (sg << 31);
            ++e;
        }
    }
    T.ok = true;
    return T;
}

// Host evaluation (tests): L += H_ab C, canonical order
static inline void mqc_sigma_ab_apply_host(const MqcSigmaAB& T, const std::vector<unsigned>& astr, const double2* C, double2* L) {
    const unsigned nA = T.n_a, nB = T.n_b;
    for (unsigned r = 0; r < nA; ++r) for (unsigned c = 0; c < nB; ++c) {
        double ar = 0.0, ai = 0.0;
        for (int ea = 0; ea < T.e_a; ++ea) {
            const unsigned la = T.lnk_a[(size_t)ea * nA + r];
            const unsigned src = la & 0xffffu, ij = (la >> 16) & 0x1ffu, sa = la >> 31;
            for (int eb = 0; eb < T.e_b; ++eb) {
                const unsigned lb = T.ell_b[(size_t)eb * nB + c];
                const unsigned col = lb & 0xffffu, kl = (lb >> 16) & 0x1ffu, sb = lb >> 31;
                const unsigned th = (unsigned)__builtin_popcount(astr[src] & T.gB[col]) & 1u;
                double w = T.W[(size_t)ij * T.n_pair + kl];
                if ((sa ^ sb ^ th) & 1u) w = -w;
                ar += w * C[(size_t)src * nB + col].x; ai += w * C[(size_t)src * nB + col].y;
            }
        }
        if (__builtin_popcount(astr[r] & T.gB[c]) & 1) { ar = -ar; ai = -ai; }
        L[(size_t)r * nB + c].x += ar; L[(size_t)r * nB + c].y += ai;
    }
}

#if defined(__CUDACC__)
__device__ __forceinline__ double mqc_ab_flip(double w, unsigned s) {      // s odd -> -w
    return __hiloint2double(__double2hiint(w) ^ (int)(s << 31), __double2loint(w));
}

// One CTA per (output alpha string, block of THREADS * R output beta strings).  L += H_ab C and the energy
// partial Re / Im <C|H_ab|C> into e_out[0..1].  STAGE: source rows in shared memory.
template <bool STAGE>
__global__ void __launch_bounds__(MQC_AB_THREADS, 3)
k_sigma_ab(const double2* __restrict__ C, double2* __restrict__ L, const __grid_constant__ MqcSigmaABDev S, double* __restrict__ e_out) {
    extern __shared__ __align__(16) unsigned char ab_smem[];
    const unsigned nA = S.n_a, nB = S.n_b;
    double* wl = reinterpret_cast<double*>(ab_smem);                                        // [LB][WSTRIDE]
    double2* srow = reinterpret_cast<double2*>(ab_smem + (size_t)MQC_AB_LB * MQC_AB_WSTRIDE * sizeof(double));   // [LB][nB]
    __shared__ double red[2][MQC_AB_THREADS / 32];
    const int tid = threadIdx.x;
    const unsigned i = blockIdx.x;
    const unsigned ap = __ldg(S.aorb + i);
    unsigned jb[MQC_AB_R], jc[MQC_AB_R];
    double ar[MQC_AB_R], ai[MQC_AB_R];
#pragma unroll
    for (int r = 0; r < MQC_AB_R; ++r) {
        jb[r] = (blockIdx.y * MQC_AB_R + r) * MQC_AB_THREADS + tid;
        jc[r] = jb[r] < nB ? jb[r] : nB - 1u;
        ar[r] = 0.0; ai[r] = 0.0;
    }
    for (int l0 = 0; l0 < S.e_a; l0 += MQC_AB_LB) {
        unsigned src[MQC_AB_LB], ij[MQC_AB_LB], sa[MQC_AB_LB];
        bool ok[MQC_AB_LB];
#pragma unroll
        for (int l = 0; l < MQC_AB_LB; ++l) {
            ok[l] = l0 + l < S.e_a;
            const unsigned ln = ok[l] ? __ldg(S.lnk_a + (size_t)(l0 + l) * nA + i) : 0u;
            src[l] = ln & 0xffffu; ij[l] = (ln >> 16) & 0x1ffu; sa[l] = ln >> 31;
        }
        __syncthreads();                                          // previous batch is done with wl / srow
#pragma unroll
        for (int l = 0; l < MQC_AB_LB; ++l) {
            for (int k = tid; k < S.n_pair; k += MQC_AB_THREADS)
                wl[l * MQC_AB_WSTRIDE + k] = ok[l] ? mqc_ab_flip(__ldg(S.W + (size_t)ij[l] * S.n_pair + k), sa[l]) : 0.0;
            if (STAGE) {
                const unsigned a = __ldg(S.aorb + src[l]);
                const double2* __restrict__ Crow = C + (size_t)src[l] * nB;
                for (unsigned j = tid; j < nB; j += MQC_AB_THREADS) {
                    const double2 v = Crow[j];
                    const unsigned th = (unsigned)__popc(a & __ldg(S.gB + j)) & 1u;
                    srow[(size_t)l * nB + j] = make_double2(mqc_ab_flip(v.x, th), mqc_ab_flip(v.y, th));
                }
            }
        }
        __syncthreads();
        unsigned asrc[MQC_AB_LB];
        if (!STAGE) {
#pragma unroll
            for (int l = 0; l < MQC_AB_LB; ++l) asrc[l] = __ldg(S.aorb + src[l]);
        }
        unsigned nx[MQC_AB_R];
#pragma unroll
        for (int r = 0; r < MQC_AB_R; ++r) nx[r] = __ldg(S.ell_b + jc[r]);
        for (int e = 0; e < S.e_b; ++e) {
            unsigned en[MQC_AB_R];
#pragma unroll
            for (int r = 0; r < MQC_AB_R; ++r) { en[r] = nx[r]; if (e + 1 < S.e_b) nx[r] = __ldg(S.ell_b + (size_t)(e + 1) * nB + jc[r]); }
#pragma unroll
            for (int r = 0; r < MQC_AB_R; ++r) {
                const unsigned col = en[r] & 0xffffu, kl = (en[r] >> 16) & 0x1ffu;
                double tr = 0.0, ti = 0.0;
                if (STAGE) {
#pragma unroll
                    for (int l = 0; l < MQC_AB_LB; ++l) {
                        const double w = wl[l * MQC_AB_WSTRIDE + kl];
                        const double2 v = srow[(size_t)l * nB + col];
                        tr = fma(w, v.x, tr); ti = fma(w, v.y, ti);
                    }
                } else {
                    const unsigned g = __ldg(S.gB + col);
#pragma unroll
                    for (int l = 0; l < MQC_AB_LB; ++l) {
                        const double w = mqc_ab_flip(wl[l * MQC_AB_WSTRIDE + kl], (unsigned)__popc(asrc[l] & g));
                        const double2 v = C[(size_t)src[l] * nB + col];
                        tr = fma(w, v.x, tr); ti = fma(w, v.y, ti);
                    }
                }
                const int sgn = (int)(en[r] & 0x80000000u);
                ar[r] += __hiloint2double(__double2hiint(tr) ^ sgn, __double2loint(tr));
                ai[r] += __hiloint2double(__double2hiint(ti) ^ sgn, __double2loint(ti));
            }
        }
    }
    double er = 0.0, ei = 0.0;
#pragma unroll
    for (int r = 0; r < MQC_AB_R; ++r) {
        if (jb[r] < nB) {
            const unsigned th = (unsigned)__popc(ap & __ldg(S.gB + jb[r])) & 1u;
            const double xr = mqc_ab_flip(ar[r], th), xi = mqc_ab_flip(ai[r], th);
            const size_t o = (size_t)i * nB + jb[r];
            if (L) { const double2 o0 = L[o]; L[o] = make_double2(o0.x + xr, o0.y + xi); }
            const double2 py = C[o];
            er += py.x * xr + py.y * xi;
            ei += py.x * xi - py.y * xr;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { er += __shfl_xor_sync(0xffffffffu, er, o); ei += __shfl_xor_sync(0xffffffffu, ei, o); }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { red[0][warp] = er; red[1][warp] = ei; }
    __syncthreads();
    if (warp == 0) {
        double vr = (lane < MQC_AB_THREADS / 32) ? red[0][lane] : 0.0;
        double vi = (lane < MQC_AB_THREADS / 32) ? red[1][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { vr += __shfl_xor_sync(0xffffffffu, vr, o); vi += __shfl_xor_sync(0xffffffffu, vi, o); }
        if (lane == 0) {
            if (vr != 0.0) atomicAdd(e_out, vr);
            if (vi != 0.0) atomicAdd(e_out + 1, vi);
        }
    }
}

static inline size_t mqc_sigma_ab_smem_bytes(unsigned n_b, bool stage) {
    return (size_t)MQC_AB_LB * MQC_AB_WSTRIDE * sizeof(double) + (stage ? (size_t)MQC_AB_LB * n_b * sizeof(double2) : 0);
}
#endif
