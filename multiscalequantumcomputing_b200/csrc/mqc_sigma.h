// K2, sector form, second generation: lambda = H psi on the (N_alpha, N_beta) sector matrix
// C[ia][ib] with precomputed excitation link tables.
//
// The qubit Hamiltonian arrives as Pauli terms grouped by X-flip mask (the north star's
// grouping).  An X-mask group (ma, mb) [flipped alpha / beta orbitals] acting on determinant
// x = (a, b) contributes  coefficient(x) * C[a][b]  to the output (a ^ ma, b ^ mb) with
//     coefficient(x) = (-1)^popc(x & zc) * W[pattern of x on the flipped bits]
// whenever all its terms share the parity mask zc outside the flipped bits ("simple" group),
// and for a Jordan-Wigner image of a fermionic operator zc is the JW string
//     zth(ma, mb) = XOR_{f in flipped qubits} (mask of qubits below f)   & ~flipped.
// zth is an XOR of an "alpha excitation" part and a "beta excitation" part, so the sign
// factorises into   [source alpha string, ma] x [output beta string, ma] x [entry: beta
// string, mb] x [small per-(ma, mb, pattern) corrections folded into the weights].  That is
// what lets the matrix-vector product run as
//     for every alpha link (ia' <- ia, ma):   acc[ib'] += s_a * f1(ib') * sum_e  s_e * Wl[e.w] * C[ia][e.col]
// with ONE beta-side ELL table (per class of ma with the same set of non-zero mb) that is
// independent of the link, a 1-3 KB per-link weight table in shared memory and the source
// row staged in shared memory: ~10 instructions per matrix element instead of ~50.
//
//   (|ma|,|mb|) = (2,2)  alpha-beta doubles          phase 1, ELL entries
//                 (2,0)  alpha singles (n_r-dependent) phase 1, diagonal-in-beta term:
//                        coefficient = sign * (A(a) + B_pattern(b)) (checked: every term's
//                        extra Z's sit on one spin species)
//                 (4,0)  alpha-alpha doubles          phase 2, row links, diagonal in beta
//                 (0,2)  beta singles, (0,4) beta-beta doubles, (0,0) diagonal: phase 0
// Groups that do not fit (non-JW parity masks, > 4 flipped bits, complex weights) stay on the
// first-generation kernel k_happly_sector, which then ADDS its part: any Pauli Hamiltonian is
// still handled exactly.
//
// Everything here lives in ORBITAL space (strings as n_orb-bit masks, qubit q = 2p + spin) and
// is independent of the physical layout of the state vector.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <map>
#include <vector>

#include "mqc_common.h"

#define MQC_SG_INVALID 0xffffffffu
#define MQC_SG_THREADS 256
#define MQC_SG_LB 4            // alpha links per batch (rows staged together, one pass over the ELL table)

struct MqcSigmaDev {
    unsigned n_a, n_b;             // strings
    int n_m2, n_m4;
    // phase 1 (alpha single links) and phase 0 (beta-only)
    const unsigned* lnk_a1;        // [e1a][n_a]  src(16) | maidx(8) | pav(1) | sign(1)   (INVALID: none)
    int e1a;
    const unsigned* lnk_a2;        // [e2a][n_a]  src(16) | ma4idx(12) | pa6(3) | sign(1)
    int e2a;
    const unsigned* ell_b1;        // [sum_c e1b_c][n_b]  col(16) | mbidx(8) | pbv(1) | sign(1)
    const int* cls_off;            // [n_cls + 1] first ELL row of class c
    const int* cls_of_m2;          // [n_m2 + 1] class of alpha mask index (entry n_m2: ma = 0); -1 none
    const unsigned* ell_b2;        // [e2b][n_b]  col(16) | mb4idx(12) | pb6(3) | sign(1)
    int e2b;
    const double* w22;             // [n_m2][n_m2][4]
    const double* w4a;             // [n_m4][6]
    const double* w4b;             // [n_m4][6]
    const double* ta1;             // [n_m2][n_a]      A(a), alpha singles, by SOURCE string
    const double* tb1;             // [n_m2][2][n_b]   B_pav(b)
    const double* tab;             // [n_m2][2][n_a]   A_pbv(a'), beta singles
    const double* tbb;             // [n_m2][n_b]      B(b), by SOURCE string
    const unsigned* m2_mask;       // [n_m2] orbital mask
    const unsigned* sa_b2;         // [n_m2] alpha orbitals in the JW string of a BETA 2-flip
    const unsigned* sb_a2;         // [n_m2] beta orbitals in the JW string of an ALPHA 2-flip
    const unsigned* sa_b4;         // [n_m4]
    const unsigned* sb_a4;         // [n_m4]
    const unsigned* aorb;          // [n_a]
    const unsigned* borb;          // [n_b]
    const double* hdiag;           // [rows of this rank][n_b]
    unsigned row0;                 // first row of hdiag / of this launch
    // Row-distributed sector matrix: the launch applies only the links whose SOURCE row lies in
    // [src_lo, src_hi) (the block `Csrc` holds, first row src_lo); own rows come from `Cown`
    // (first row own_lo).  Replicated matrix: one block [0, n_a), all three offsets 0.
    unsigned src_lo, src_hi, own_lo, l_lo;
    int accumulate;                // L += instead of L =  (rounds after the first)
    unsigned pitch;                // elements between consecutive rows of Cown / Csrc / L (0: n_b)
    // reproducible energy: every CTA leaves its partial in e_part[2 * cta], the last CTA to finish (ticket) adds them
    // in a fixed order into e_out; NULL: one atomicAdd per CTA
    double* e_part;
    unsigned* ticket;
    // Row split (replicated matrix only): gridDim.z == 2, part z = 0 does phase 0 and the alpha single
    // links [0, split_l), part 1 the remaining links and phase 2; both ADD into a zeroed L with
    // atomics (two addends per element: the sum does not depend on their order).  0: one CTA per row.
    int split_l;
};

// ---- host side -----------------------------------------------------------------------------
struct MqcSigmaHost {
    int n_orb = 0, na = 0, nb = 0;
    unsigned n_a = 0, n_b = 0;
    int n_m2 = 0, n_m4 = 0, e1a = 0, e2a = 0, e2b = 0, n_cls = 0;
    std::vector<unsigned> lnk_a1, lnk_a2, ell_b1, ell_b2, m2_mask, m4_mask, sa_b2, sb_a2, sa_b4, sb_a4, aorb, borb;
    std::vector<int> cls_off, cls_of_m2;
    std::vector<double> w22, w4a, w4b, ta1, tb1, tab, tbb;
    // terms the fast path does not cover (logical masks / coefficients as given by the user)
    std::vector<size_t> fallback_terms;
    // diagonal group terms in qubit space, evaluated once per Hamiltonian on the device
    std::vector<u64> diag_z;
    std::vector<double> diag_c;
    std::vector<u64> aq, bq;       // qubit-space images of the strings
    size_t n_fast_terms = 0;
};

static inline u64 mqc_sg_below_xor(u64 mq) {       // XOR over flipped qubits f of (qubits below f)
    u64 z = 0, m = mq;
    while (m) { const int f = __builtin_ctzll(m); m &= m - 1; z ^= (1ull << f) - 1ull; }
    return z;
}
static inline unsigned mqc_sg_spin_part(u64 zq, int spin, int n_orb) {   // qubit mask -> orbital mask of one spin
    unsigned o = 0;
    for (int p = 0; p < n_orb; ++p) if ((zq >> (2 * p + spin)) & 1ull) o |= 1u << p;
    return o;
}
static inline u64 mqc_sg_spread(unsigned orb, int spin, int n_orb) {
    u64 q = 0;
    for (int p = 0; p < n_orb; ++p) if ((orb >> p) & 1u) q |= 1ull << (2 * p + spin);
    return q;
}
static inline unsigned mqc_sg_pext(unsigned x, unsigned m) {
    unsigned out = 0, k = 0;
    while (m) { const unsigned b = __builtin_ctz(m); m &= m - 1; out |= ((x >> b) & 1u) << k; ++k; }
    return out;
}
static inline int mqc_sg_pat6(unsigned p4) {        // 4-bit pattern with two bits set -> 0..5
    switch (p4) { case 3: return 0; case 5: return 1; case 6: return 2; case 9: return 3; case 10: return 4; case 12: return 5; }
    return -1;
}

static inline void mqc_sg_strings(int n_orb, int k, std::vector<unsigned>& out, std::vector<int>& rank) {
    out.clear();
    rank.assign((size_t)1 << n_orb, -1);
    if (k < 0 || k > n_orb) return;
    if (k == 0) { out.push_back(0); rank[0] = 0; return; }
    u64 v = (1ull << k) - 1ull;
    const u64 limit = 1ull << n_orb;
    while (v < limit) {
        rank[v] = (int)out.size();
        out.push_back((unsigned)v);
        const u64 c = v & (~v + 1ull), r = v + c;
        v = (((r ^ v) >> 2) / c) | r;
    }
}

// The order of a column's ELL entries is free (a sum).  Sixteen consecutive output columns are the lanes of one
// shared-memory wavefront of 8-byte gathers: permute every column's entries so that, slot by slot, the sixteen
// source columns fall into different bank pairs (col mod 16) and - for the alpha-beta table - so do the weight
// indices (greedy; equal addresses are a broadcast, not a conflict).  key2 < 0: no second key.
template <class Key2>
static inline void mqc_sg_bank_order(std::vector<std::vector<unsigned>>& rows, Key2 key2) {
    const size_t nb = rows.size();
    for (size_t g0 = 0; g0 < nb; g0 += 16) {
        const size_t g1 = std::min(nb, g0 + 16);
        size_t emax = 0;
        for (size_t j = g0; j < g1; ++j) emax = std::max(emax, rows[j].size());
        std::vector<std::vector<unsigned>> out(g1 - g0);
        std::vector<std::vector<char>> used(g1 - g0);
        for (size_t j = g0; j < g1; ++j) { used[j - g0].assign(rows[j].size(), 0); out[j - g0].reserve(rows[j].size()); }
        for (size_t e = 0; e < emax; ++e) {
            int bank_col[16], bank_k2[16];
            for (int b = 0; b < 16; ++b) { bank_col[b] = -1; bank_k2[b] = -1; }
            for (size_t j = g0; j < g1; ++j) {
                const std::vector<unsigned>& r = rows[j];
                std::vector<char>& u = used[j - g0];
                if (out[j - g0].size() >= r.size()) continue;
                int best = -1, best_cost = 99;
                for (size_t c = 0; c < r.size(); ++c) {
                    if (u[c]) continue;
                    const int col = (int)(r[c] & 0xffffu), k2 = key2(r[c]);
                    int cost = 0;
                    if (bank_col[col & 15] >= 0 && bank_col[col & 15] != col) cost += 2;
                    if (k2 >= 0 && bank_k2[k2 & 15] >= 0 && bank_k2[k2 & 15] != k2) cost += 1;
                    if (cost < best_cost) { best_cost = cost; best = (int)c; if (!cost) break; }
                }
                u[best] = 1;
                const int col = (int)(r[best] & 0xffffu), k2 = key2(r[best]);
                if (bank_col[col & 15] < 0) bank_col[col & 15] = col;
                if (k2 >= 0 && bank_k2[k2 & 15] < 0) bank_k2[k2 & 15] = k2;
                out[j - g0].push_back(r[best]);
            }
        }
        for (size_t j = g0; j < g1; ++j) rows[j].swap(out[j - g0]);
    }
}

// xq / zq: Pauli masks in QUBIT space (bit q = qubit q), c: coefficients.  Returns the tables;
// terms that cannot take the fast path are listed in fallback_terms.
static inline MqcSigmaHost mqc_sigma_build(int n_orb, int na, int nb, const std::vector<u64>& xq,
                                           const std::vector<u64>& zq, const std::vector<double2>& cf) {
    MqcSigmaHost S;
    S.n_orb = n_orb; S.na = na; S.nb = nb;
    std::vector<int> rank_a, rank_b;
    mqc_sg_strings(n_orb, na, S.aorb, rank_a);
    mqc_sg_strings(n_orb, nb, S.borb, rank_b);
    S.n_a = (unsigned)S.aorb.size(); S.n_b = (unsigned)S.borb.size();
    for (unsigned a : S.aorb) S.aq.push_back(mqc_sg_spread(a, 0, n_orb));
    for (unsigned b : S.borb) S.bq.push_back(mqc_sg_spread(b, 1, n_orb));
    // mask index tables
    std::vector<int> idx2((size_t)1 << n_orb, -1), idx4((size_t)1 << n_orb, -1);
    for (unsigned m = 0; m < (1u << n_orb); ++m) {
        const int pc = __builtin_popcount(m);
        if (pc == 2) { idx2[m] = (int)S.m2_mask.size(); S.m2_mask.push_back(m); }
        if (pc == 4) { idx4[m] = (int)S.m4_mask.size(); S.m4_mask.push_back(m); }
    }
    S.n_m2 = (int)S.m2_mask.size(); S.n_m4 = (int)S.m4_mask.size();
    const int M2 = std::max(1, S.n_m2), M4 = std::max(1, S.n_m4);
    S.sa_b2.assign(M2, 0); S.sb_a2.assign(M2, 0); S.sa_b4.assign(M4, 0); S.sb_a4.assign(M4, 0);
    std::vector<unsigned> sa_a2(M2, 0), sb_b2(M2, 0), sa_a4(M4, 0), sb_b4(M4, 0);
    for (int i = 0; i < S.n_m2; ++i) {
        const u64 za = mqc_sg_below_xor(mqc_sg_spread(S.m2_mask[i], 0, n_orb)), zb = mqc_sg_below_xor(mqc_sg_spread(S.m2_mask[i], 1, n_orb));
        sa_a2[i] = mqc_sg_spin_part(za, 0, n_orb); S.sb_a2[i] = mqc_sg_spin_part(za, 1, n_orb);
        S.sa_b2[i] = mqc_sg_spin_part(zb, 0, n_orb); sb_b2[i] = mqc_sg_spin_part(zb, 1, n_orb);
    }
    for (int i = 0; i < S.n_m4; ++i) {
        const u64 za = mqc_sg_below_xor(mqc_sg_spread(S.m4_mask[i], 0, n_orb)), zb = mqc_sg_below_xor(mqc_sg_spread(S.m4_mask[i], 1, n_orb));
        sa_a4[i] = mqc_sg_spin_part(za, 0, n_orb); S.sb_a4[i] = mqc_sg_spin_part(za, 1, n_orb);
        S.sa_b4[i] = mqc_sg_spin_part(zb, 0, n_orb); sb_b4[i] = mqc_sg_spin_part(zb, 1, n_orb);
    }
    S.w22.assign((size_t)M2 * M2 * 4, 0.0);
    // one extra all-zero mask index (n_m2 / n_m4) is what the padding entries of the beta ELL
    // tables point to: the hot loops need no validity test
    S.w4a.assign((size_t)M4 * 6, 0.0); S.w4b.assign((size_t)(M4 + 1) * 6, 0.0);
    S.ta1.assign((size_t)M2 * S.n_a, 0.0); S.tb1.assign((size_t)M2 * 2 * S.n_b, 0.0);
    S.tab.assign((size_t)M2 * 2 * S.n_a, 0.0); S.tbb.assign((size_t)(M2 + 1) * S.n_b, 0.0);
    S.sa_b4.push_back(0);

    // ---- group the terms by (ma, mb) ---------------------------------------------------------
    const u64 amask_q = mqc_sg_spread((1u << n_orb) - 1u, 0, n_orb), bmask_q = amask_q << 1;
    std::map<std::pair<unsigned, unsigned>, std::vector<size_t>> groups;
    for (size_t t = 0; t < xq.size(); ++t) {
        const unsigned ma = mqc_sg_spin_part(xq[t], 0, n_orb), mb = mqc_sg_spin_part(xq[t], 1, n_orb);
        if ((__builtin_popcount(ma) | __builtin_popcount(mb)) & 1) continue;      // leaves the sector: never contributes
        groups[{ma, mb}].push_back(t);
    }
    std::vector<char> has_a1(M2, 0), has_b1(M2, 0);       // single-excitation groups on the fast path
    std::vector<std::vector<char>> nz22(M2, std::vector<char>(M2, 0));
    for (auto& kv : groups) {
        const unsigned ma = kv.first.first, mb = kv.first.second;
        const std::vector<size_t>& tt = kv.second;
        const int pa = __builtin_popcount(ma), pb = __builtin_popcount(mb);
        const u64 mq = mqc_sg_spread(ma, 0, n_orb) | mqc_sg_spread(mb, 1, n_orb);
        const u64 zth = mqc_sg_below_xor(mq) & ~mq;
        bool real = true;
        for (size_t t : tt) if (cf[t].y != 0.0) real = false;
        bool ok = real;
        if (ok && pa == 0 && pb == 0) {
            for (size_t t : tt) { S.diag_z.push_back(zq[t]); S.diag_c.push_back(cf[t].x); }
            S.n_fast_terms += tt.size();
            continue;
        }
        const bool dbl = (pa == 2 && pb == 2) || (pa == 4 && pb == 0) || (pa == 0 && pb == 4);
        const bool sgl = (pa == 2 && pb == 0) || (pa == 0 && pb == 2);
        if (!dbl && !sgl) ok = false;
        if (ok && dbl) for (size_t t : tt) if ((zq[t] & ~mq) != zth) { ok = false; break; }
        if (ok && sgl) for (size_t t : tt) {
            const u64 e = (zq[t] & ~mq) ^ zth;
            if ((e & amask_q) && (e & bmask_q)) { ok = false; break; }
        }
        if (!ok) { for (size_t t : tt) S.fallback_terms.push_back(t); continue; }
        S.n_fast_terms += tt.size();
        // flipped qubits in pattern order: alpha orbitals ascending, then beta orbitals ascending
        u64 pbit[4]; int np = 0;
        for (int p = 0; p < n_orb; ++p) if ((ma >> p) & 1u) pbit[np++] = 1ull << (2 * p);
        for (int p = 0; p < n_orb; ++p) if ((mb >> p) & 1u) pbit[np++] = 1ull << (2 * p + 1);
        auto weight = [&](unsigned pat) {
            u64 xm = 0;
            for (int k = 0; k < np; ++k) if ((pat >> k) & 1u) xm |= pbit[k];
            double w = 0;
            for (size_t t : tt) w += (__builtin_popcountll(xm & zq[t]) & 1) ? -cf[t].x : cf[t].x;
            return w;
        };
        if (pa == 2 && pb == 2) {
            const int ia = idx2[ma], ib = idx2[mb];
            bool any = false;
            for (int pav = 0; pav < 2; ++pav) for (int pbv = 0; pbv < 2; ++pbv) {
                const double w = weight((unsigned)(pav + 1) | ((unsigned)(pbv + 1) << 2));
                S.w22[((size_t)ia * M2 + ib) * 4 + pav * 2 + pbv] = w;
                if (w != 0.0) any = true;
            }
            nz22[ia][ib] = any;
        } else if (pa == 4) {
            static const unsigned pats[6] = {3, 5, 6, 9, 10, 12};
            for (int k = 0; k < 6; ++k) S.w4a[(size_t)idx4[ma] * 6 + k] = weight(pats[k]);
        } else if (pb == 4) {
            static const unsigned pats[6] = {3, 5, 6, 9, 10, 12};
            for (int k = 0; k < 6; ++k) S.w4b[(size_t)idx4[mb] * 6 + k] = weight(pats[k]);
        } else if (pa == 2) {
            // coefficient(x) = (-1)^popc(x & zth) [ A(a) + B_pav(b) ]
            const int ia = idx2[ma];
            has_a1[ia] = 1;
            for (size_t t : tt) {
                const u64 e = (zq[t] & ~mq) ^ zth, zin = zq[t] & mq;
                if (e & bmask_q) {
                    for (int pav = 0; pav < 2; ++pav) {
                        const u64 xm = pbit[pav];                      // source pattern: that flipped qubit occupied
                        const double c0 = (__builtin_popcountll(xm & zin) & 1) ? -cf[t].x : cf[t].x;
                        double* dst = &S.tb1[((size_t)ia * 2 + pav) * S.n_b];
                        for (unsigned j = 0; j < S.n_b; ++j) dst[j] += (__builtin_popcountll(S.bq[j] & e) & 1) ? -c0 : c0;
                    }
                } else {
                    double* dst = &S.ta1[(size_t)ia * S.n_a];
                    for (unsigned i = 0; i < S.n_a; ++i) {
                        if (__builtin_popcount(S.aorb[i] & ma) != 1) continue;
                        dst[i] += (__builtin_popcountll(S.aq[i] & (e | zin)) & 1) ? -cf[t].x : cf[t].x;
                    }
                }
            }
        } else {
            // beta single: coefficient(x) = (-1)^popc(x & zth) [ A_pbv(a) + B(b) ]
            const int ib = idx2[mb];
            has_b1[ib] = 1;
            for (size_t t : tt) {
                const u64 e = (zq[t] & ~mq) ^ zth, zin = zq[t] & mq;
                if (e & amask_q) {
                    for (int pbv = 0; pbv < 2; ++pbv) {
                        const u64 xm = pbit[pbv];
                        const double c0 = (__builtin_popcountll(xm & zin) & 1) ? -cf[t].x : cf[t].x;
                        double* dst = &S.tab[((size_t)ib * 2 + pbv) * S.n_a];
                        for (unsigned i = 0; i < S.n_a; ++i) dst[i] += (__builtin_popcountll(S.aq[i] & e) & 1) ? -c0 : c0;
                    }
                } else {
                    double* dst = &S.tbb[(size_t)ib * S.n_b];
                    for (unsigned j = 0; j < S.n_b; ++j) {
                        if (__builtin_popcount(S.borb[j] & mb) != 1) continue;
                        dst[j] += (__builtin_popcountll(S.bq[j] & (e | zin)) & 1) ? -cf[t].x : cf[t].x;
                    }
                }
            }
        }
    }
    // ---- classes of alpha 2-masks by their set of non-zero beta partners -------------------
    // class key: the list of beta mask indices with a non-zero (2,2) weight; index n_m2 is the
    // pseudo mask "ma = 0" whose partners are the beta single groups.
    std::map<std::vector<int>, int> cls_id;
    std::vector<std::vector<int>> cls_members;
    S.cls_of_m2.assign(S.n_m2 + 1, -1);
    for (int ia = 0; ia <= S.n_m2; ++ia) {
        std::vector<int> key;
        for (int ib = 0; ib < S.n_m2; ++ib) if (ia < S.n_m2 ? nz22[ia][ib] : has_b1[ib]) key.push_back(ib);
        if (key.empty() && !(ia < S.n_m2 && has_a1[ia])) continue;
        auto it = cls_id.find(key);
        if (it == cls_id.end()) { it = cls_id.emplace(key, (int)cls_members.size()).first; cls_members.push_back(key); }
        S.cls_of_m2[ia] = it->second;
    }
    if ((int)cls_members.size() > 12) {
        // too fragmented to be worth separate tables: one class with every beta mask
        std::vector<int> all;
        for (int ib = 0; ib < S.n_m2; ++ib) all.push_back(ib);
        cls_members.assign(1, all);
        for (int ia = 0; ia <= S.n_m2; ++ia) if (S.cls_of_m2[ia] >= 0) S.cls_of_m2[ia] = 0;
    }
    S.n_cls = (int)cls_members.size();
    S.cls_off.assign(S.n_cls + 1, 0);
    // beta ELL per class: entries of output string jb' = sources jb = rank(b' ^ mb), b has exactly one bit of mb
    for (int c = 0; c < S.n_cls; ++c) {
        std::vector<std::vector<unsigned>> rows(S.n_b);
        size_t emax = 0;
        for (unsigned j = 0; j < S.n_b; ++j) {
            const unsigned bp = S.borb[j];
            for (int ib : cls_members[c]) {
                const unsigned mb = S.m2_mask[ib];
                if (__builtin_popcount(bp & mb) != 1) continue;
                const unsigned b = bp ^ mb;
                const unsigned pbv = mqc_sg_pext(b, mb) - 1u;
                const unsigned sgn = __builtin_popcount(b & sb_b2[ib] & ~mb) & 1u;
                rows[j].push_back((unsigned)rank_b[b] | ((unsigned)ib << 16) | (pbv << 24) | (sgn << 25));
            }
            emax = std::max(emax, rows[j].size());
        }
        mqc_sg_bank_order(rows, [](unsigned en) { return (int)(((en >> 16) & 0xffu) * 2u + ((en >> 24) & 1u)); });
        S.cls_off[c + 1] = S.cls_off[c] + (int)emax;
        const size_t base = S.ell_b1.size();
        S.ell_b1.resize(base + emax * S.n_b, (unsigned)S.n_m2 << 16);
        for (unsigned j = 0; j < S.n_b; ++j)
            for (size_t e = 0; e < rows[j].size(); ++e) S.ell_b1[base + e * S.n_b + j] = rows[j][e];
    }
    // beta-beta doubles
    {
        std::vector<std::vector<unsigned>> rows(S.n_b);
        size_t emax = 0;
        for (int i4 = 0; i4 < S.n_m4; ++i4) {
            bool any = false;
            for (int k = 0; k < 6; ++k) if (S.w4b[(size_t)i4 * 6 + k] != 0.0) any = true;
            if (!any) continue;
            const unsigned mb = S.m4_mask[i4];
            for (unsigned j = 0; j < S.n_b; ++j) {
                const unsigned bp = S.borb[j];
                if (__builtin_popcount(bp & mb) != 2) continue;
                const unsigned b = bp ^ mb;
                const int p6 = mqc_sg_pat6(mqc_sg_pext(b, mb));
                if (S.w4b[(size_t)i4 * 6 + p6] == 0.0) continue;
                const unsigned sgn = __builtin_popcount(b & sb_b4[i4] & ~mb) & 1u;
                rows[j].push_back((unsigned)rank_b[b] | ((unsigned)i4 << 16) | ((unsigned)p6 << 28) | (sgn << 31));
            }
        }
        for (auto& r : rows) emax = std::max(emax, r.size());
        mqc_sg_bank_order(rows, [](unsigned) { return -1; });
        S.e2b = (int)emax;
        S.ell_b2.assign(std::max<size_t>(1, emax * S.n_b), (unsigned)S.n_m4 << 16);
        for (unsigned j = 0; j < S.n_b; ++j)
            for (size_t e = 0; e < rows[j].size(); ++e) S.ell_b2[e * S.n_b + j] = rows[j][e];
    }
    // alpha single links (only masks that have any partner) and alpha-alpha double links
    {
        std::vector<std::vector<unsigned>> r1(S.n_a), r2(S.n_a);
        for (unsigned i = 0; i < S.n_a; ++i) {
            const unsigned ap = S.aorb[i];
            for (int ia = 0; ia < S.n_m2; ++ia) {
                if (S.cls_of_m2[ia] < 0) continue;
                const unsigned ma = S.m2_mask[ia];
                if (__builtin_popcount(ap & ma) != 1) continue;
                const unsigned a = ap ^ ma;
                const unsigned pav = mqc_sg_pext(a, ma) - 1u;
                const unsigned sgn = __builtin_popcount(a & sa_a2[ia] & ~ma) & 1u;
                r1[i].push_back((unsigned)rank_a[a] | ((unsigned)ia << 16) | (pav << 24) | (sgn << 25));
            }
            for (int i4 = 0; i4 < S.n_m4; ++i4) {
                const unsigned ma = S.m4_mask[i4];
                if (__builtin_popcount(ap & ma) != 2) continue;
                const unsigned a = ap ^ ma;
                const int p6 = mqc_sg_pat6(mqc_sg_pext(a, ma));
                if (S.w4a[(size_t)i4 * 6 + p6] == 0.0) continue;
                const unsigned sgn = __builtin_popcount(a & sa_a4[i4] & ~ma) & 1u;
                r2[i].push_back((unsigned)rank_a[a] | ((unsigned)i4 << 16) | ((unsigned)p6 << 28) | (sgn << 31));
            }
        }
        // links of one output row are processed MQC_SG_LB at a time against ONE beta ELL table:
        // order them by class and pad so that no batch mixes classes
        for (auto& r : r1) {
            std::stable_sort(r.begin(), r.end(), [&](unsigned x, unsigned y) {
                return S.cls_of_m2[(x >> 16) & 0xffu] < S.cls_of_m2[(y >> 16) & 0xffu]; });
            std::vector<unsigned> padded;
            for (size_t k = 0; k < r.size(); ++k) {
                if (k && S.cls_of_m2[(r[k] >> 16) & 0xffu] != S.cls_of_m2[(r[k - 1] >> 16) & 0xffu])
                    while (padded.size() % MQC_SG_LB) padded.push_back(MQC_SG_INVALID);
                padded.push_back(r[k]);
            }
            while (padded.size() % MQC_SG_LB) padded.push_back(MQC_SG_INVALID);
            r.swap(padded);
        }
        size_t e1 = 0, e2 = 0;
        for (auto& r : r1) e1 = std::max(e1, r.size());
        for (auto& r : r2) e2 = std::max(e2, r.size());
        S.e1a = (int)e1; S.e2a = (int)e2;
        S.lnk_a1.assign(std::max<size_t>(1, e1 * S.n_a), MQC_SG_INVALID);
        S.lnk_a2.assign(std::max<size_t>(1, e2 * S.n_a), MQC_SG_INVALID);
        for (unsigned i = 0; i < S.n_a; ++i) {
            for (size_t e = 0; e < r1[i].size(); ++e) S.lnk_a1[e * S.n_a + i] = r1[i][e];
            for (size_t e = 0; e < r2[i].size(); ++e) S.lnk_a2[e * S.n_a + i] = r2[i][e];
        }
    }
    if (S.ell_b1.empty()) S.ell_b1.push_back((unsigned)S.n_m2 << 16);
    return S;
}

// Reference evaluation of the tables on the host (one thread; tests and debugging only).
// L = (fast part of H) C for rows [0, n_a); `hdiag` is evaluated from the diagonal terms.
static inline void mqc_sigma_apply_host(const MqcSigmaHost& S, const double2* C, double2* L) {
    const unsigned nA = S.n_a, nB = S.n_b;
    const int M2 = std::max(1, S.n_m2);
    for (unsigned i = 0; i < nA; ++i) {
        const unsigned ap = S.aorb[i];
        for (unsigned j = 0; j < nB; ++j) {
            const unsigned bp = S.borb[j];
            double ar = 0, ai = 0;
            auto add = [&](double w, unsigned si, unsigned sj) { ar += w * C[(size_t)si * nB + sj].x; ai += w * C[(size_t)si * nB + sj].y; };
            // diagonal
            {
                const u64 x = S.aq[i] | S.bq[j];
                double d = 0;
                for (size_t t = 0; t < S.diag_z.size(); ++t) d += (__builtin_popcountll(x & S.diag_z[t]) & 1) ? -S.diag_c[t] : S.diag_c[t];
                add(d, i, j);
            }
            // phase 0: beta singles / beta-beta doubles on the own row
            const int c0 = S.cls_of_m2[S.n_m2];
            if (c0 >= 0) for (int e = S.cls_off[c0]; e < S.cls_off[c0 + 1]; ++e) {
                const unsigned en = S.ell_b1[(size_t)e * nB + j];
                const unsigned col = en & 0xffffu, ib = (en >> 16) & 0xffu, pbv = (en >> 24) & 1u, sg = (en >> 25) & 1u;
                if (ib == (unsigned)S.n_m2) continue;                    // padding
                const unsigned sx = __builtin_popcount(ap & S.sa_b2[ib]) & 1u;
                double w = S.tab[((size_t)ib * 2 + pbv) * nA + i] + S.tbb[(size_t)ib * nB + col];
                if ((sx + sg) & 1u) w = -w;
                add(w, i, col);
            }
            for (int e = 0; e < S.e2b; ++e) {
                const unsigned en = S.ell_b2[(size_t)e * nB + j];
                const unsigned col = en & 0xffffu, i4 = (en >> 16) & 0xfffu, p6 = (en >> 28) & 7u, sg = en >> 31;
                if (i4 == (unsigned)S.n_m4) continue;                    // padding
                const unsigned sx = __builtin_popcount(ap & S.sa_b4[i4]) & 1u;
                double w = S.w4b[(size_t)i4 * 6 + p6];
                if ((sx + sg) & 1u) w = -w;
                add(w, i, col);
            }
            // phase 1: alpha single links
            for (int l = 0; l < S.e1a; ++l) {
                const unsigned ln = S.lnk_a1[(size_t)l * nA + i];
                if (ln == MQC_SG_INVALID) continue;
                const unsigned src = ln & 0xffffu, ia = (ln >> 16) & 0xffu, pav = (ln >> 24) & 1u, sga = (ln >> 25) & 1u;
                const unsigned a = S.aorb[src], ma = S.m2_mask[ia];
                const unsigned f1 = __builtin_popcount(bp & S.sb_a2[ia]) & 1u;
                double accr = 0, acci = 0;
                const int c = S.cls_of_m2[ia];
                if (c >= 0) for (int e = S.cls_off[c]; e < S.cls_off[c + 1]; ++e) {
                    const unsigned en = S.ell_b1[(size_t)e * nB + j];
                    const unsigned col = en & 0xffffu, ib = (en >> 16) & 0xffu, pbv = (en >> 24) & 1u, sg = (en >> 25) & 1u;
                    if (ib == (unsigned)S.n_m2) continue;                // padding
                    const unsigned mb = S.m2_mask[ib];
                    double w = S.w22[((size_t)ia * M2 + ib) * 4 + pav * 2 + pbv];
                    const unsigned sxa = __builtin_popcount(a & S.sa_b2[ib] & ~ma) & 1u;
                    const unsigned destpat = bp & mb;                         // output string's bits on mb
                    const unsigned sxb = __builtin_popcount(destpat & S.sb_a2[ia]) & 1u;
                    if ((sxa + sxb + sg) & 1u) w = -w;
                    accr += w * C[(size_t)src * nB + col].x; acci += w * C[(size_t)src * nB + col].y;
                }
                {   // alpha single group itself: diagonal in beta
                    const double w = S.ta1[(size_t)ia * nA + src] + S.tb1[((size_t)ia * 2 + pav) * nB + j];
                    accr += w * C[(size_t)src * nB + j].x; acci += w * C[(size_t)src * nB + j].y;
                }
                if ((f1 + sga) & 1u) { accr = -accr; acci = -acci; }
                ar += accr; ai += acci;
            }
            // phase 2: alpha-alpha double links
            for (int l = 0; l < S.e2a; ++l) {
                const unsigned ln = S.lnk_a2[(size_t)l * nA + i];
                if (ln == MQC_SG_INVALID) continue;
                const unsigned src = ln & 0xffffu, i4 = (ln >> 16) & 0xfffu, p6 = (ln >> 28) & 7u, sga = ln >> 31;
                const unsigned f1 = __builtin_popcount(bp & S.sb_a4[i4]) & 1u;
                double w = S.w4a[(size_t)i4 * 6 + p6];
                if ((f1 + sga) & 1u) w = -w;
                add(w, src, j);
            }
            L[(size_t)i * nB + j] = make_double2(ar, ai);
        }
    }
}

// ---- device side ---------------------------------------------------------------------------
#if defined(__CUDACC__)
#define MQC_SG_R 4           // output beta strings per thread
struct __align__(16) MqcSgA2 { unsigned src, zb; double w; };   // one resolved alpha-alpha double link

// hdiag[row - row0][jb] = sum over diagonal Pauli terms c_t (-1)^popc(x & z_t)
__global__ void k_sigma_hdiag(const u64* __restrict__ aq, const u64* __restrict__ bq, const u64* __restrict__ dz,
                              const double* __restrict__ dc, const int n_terms, const unsigned row0, const unsigned row1,
                              const unsigned n_b, double* __restrict__ hdiag) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (u64)(row1 - row0) * n_b) return;
    const unsigned i = row0 + (unsigned)(tid / n_b), j = (unsigned)(tid % n_b);
    const u64 x = aq[i] | bq[j];
    double d = 0.0;
    for (int t = 0; t < n_terms; ++t) {
        const double c = __ldg(dc + t);
        d += (__popcll(x & __ldg(dz + t)) & 1) ? -c : c;
    }
    hdiag[tid] = d;
}

__device__ __forceinline__ double mqc_sg_flip(double w, unsigned s) {      // s odd -> -w
    return __hiloint2double(__double2hiint(w) ^ (int)(s << 31), __double2loint(w));
}

// One CTA per (output alpha string, block of MQC_SG_THREADS * MQC_SG_R output beta strings).
// STAGE: the source row of C is copied to shared memory per link (rows up to 64 KB); otherwise
// the gathers go to L1/L2 directly.
template <bool REAL> struct MqcSgAmp { typedef double2 T; };
template <> struct MqcSgAmp<true> { typedef double T; };
__device__ __forceinline__ void mqc_sg_acc(const double w, const double2 v, double& ar, double& ai) { ar = fma(w, v.x, ar); ai = fma(w, v.y, ai); }
__device__ __forceinline__ void mqc_sg_acc(const double w, const double v, double& ar, double&) { ar = fma(w, v, ar); }
__device__ __forceinline__ double2 mqc_sg_flipv(const double2 v, const unsigned s) { return make_double2(mqc_sg_flip(v.x, s), mqc_sg_flip(v.y, s)); }
__device__ __forceinline__ double mqc_sg_flipv(const double v, const unsigned s) { return mqc_sg_flip(v, s); }
__device__ __forceinline__ void mqc_sg_store(double2* p, double ar, double ai) { *p = make_double2(ar, ai); }
__device__ __forceinline__ void mqc_sg_store(double* p, double ar, double) { *p = ar; }
__device__ __forceinline__ void mqc_sg_add(double2* p, double ar, double ai) { const double2 o = *p; *p = make_double2(o.x + ar, o.y + ai); }
__device__ __forceinline__ void mqc_sg_add(double* p, double ar, double) { *p += ar; }
__device__ __forceinline__ void mqc_sg_atomic(double2* p, double ar, double ai) { atomicAdd(&p->x, ar); atomicAdd(&p->y, ai); }
__device__ __forceinline__ void mqc_sg_atomic(double* p, double ar, double) { atomicAdd(p, ar); }
__device__ __forceinline__ void mqc_sg_dot(const double2 py, double ar, double ai, double& er, double& ei) {
    er += py.x * ar + py.y * ai; ei += py.x * ai - py.y * ar;
}
__device__ __forceinline__ void mqc_sg_dot(const double py, double ar, double, double& er, double&) { er += py * ar; }

template <bool STAGE, int LB, bool REAL>
__global__ void __launch_bounds__(MQC_SG_THREADS, 3)
k_sigma(const typename MqcSgAmp<REAL>::T* __restrict__ Cown, const typename MqcSgAmp<REAL>::T* __restrict__ Csrc,
        typename MqcSgAmp<REAL>::T* __restrict__ L, const __grid_constant__ MqcSigmaDev S, double* __restrict__ e_out) {
    typedef typename MqcSgAmp<REAL>::T AT;
    extern __shared__ __align__(16) unsigned char sg_smem[];
    const size_t pitch = S.pitch ? S.pitch : S.n_b;                       // row pitch of the three state buffers
    const AT* __restrict__ C = Csrc - (size_t)S.src_lo * pitch;           // C[src * pitch + j] for src in the block
    const unsigned nA = S.n_a, nB = S.n_b;
    const int M2 = S.n_m2 > 0 ? S.n_m2 : 1;
    const int WS = 2 * (S.n_m2 + 1);                               // weight-table stride; last pair = 0 (padding)
    AT* srow = reinterpret_cast<AT*>(sg_smem);
    double* wl = reinterpret_cast<double*>(sg_smem + (STAGE ? (((size_t)LB * nB * sizeof(AT) + 15) & ~(size_t)15) : 0));
    double* sxd = wl + (size_t)LB * WS;
    unsigned char* sg4 = reinterpret_cast<unsigned char*>(sxd + (S.n_m2 + 1));
    __shared__ double red[2][MQC_SG_THREADS / 32];

    const int tid = threadIdx.x;
    const unsigned i = S.row0 + blockIdx.x;
    const unsigned ap = __ldg(S.aorb + i);
    unsigned jb[MQC_SG_R], jc[MQC_SG_R], bp[MQC_SG_R];              // jc: clamped column for loads of idle threads
    double ar[MQC_SG_R], ai[MQC_SG_R];
#pragma unroll
    for (int r = 0; r < MQC_SG_R; ++r) {
        jb[r] = (blockIdx.y * MQC_SG_R + r) * MQC_SG_THREADS + tid;
        jc[r] = jb[r] < nB ? jb[r] : nB - 1u;
        bp[r] = __ldg(S.borb + jc[r]);
        ar[r] = 0.0; ai[r] = 0.0;
    }

    // ---- phase 0: beta-only part on the own row ---------------------------------------------
    const bool part0 = blockIdx.z == 0, part1 = blockIdx.z + 1 == gridDim.z;
    const int l_begin = part0 ? 0 : S.split_l, l_end = (part0 && !part1) ? S.split_l : S.e1a;
    if (part0 && i >= S.src_lo && i < S.src_hi) {               // uniform over the CTA
        const int c0 = __ldg(S.cls_of_m2 + S.n_m2);
        for (int k = tid; k < WS; k += MQC_SG_THREADS) {
            const int ib = k >> 1, pbv = k & 1;
            double w = 0.0, sd = 0.0;
            if (ib < S.n_m2) {
                const unsigned sx = __popc(ap & __ldg(S.sa_b2 + ib)) & 1u;
                w = mqc_sg_flip(__ldg(S.tab + ((size_t)ib * 2 + pbv) * nA + i), sx);
                sd = sx ? -1.0 : 1.0;
            }
            wl[k] = w;
            if (pbv == 0) sxd[ib] = sd;
        }
        for (int k = tid; k <= S.n_m4; k += MQC_SG_THREADS) sg4[k] = (unsigned char)(__popc(ap & __ldg(S.sa_b4 + k)) & 1u);
        const AT* __restrict__ Crow = C + (size_t)i * pitch;
        if (STAGE) for (unsigned j = tid; j < nB; j += MQC_SG_THREADS) srow[j] = Crow[j];
        __syncthreads();
        const AT* __restrict__ row = STAGE ? srow : Crow;
        if (c0 >= 0) {
            const int e0 = __ldg(S.cls_off + c0), e1 = __ldg(S.cls_off + c0 + 1);
            unsigned nx[MQC_SG_R];
#pragma unroll
            for (int r = 0; r < MQC_SG_R; ++r) nx[r] = e0 < e1 ? __ldg(S.ell_b1 + (size_t)e0 * nB + jc[r]) : 0u;
            for (int e = e0; e < e1; ++e) {
                unsigned en[MQC_SG_R];
#pragma unroll
                for (int r = 0; r < MQC_SG_R; ++r) { en[r] = nx[r]; if (e + 1 < e1) nx[r] = __ldg(S.ell_b1 + (size_t)(e + 1) * nB + jc[r]); }
#pragma unroll
                for (int r = 0; r < MQC_SG_R; ++r) {
                    const unsigned col = en[r] & 0xffffu, ib = (en[r] >> 16) & 0xffu;
                    const unsigned widx = (ib << 1) | ((en[r] >> 24) & 1u);
                    double w = fma(sxd[ib], __ldg(S.tbb + (size_t)ib * nB + col), wl[widx]);
                    w = mqc_sg_flip(w, (en[r] >> 25) & 1u);
                    const AT v = row[col];
                    mqc_sg_acc(w, v, ar[r], ai[r]);
                }
            }
        }
        {
            // entry e + 2 and the weight of entry e + 1 are in flight while entry e is applied
            unsigned nx[MQC_SG_R], en[MQC_SG_R];
            double wn[MQC_SG_R];
#pragma unroll
            for (int r = 0; r < MQC_SG_R; ++r) {
                en[r] = S.e2b > 0 ? __ldg(S.ell_b2 + jc[r]) : 0u;
                nx[r] = S.e2b > 1 ? __ldg(S.ell_b2 + (size_t)nB + jc[r]) : 0u;
                wn[r] = S.e2b > 0 ? __ldg(S.w4b + (size_t)((en[r] >> 16) & 0xfffu) * 6 + ((en[r] >> 28) & 7u)) : 0.0;
            }
            for (int e = 0; e < S.e2b; ++e) {
                unsigned cur[MQC_SG_R];
                double wc[MQC_SG_R];
#pragma unroll
                for (int r = 0; r < MQC_SG_R; ++r) {
                    cur[r] = en[r]; wc[r] = wn[r]; en[r] = nx[r];
                    if (e + 2 < S.e2b) nx[r] = __ldg(S.ell_b2 + (size_t)(e + 2) * nB + jc[r]);
                    if (e + 1 < S.e2b) wn[r] = __ldg(S.w4b + (size_t)((en[r] >> 16) & 0xfffu) * 6 + ((en[r] >> 28) & 7u));
                }
#pragma unroll
                for (int r = 0; r < MQC_SG_R; ++r) {
                    const unsigned col = cur[r] & 0xffffu, i4 = (cur[r] >> 16) & 0xfffu;
                    const double w = mqc_sg_flip(wc[r], (unsigned)sg4[i4] ^ (cur[r] >> 31));
                    const AT v = row[col];
                    mqc_sg_acc(w, v, ar[r], ai[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < MQC_SG_R; ++r) {
            const double d = __ldg(S.hdiag + (size_t)(i - S.row0) * nB + jc[r]);
            const AT v = row[jc[r]];
            mqc_sg_acc(d, v, ar[r], ai[r]);
        }
    }

    // ---- phase 1: alpha single links (alpha-beta doubles + the alpha single groups) ----------
    // LB links per batch: their source rows are staged together and every ELL entry (column,
    // weight index, sign) is loaded and decoded once for all of them; the next entry is in
    // flight while the current one is applied.
    for (int l0 = l_begin; l0 < l_end; l0 += LB) {
        unsigned src[LB], ia[LB], pav[LB], sga[LB], zb[LB];
        bool ok[LB];
        bool any = false;
        int c = -1;
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const unsigned ln = l0 + l < l_end ? __ldg(S.lnk_a1 + (size_t)(l0 + l) * nA + i) : MQC_SG_INVALID;
            ok[l] = ln != MQC_SG_INVALID && (ln & 0xffffu) >= S.src_lo && (ln & 0xffffu) < S.src_hi;
            src[l] = ok[l] ? (ln & 0xffffu) : S.src_lo;
            ia[l] = ok[l] ? ((ln >> 16) & 0xffu) : 0u;
            pav[l] = (ln >> 24) & 1u; sga[l] = (ln >> 25) & 1u;
            zb[l] = __ldg(S.sb_a2 + ia[l]);
            if (ok[l]) { any = true; c = __ldg(S.cls_of_m2 + ia[l]); }
        }
        if (!any) continue;                                       // uniform over the CTA
        __syncthreads();                                          // previous consumers of srow / wl are done
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const unsigned a = __ldg(S.aorb + src[l]), ma = __ldg(S.m2_mask + ia[l]);
            for (int k = tid; k < WS; k += MQC_SG_THREADS) {
                const int ib = k >> 1, pbv = k & 1;
                double w = 0.0;
                if (ok[l] && ib < S.n_m2) {
                    const unsigned mb = __ldg(S.m2_mask + ib);
                    const unsigned dest = pbv ? (mb & (0u - mb)) : (mb & (mb - 1u));   // output string's bit on mb
                    // STAGE: the staged row carries (-1)^popc(source beta string & zb), the output string differs
                    // from it by mb, and the link's own sign goes here too: no per-(output, link) sign is left
                    const unsigned sx = (__popc(a & __ldg(S.sa_b2 + ib) & ~ma) + __popc(dest & zb[l]) +
                                         (STAGE ? __popc(mb & zb[l]) + sga[l] : 0u)) & 1u;
                    w = mqc_sg_flip(__ldg(S.w22 + ((size_t)ia[l] * M2 + ib) * 4 + pav[l] * 2 + pbv), sx);
                }
                wl[l * WS + k] = w;
            }
            if (STAGE) {
                const AT* __restrict__ Crow = C + (size_t)src[l] * pitch;
                const unsigned z = zb[l];
#pragma unroll 4
                for (unsigned j = tid; j < nB; j += MQC_SG_THREADS) {
                    const AT v = Crow[j];
                    const unsigned sg = __popc(__ldg(S.borb + j) & z) & 1u;
                    srow[(size_t)l * nB + j] = mqc_sg_flipv(v, sg);
                }
            }
        }
        __syncthreads();
        const AT* row[LB];
        unsigned sl[MQC_SG_R];                                    // bit l: sign of link l for output r
#pragma unroll
        for (int r = 0; r < MQC_SG_R; ++r) sl[r] = 0u;
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            row[l] = STAGE ? (srow + (size_t)l * nB) : (C + (size_t)src[l] * pitch);
            const double al = ok[l] ? __ldg(S.ta1 + (size_t)ia[l] * nA + src[l]) : 0.0;
#pragma unroll
            for (int r = 0; r < MQC_SG_R; ++r) {
                const unsigned sg = STAGE ? sga[l] : ((__popc(bp[r] & zb[l]) + sga[l]) & 1u);
                if (!STAGE) sl[r] |= sg << l;
                if (ok[l]) {
                    const double w = mqc_sg_flip(al + __ldg(S.tb1 + ((size_t)ia[l] * 2 + pav[l]) * nB + jc[r]), sg);
                    const AT v = row[l][jc[r]];
                    mqc_sg_acc(w, v, ar[r], ai[r]);
                }
            }
        }
        // c < 0: these alpha masks have no beta partner (the opposite-spin doubles live elsewhere or vanish)
        const int e0 = c >= 0 ? __ldg(S.cls_off + c) : 0, e1 = c >= 0 ? __ldg(S.cls_off + c + 1) : 0;
        unsigned nx[MQC_SG_R];
#pragma unroll
        for (int r = 0; r < MQC_SG_R; ++r) nx[r] = e0 < e1 ? __ldg(S.ell_b1 + (size_t)e0 * nB + jc[r]) : 0u;
        for (int e = e0; e < e1; ++e) {
            unsigned en[MQC_SG_R];
#pragma unroll
            for (int r = 0; r < MQC_SG_R; ++r) { en[r] = nx[r]; if (e + 1 < e1) nx[r] = __ldg(S.ell_b1 + (size_t)(e + 1) * nB + jc[r]); }
#pragma unroll
            for (int r = 0; r < MQC_SG_R; ++r) {
                const unsigned col = en[r] & 0xffffu;
                const unsigned widx = ((en[r] >> 16) & 0xffu) * 2u + ((en[r] >> 24) & 1u);
                const unsigned sg = (en[r] >> 25) & 1u;
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    const double w = mqc_sg_flip(wl[l * WS + widx], STAGE ? sg : (sg ^ (sl[r] >> l)));
                    const AT v = row[l][col];
                    mqc_sg_acc(w, v, ar[r], ai[r]);
                }
            }
        }
    }

    // ---- phase 2: alpha-alpha double links (diagonal in beta, coalesced row reads) -----------
    // The row's links are resolved once into shared memory (source row, signed weight, JW beta mask; links
    // that are absent or whose source row another rank holds are squeezed out, in order), so the row
    // reads of consecutive links are independent loads.
    if (part1) {
        MqcSgA2* a2 = reinterpret_cast<MqcSgA2*>((reinterpret_cast<uintptr_t>(sg4) + (size_t)S.n_m4 + 1 + 15) & ~(uintptr_t)15);
        __shared__ int wcnt[MQC_SG_THREADS / 32];
        const int warp = tid >> 5, lane = tid & 31;
        int n2 = 0;
        __syncthreads();                                          // phase 1 is done with the shared tables
        for (int l0 = 0; l0 < S.e2a; l0 += MQC_SG_THREADS) {
            const int l = l0 + tid;
            MqcSgA2 ent;
            bool ok = false;
            if (l < S.e2a) {
                const unsigned ln = __ldg(S.lnk_a2 + (size_t)l * nA + i);
                const unsigned src = ln & 0xffffu, i4 = (ln >> 16) & 0xfffu, p6 = (ln >> 28) & 7u;
                ok = ln != MQC_SG_INVALID && src >= S.src_lo && src < S.src_hi;
                if (ok) { ent.src = src; ent.zb = __ldg(S.sb_a4 + i4); ent.w = mqc_sg_flip(__ldg(S.w4a + (size_t)i4 * 6 + p6), ln >> 31); }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) wcnt[warp] = __popc(bal);
            __syncthreads();
            int off = n2 + __popc(bal & ((1u << lane) - 1u)), tot = 0;
#pragma unroll
            for (int w = 0; w < MQC_SG_THREADS / 32; ++w) { const int c = wcnt[w]; if (w < warp) off += c; tot += c; }
            if (ok) a2[off] = ent;
            n2 += tot;
            __syncthreads();
        }
#pragma unroll 2
        for (int l = 0; l < n2; ++l) {
            const MqcSgA2 ent = a2[l];
            const AT* __restrict__ Crow = C + (size_t)ent.src * pitch;
#pragma unroll
            for (int r = 0; r < MQC_SG_R; ++r) {
                const AT v = Crow[jc[r]];
                const double ws = mqc_sg_flip(ent.w, __popc(bp[r] & ent.zb) & 1u);
                mqc_sg_acc(ws, v, ar[r], ai[r]);
            }
        }
    }

    // ---- output ----------------------------------------------------------------------------
    double er = 0.0, ei = 0.0;
#pragma unroll
    for (int r = 0; r < MQC_SG_R; ++r) {
        if (jb[r] < nB) {
            if (L) {
                const size_t o = (size_t)(i - S.l_lo) * pitch + jb[r];
                if (gridDim.z > 1) mqc_sg_atomic(L + o, ar[r], ai[r]);
                else if (S.accumulate) mqc_sg_add(L + o, ar[r], ai[r]);
                else mqc_sg_store(L + o, ar[r], ai[r]);
            }
            mqc_sg_dot(Cown[(size_t)(i - S.own_lo) * pitch + jb[r]], ar[r], ai[r], er, ei);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { er += __shfl_xor_sync(0xffffffffu, er, o); ei += __shfl_xor_sync(0xffffffffu, ei, o); }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { red[0][warp] = er; red[1][warp] = ei; }
    __syncthreads();
    if (warp == 0) {
        double vr = (lane < MQC_SG_THREADS / 32) ? red[0][lane] : 0.0;
        double vi = (lane < MQC_SG_THREADS / 32) ? red[1][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { vr += __shfl_xor_sync(0xffffffffu, vr, o); vi += __shfl_xor_sync(0xffffffffu, vi, o); }
        if (!S.e_part) {
            if (lane == 0) {
                if (vr != 0.0) atomicAdd(e_out, vr);
                if (vi != 0.0) atomicAdd(e_out + 1, vi);
            }
        } else {
            const unsigned total = gridDim.x * gridDim.y * gridDim.z;
            const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            unsigned last = 0u;
            if (lane == 0) {
                S.e_part[2 * (size_t)cta] = vr; S.e_part[2 * (size_t)cta + 1] = vi;
                __threadfence();
                last = atomicAdd(S.ticket, 1u) == total - 1u ? 1u : 0u;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {                                               // every other CTA's partial is visible (fence + ticket)
                __threadfence();
                double sr = 0.0, si = 0.0;
                for (unsigned c = lane; c < total; c += 32) { sr += __ldcg(S.e_part + 2 * (size_t)c); si += __ldcg(S.e_part + 2 * (size_t)c + 1); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
                if (lane == 0) {
                    if (sr != 0.0) atomicAdd(e_out, sr);
                    if (si != 0.0) atomicAdd(e_out + 1, si);
                    *S.ticket = 0u;                                   // ready for the next launch on this stream
                }
            }
        }
    }
}

static inline size_t mqc_sigma_smem_bytes(const MqcSigmaDev& S, bool stage, int lb, int elem = 16) {
    const int M2 = S.n_m2 > 0 ? S.n_m2 : 1;
    size_t b = stage ? (((size_t)lb * S.n_b * elem + 15) & ~(size_t)15) : 0;
    (void)M2;
    b += ((size_t)lb * 2 * (S.n_m2 + 1) + (size_t)(S.n_m2 + 1)) * sizeof(double);
    b += (((size_t)S.n_m4 + 1 + 15) & ~(size_t)15) + 16;
    b += (size_t)(S.e2a > 0 ? S.e2a : 1) * 16;                  // resolved alpha-alpha links (MqcSgA2)
    return (b + 15) & ~(size_t)15;
}
#endif
