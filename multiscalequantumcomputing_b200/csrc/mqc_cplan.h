// Compact-sector sweep plan (host side, pure C++).
//
// A number- and S_z-conserving ansatz only ever populates the determinants of one
// (N_alpha, N_beta) sector.  The compact sweep keeps the state as the dense matrix
//     C[row][col],   row <-> alpha string,   col <-> beta string          (16 B per determinant)
// (13.7 MB at 24 qubits instead of the 268 MB dense register, 37.8 GB at 36 qubits instead of
// 1.1 TB), i.e. the (reordered) CI vector that K2 (mqc_sigma.h) and K4 consume directly.
//
// The pass structure is the scheduler's (mqc_schedule.h): a pass keeps a set of ACTIVE alpha
// orbitals S_A and beta orbitals S_B and applies every factor that moves only active orbitals.
// All alpha strings with the same occupation OUTSIDE S_A form a class; within a class the
// strings differ by a k-subset of S_A ("pattern").  The layout of pass p orders rows class by
// class (ascending inactive mask), patterns ascending inside a class - and likewise columns -
// so that a TILE = (alpha class) x (beta class) is a contiguous box of C: cnt_A rows of cnt_B
// contiguous amplitudes each, which the kernel fetches with one bulk (TMA) copy per row.  A
// pass is out of place: it reads the layout of pass p and scatters its result into the layout
// of pass p + 1 through a row map and a column map (the last pass writes the canonical
// ascending-string order).  The reverse (adjoint) sweep runs the same tiles backwards:
// gather through the maps, store contiguous.
//
// Per (pass, ka, kb) - the tile CLASS: all tiles with ka / kb electrons on the active alpha /
// beta orbitals have the same internal structure - the host compiles a PROGRAM: the live factors
// in sequence order, each with the explicit list of amplitude pairs it rotates,
//     pairs of factor f in tile (ka, kb)  =  listA(f, ka)  x  listB(f, kb)
// (a factor that does not act on a spin species couples every pattern with itself), one 32-bit
// word per pair: slot of the lower partner | slot of the upper partner << 14 | in-tile parity
// << 28.  A thread of the kernel fetches its word, coalesced and several factors ahead, and
// has nothing to decode.  ~8 MB for the whole 24-qubit UCCSD sweep (L2 resident).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "mqc_common.h"
#include "mqc_schedule.h"

#define MQC_C_KMAX 17

struct MqcCClass {             // strings of one spin species with the same inactive occupation
    uint32_t row0;             // first row (column) of the class in this pass's layout
    uint32_t mask;             // occupied inactive orbitals (orbital-space mask)
    uint16_t k, cnt;           // electrons on the active orbitals, number of patterns C(act, k)
    uint32_t pad;
};
struct MqcCFactor {
    uint32_t zA_out, zB_out;   // Jordan-Wigner parity masks over the inactive alpha / beta orbitals
    int32_t slot;              // cos/sin table entry
    int32_t theta_sign;        // theta index | sign0 << 31
};
struct MqcCProgHdr {           // one live factor (step) of a class program
    uint32_t off;              // first pair word, relative to the program's first word
    uint32_t np_f;             // pairs (11 bits) | pass-local factor index << 11 (12 bits) | ring bookkeeping << 23:
                               //   2 bits each: chunks to wait for before this step (forward order / adjoint order),
                               //   chunks to issue after this step (forward / adjoint); see mqc_c_ring_flags
};
#define MQC_C_NP(y) ((y) & 0x7ffu)
#define MQC_C_F(y) (((y) >> 11) & 0xfffu)
struct MqcCLive {              // class (ka, kb) -> its program
    uint32_t off, cnt;         // headers in the header pool
    uint32_t word_base;        // first pair word in the program pool (multiple of 4: 16-byte aligned)
    uint32_t n_words;          // pair words of all steps, padded to a multiple of 4
};
// The pair words of a class program are contiguous in step order: the kernel streams them through
// a shared-memory ring of MQC_C_RING_SLOTS chunks of MQC_C_CHUNK words (bulk copies, one mbarrier
// per slot); a step may straddle at most MQC_C_RING_SLOTS - 1 chunks.
// The chunk is 256 words when every step has <= 384 pairs (7 + 7 active orbitals), else 1024.
#define MQC_C_RING_SLOTS 4
#define MQC_C_MAX_STEP_PAIRS(chunk) (((MQC_C_RING_SLOTS - 1) * (chunk)) / 2)     // two consecutive steps fit the ring
struct MqcCPassDev {
    int32_t nf, n_clsA, n_clsB, strideB;
    int32_t capA, capB, row_identity, col_identity;
    uint32_t n_tiles, tiles_explicit;
    int32_t max_live, f0, pad0, pad1;                    // longest class program (factors); first factor (sequence position)
    uint64_t clsA_off, clsB_off, fac_off, map_row_off, map_col_off, tile_off, live_tab_off;
};

struct MqcCPlan {
    bool ok = false;
    int n_orb = 0, na = 0, nb = 0;
    uint32_t nA = 0, nB = 0;                 // rows, columns
    uint32_t ref_row = 0, ref_col = 0;       // reference determinant in the layout of pass 0
    std::vector<MqcCPassDev> pass;
    std::vector<MqcCClass> cls;
    std::vector<MqcCFactor> fac;
    std::vector<uint32_t> maps, tiles;
    std::vector<MqcCLive> live_tab;
    std::vector<MqcCProgHdr> prog_hdr;       // per class: headers of its live factors, sequence order (even count: 16-byte units)
    std::vector<uint32_t> prog;              // pair words
    std::vector<uint32_t> astr, bstr;        // canonical (ascending) strings, orbital space
    int chunk = 256;                         // pair words per ring slot (256 or 1024)
    uint32_t max_step_pairs = 0;
    uint32_t pitch = 0;                      // elements between consecutive rows of the device matrix: nB, rounded up to even for 8-byte amplitudes
    int elem = 16;                           // bytes per amplitude the programs were laid out for (row stride, bank order)
    size_t max_smem[2] = {0, 0};             // shared memory a CTA needs: forward, adjoint (any pass)
    const char* why = "";                    // why ok == false
};

MQC_HD uint32_t mqc_c_magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)(0xFFFFFFFFu / d) + 1u; }

static inline void mqc_c_strings(int n_orb, int k, std::vector<uint32_t>& out) {
    out.clear();
    if (k < 0 || k > n_orb) return;
    if (k == 0) { out.push_back(0); return; }
    u64 v = (1ull << k) - 1ull;
    const u64 limit = 1ull << n_orb;
    while (v < limit) {
        out.push_back((uint32_t)v);
        const u64 c = v & (~v + 1ull), r = v + c;
        v = (((r ^ v) >> 2) / c) | r;
    }
}

// The pairs of a step are independent, so their order is free: arrange them so that the eight
// threads of a quarter-warp (one 128-byte shared-memory wavefront of 16-byte amplitudes) hit eight
// different bank groups, for the lower AND the upper partner (greedy, bounded look-ahead).
static inline void mqc_c_bank_order(std::vector<uint32_t>& w, int elem = 16) {
    const size_t n = w.size();
    if (n < 3) return;
    std::vector<char> used(n, 0);
    std::vector<uint32_t> out;
    out.reserve(n);
    size_t first = 0;
    while (out.size() < n) {
        unsigned m0 = 0, m1 = 0;
        // 8-byte (real) amplitudes: sixteen threads share a wavefront, bank pair = slot mod 16
        const int lanes = elem == 8 ? 16 : 8;
        const unsigned bm = (unsigned)lanes - 1u;
        for (int slot = 0; slot < lanes && out.size() < n; ++slot) {
            while (first < n && used[first]) ++first;
            size_t pick = n, seen = 0;
            for (size_t c = first; c < n && seen < 96; ++c) {
                if (used[c]) continue;
                ++seen;
                const unsigned b0 = 1u << (w[c] & bm), b1 = 1u << ((w[c] >> 14) & bm);
                if (!(m0 & b0) && !(m1 & b1)) { pick = c; break; }
            }
            if (pick == n) pick = first;
            used[pick] = 1;
            m0 |= 1u << (w[pick] & bm); m1 |= 1u << ((w[pick] >> 14) & bm);
            out.push_back(w[pick]);
        }
    }
    w.swap(out);
}

// Ring bookkeeping of a class program, precomputed so that the kernel neither divides nor compares:
// the chunks arrive in processing order (forward: c = 0, 1, ..; adjoint: c = NC - 1, NC - 2, ..); a step
// waits for the chunks it is the first to need and, once it is done, the chunks that the NEXT step
// no longer needs make room for as many new ones.
static inline void mqc_c_ring_flags(MqcCProgHdr* H, int n, uint32_t n_words, int chunk) {
    const int NC = (int)((n_words + chunk - 1) / chunk);
    auto c_lo = [&](int i) { return (int)(H[i].off / chunk); };
    auto c_hi = [&](int i) { return (int)((H[i].off + MQC_C_NP(H[i].np_f) - 1u) / chunk); };
    int have = 0, issued = std::min(NC, MQC_C_RING_SLOTS);
    for (int i = 0; i < n; ++i) {                              // forward order
        const int need = c_hi(i) + 1;
        const int w = std::max(0, need - have);
        have = std::max(have, need);
        int is = 0;
        if (i + 1 < n) { const int target = std::min(NC, c_lo(i + 1) + MQC_C_RING_SLOTS); is = std::max(0, target - issued); issued += is; }
        if (w > 3 || is > 3) throw std::runtime_error("program ring bookkeeping overflow");
        H[i].np_f |= ((uint32_t)w << 23) | ((uint32_t)is << 27);
    }
    have = 0; issued = std::min(NC, MQC_C_RING_SLOTS);
    for (int i = n - 1; i >= 0; --i) {                         // adjoint order, k = NC - 1 - c
        const int need = NC - 1 - c_lo(i) + 1;
        const int w = std::max(0, need - have);
        have = std::max(have, need);
        int is = 0;
        if (i > 0) { const int target = std::min(NC, NC - 1 - c_hi(i - 1) + MQC_C_RING_SLOTS); is = std::max(0, target - issued); issued += is; }
        if (w > 3 || is > 3) throw std::runtime_error("program ring bookkeeping overflow");
        H[i].np_f |= ((uint32_t)w << 25) | ((uint32_t)is << 29);
    }
}

// shared memory of one CTA (bytes) for a pass; must match the carve-up in k_csweep
static inline size_t mqc_c_smem_bytes(const MqcCPassDev& P, bool adj, int n_warps, int chunk, int elem = 16) {
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t b = 0;
    b += (size_t)(adj ? 2 : 1) * al((size_t)P.capA * P.strideB * elem);
    b += al(((size_t)P.max_live + 9) * 8);        // class program headers + sentinel slots
    b += (size_t)MQC_C_RING_SLOTS * chunk * 4;    // pair-word ring
    b += (size_t)P.nf * 16;                       // cos / sin
    b += al((size_t)P.nf);                        // per-tile factor signs
    b += al((size_t)(P.capA + P.capB) * 4);       // row / column maps of the tile
    if (adj) b += (size_t)n_warps * al((size_t)P.nf * 8);
    b += 64;                                      // mbarriers: tile + ring slots
    return b;
}

// fac: logical-coordinate factors (make_factor), sched: the pass structure built from them
// elem: bytes per amplitude in the device buffers, 16 (complex) or 8 (real: the ansatz factors are real
// rotations, so with a real Hamiltonian nothing ever has an imaginary part).  Real tiles take an even row
// stride with two spare slots per row: the bulk copies move whole 16-byte units, a row that starts on an odd
// column is loaded from the element before it (k_csweep).
static inline MqcCPlan mqc_build_cplan(const MqcSchedule& S, const std::vector<MqcFactor>& fac, int na, int nb, u64 ref_index, int elem = 16) {
    MqcCPlan PL;
    PL.elem = elem;
    const int n = S.n_bits;
    if (n & 1) return PL;
    const int n_orb = n / 2;
    if (n_orb > 31) return PL;
    PL.n_orb = n_orb; PL.na = na; PL.nb = nb;
    mqc_c_strings(n_orb, na, PL.astr);
    mqc_c_strings(n_orb, nb, PL.bstr);
    if (PL.astr.empty() || PL.bstr.empty()) return PL;
    if ((double)PL.astr.size() * (double)PL.bstr.size() > 4.0e9) return PL;
    PL.nA = (uint32_t)PL.astr.size(); PL.nB = (uint32_t)PL.bstr.size();
    PL.pitch = elem == 8 ? ((PL.nB + 1u) & ~1u) : PL.nB;
    // logical bit l = n - 1 - q, qubit q = 2 p + spin
    auto split = [&](u64 lmask, uint32_t& a, uint32_t& b) {
        a = b = 0;
        for (int q = 0; q < n; ++q)
            if ((lmask >> (n - 1 - q)) & 1ull) { if (q & 1) b |= 1u << (q / 2); else a |= 1u << (q / 2); }
    };
    const size_t np = S.passes.size();
    // active orbital sets per pass
    std::vector<uint32_t> SA(np), SB(np);
    for (size_t p = 0; p < np; ++p) {
        const std::vector<int>& ph = S.pass_phys_before[p];
        u64 lact = 0;
        for (int l = 0; l < n; ++l) if ((S.passes[p].act_mask >> ph[l]) & 1ull) lact |= 1ull << l;
        split(lact, SA[p], SB[p]);
    }
    // position of every canonical string in the layout of pass p (p == np: canonical order)
    auto layout_pos = [&](const std::vector<uint32_t>& str, uint32_t act, std::vector<uint32_t>& pos, std::vector<MqcCClass>& classes) {
        const size_t ns = str.size();
        std::vector<uint32_t> order(ns);
        for (size_t i = 0; i < ns; ++i) order[i] = (uint32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
            const uint32_t ux = str[x] & ~act, uy = str[y] & ~act;
            if (ux != uy) return ux < uy;
            return (str[x] & act) < (str[y] & act);
        });
        pos.assign(ns, 0);
        classes.clear();
        for (size_t r = 0; r < ns; ++r) {
            const uint32_t s = str[order[r]];
            pos[order[r]] = (uint32_t)r;
            if (classes.empty() || classes.back().mask != (s & ~act)) {
                MqcCClass c{};
                c.row0 = (uint32_t)r; c.mask = s & ~act; c.k = (uint16_t)__builtin_popcount(s & act); c.cnt = 0;
                classes.push_back(c);
            }
            classes.back().cnt++;
        }
    };
    std::vector<std::vector<uint32_t>> posA(np + 1), posB(np + 1);
    std::vector<std::vector<MqcCClass>> clsA(np), clsB(np);
    for (size_t p = 0; p < np; ++p) { layout_pos(PL.astr, SA[p], posA[p], clsA[p]); layout_pos(PL.bstr, SB[p], posB[p], clsB[p]); }
    posA[np].resize(PL.nA); posB[np].resize(PL.nB);
    for (uint32_t i = 0; i < PL.nA; ++i) posA[np][i] = i;
    for (uint32_t i = 0; i < PL.nB; ++i) posB[np][i] = i;
    {
        uint32_t ra, rb;
        split(ref_index, ra, rb);
        const auto ia = std::lower_bound(PL.astr.begin(), PL.astr.end(), ra), ib = std::lower_bound(PL.bstr.begin(), PL.bstr.end(), rb);
        if (ia == PL.astr.end() || *ia != ra || ib == PL.bstr.end() || *ib != rb) return PL;
        PL.ref_row = posA[0][ia - PL.astr.begin()];
        PL.ref_col = posB[0][ib - PL.bstr.begin()];
    }
    for (size_t p = 0; p < np; ++p) {
        const MqcPass& P = S.passes[p];
        MqcCPassDev D{};
        D.nf = P.f1 - P.f0;
        D.f0 = P.f0;
        if (D.nf > 4095) { PL.why = "more than 4095 factors in a pass"; return PL; }
        const int a_act = __builtin_popcount(SA[p]), b_act = __builtin_popcount(SB[p]);
        if (a_act >= MQC_C_KMAX || b_act >= MQC_C_KMAX) { PL.why = "too many active orbitals of one spin"; return PL; }
        D.n_clsA = (int32_t)clsA[p].size(); D.n_clsB = (int32_t)clsB[p].size();
        D.clsA_off = PL.cls.size(); PL.cls.insert(PL.cls.end(), clsA[p].begin(), clsA[p].end());
        D.clsB_off = PL.cls.size(); PL.cls.insert(PL.cls.end(), clsB[p].begin(), clsB[p].end());
        int capA = 1, capB = 1;
        bool kA[MQC_C_KMAX] = {}, kB[MQC_C_KMAX] = {};
        for (const MqcCClass& c : clsA[p]) { capA = std::max(capA, (int)c.cnt); kA[c.k] = true; }
        for (const MqcCClass& c : clsB[p]) { capB = std::max(capB, (int)c.cnt); kB[c.k] = true; }
        if (capA > 127 || capB > 127) { PL.why = "more than 127 patterns per class"; return PL; }   // 7-bit pattern ranks
        D.capA = capA; D.capB = capB;
        D.strideB = capB | 1;                                 // odd row stride: bank-conflict-free row pairs
        if (elem == 8) { D.strideB = (capB + 3) & ~1; if (!(D.strideB & 2)) D.strideB += 2; }   // even, >= capB + 2, not a multiple of 4
        if (capA * D.strideB > 16383) { PL.why = "tile larger than 16383 slots"; return PL; }      // 14-bit slots in the pair words
        // maps: layout p position -> layout p + 1 position
        D.map_row_off = PL.maps.size();
        { std::vector<uint32_t> m(PL.nA); bool id = true; for (uint32_t i = 0; i < PL.nA; ++i) { m[posA[p][i]] = posA[p + 1][i]; id = id && posA[p][i] == posA[p + 1][i]; }
          PL.maps.insert(PL.maps.end(), m.begin(), m.end()); D.row_identity = id; }
        D.map_col_off = PL.maps.size();
        { std::vector<uint32_t> m(PL.nB); bool id = true; for (uint32_t i = 0; i < PL.nB; ++i) { m[posB[p][i]] = posB[p + 1][i]; id = id && posB[p][i] == posB[p + 1][i]; }
          PL.maps.insert(PL.maps.end(), m.begin(), m.end()); D.col_identity = id; }
        // tiles, heaviest first
        const u64 nt = (u64)D.n_clsA * (u64)D.n_clsB;
        if (nt > 0xFFFFFFFFull) return PL;
        D.n_tiles = (uint32_t)nt;
        D.tile_off = PL.tiles.size();
        D.tiles_explicit = (nt <= (1ull << 22) && D.n_clsA <= 65535 && D.n_clsB <= 65535) ? 1u : 0u;
        if (D.tiles_explicit) {
            std::vector<std::pair<int, uint32_t>> tmp;
            tmp.reserve(nt);
            for (int i = 0; i < D.n_clsA; ++i) for (int j = 0; j < D.n_clsB; ++j)
                tmp.emplace_back((int)clsA[p][i].cnt * (int)clsB[p][j].cnt, (uint32_t)i | ((uint32_t)j << 16));
            std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, uint32_t>& x, const std::pair<int, uint32_t>& y) { return x.first > y.first; });
            for (auto& kv : tmp) PL.tiles.push_back(kv.second);
        }
        // factors in orbital space
        D.fac_off = PL.fac.size();
        std::vector<uint32_t> mA(D.nf), vA(D.nf), zAi(D.nf), mB(D.nf), vB(D.nf), zBi(D.nf), cA(D.nf), cB(D.nf);
        for (int f = 0; f < D.nf; ++f) {
            const MqcFactor& F = fac[S.tf[P.f0 + f].slot];
            uint32_t za, zb;
            split(F.m, mA[f], mB[f]); split(F.v, vA[f], vB[f]); split(F.z, za, zb); split(F.ctrl, cA[f], cB[f]);
            if ((mA[f] & ~SA[p]) || (mB[f] & ~SB[p]) || (cA[f] & ~SA[p]) || (cB[f] & ~SB[p])) { PL.why = "factor touches an inactive orbital"; return PL; }      // the schedule guarantees this
            zAi[f] = za & SA[p]; zBi[f] = zb & SB[p];
            MqcCFactor C{};
            C.zA_out = za & ~SA[p]; C.zB_out = zb & ~SB[p];
            C.slot = F.slot;
            C.theta_sign = (int32_t)((uint32_t)F.theta | ((uint32_t)(F.sign0 & 1) << 31));
            PL.fac.push_back(C);
        }
        // pattern-pair lists per (side, factor, k), then the class programs
        std::vector<std::vector<uint32_t>> lists[2];             // [side][f * KMAX + k] -> entries lo | hi << 7 | par << 14
        for (int side = 0; side < 2; ++side) {
            const uint32_t act = side ? SB[p] : SA[p];
            const int nact = side ? b_act : a_act;
            const bool* kuse = side ? kB : kA;
            const std::vector<uint32_t>& fm = side ? mB : mA;
            const std::vector<uint32_t>& fv = side ? vB : vA;
            const std::vector<uint32_t>& fz = side ? zBi : zAi;
            const std::vector<uint32_t>& fc = side ? cB : cA;
            std::vector<std::vector<uint32_t>> pats(MQC_C_KMAX);
            {
                uint32_t sub = 0;
                while (true) { pats[__builtin_popcount(sub)].push_back(sub); if (sub == act) break; sub = (sub - act) & act; }
            }
            lists[side].assign((size_t)D.nf * MQC_C_KMAX, std::vector<uint32_t>());
            for (int k = 0; k <= nact; ++k) {
                if (!kuse[k]) continue;
                const std::vector<uint32_t>& pk = pats[k];       // ascending
                auto rank = [&](uint32_t t) { return (uint32_t)(std::lower_bound(pk.begin(), pk.end(), t) - pk.begin()); };
                for (int f = 0; f < D.nf; ++f) {
                    std::vector<uint32_t>& L = lists[side][(size_t)f * MQC_C_KMAX + k];
                    for (uint32_t t : pk) {
                        if ((t & fm[f]) != fv[f] || (t & fc[f]) != fc[f]) continue;      // lower partner; control orbitals occupied
                        L.push_back(rank(t) | (rank(t ^ fm[f]) << 7) | (((uint32_t)__builtin_popcount(t & fz[f]) & 1u) << 14));
                    }
                }
            }
        }
        D.live_tab_off = PL.live_tab.size();
        PL.live_tab.resize(PL.live_tab.size() + (size_t)MQC_C_KMAX * MQC_C_KMAX, MqcCLive{0, 0, 0, 0});
        for (int ka = 0; ka <= a_act; ++ka) for (int kb = 0; kb <= b_act; ++kb) {
            if (!kA[ka] || !kB[kb]) continue;
            if (PL.prog_hdr.size() & 1) PL.prog_hdr.push_back(MqcCProgHdr{0, 0});
            MqcCLive& Lv = PL.live_tab[D.live_tab_off + (size_t)ka * MQC_C_KMAX + kb];
            Lv.off = (uint32_t)PL.prog_hdr.size();
            while (PL.prog.size() & 3) PL.prog.push_back(0);
            Lv.word_base = (uint32_t)PL.prog.size();
            for (int f = 0; f < D.nf; ++f) {
                const std::vector<uint32_t>& LA = lists[0][(size_t)f * MQC_C_KMAX + ka];
                const std::vector<uint32_t>& LB = lists[1][(size_t)f * MQC_C_KMAX + kb];
                if (LA.empty() || LB.empty()) continue;
                if (PL.prog.size() + LA.size() * LB.size() > 0xFFFFFFF0ull) { PL.why = "program pool too large"; return PL; }
                MqcCProgHdr H;
                H.off = (uint32_t)PL.prog.size() - Lv.word_base;
                H.np_f = (uint32_t)(LA.size() * LB.size()) | ((uint32_t)f << 11);
                if (LA.size() * LB.size() > (size_t)MQC_C_MAX_STEP_PAIRS(1024)) { PL.why = "a step has more pairs than the ring holds"; return PL; }
                PL.max_step_pairs = std::max(PL.max_step_pairs, (uint32_t)(LA.size() * LB.size()));
                PL.prog_hdr.push_back(H);
                std::vector<uint32_t> words;
                words.reserve(LA.size() * LB.size());
                for (uint32_t la : LA) for (uint32_t lb : LB) {
                    const uint32_t e0 = (la & 0x7fu) * (uint32_t)D.strideB + (lb & 0x7fu);
                    const uint32_t e1 = ((la >> 7) & 0x7fu) * (uint32_t)D.strideB + ((lb >> 7) & 0x7fu);
                    words.push_back(e0 | (e1 << 14) | ((((la ^ lb) >> 14) & 1u) << 28));
                }
                mqc_c_bank_order(words, elem);
                PL.prog.insert(PL.prog.end(), words.begin(), words.end());
                Lv.cnt++;
            }
            while (PL.prog.size() & 3) PL.prog.push_back(0);
            Lv.n_words = (uint32_t)PL.prog.size() - Lv.word_base;
            D.max_live = std::max(D.max_live, (int)Lv.cnt);
        }
        if (PL.prog_hdr.size() & 1) PL.prog_hdr.push_back(MqcCProgHdr{0, 0});
        PL.pass.push_back(D);
    }
    PL.chunk = PL.max_step_pairs <= (uint32_t)MQC_C_MAX_STEP_PAIRS(256) ? 256 : 1024;
    for (const MqcCLive& Lv : PL.live_tab) if (Lv.cnt) mqc_c_ring_flags(&PL.prog_hdr[Lv.off], (int)Lv.cnt, Lv.n_words, PL.chunk);
    for (const MqcCPassDev& D : PL.pass)
        for (int adj = 0; adj < 2; ++adj) PL.max_smem[adj] = std::max(PL.max_smem[adj], mqc_c_smem_bytes(D, adj != 0, 4, PL.chunk, elem));
    if (PL.cls.empty()) PL.cls.push_back(MqcCClass{});
    if (PL.fac.empty()) PL.fac.push_back(MqcCFactor{});
    if (PL.maps.empty()) PL.maps.push_back(0);
    if (PL.tiles.empty()) PL.tiles.push_back(0);
    if (PL.live_tab.empty()) PL.live_tab.push_back(MqcCLive{0, 0, 0, 0});
    if (PL.prog_hdr.empty()) PL.prog_hdr.assign(2, MqcCProgHdr{0, 0});
    while (PL.prog.size() & 3) PL.prog.push_back(0);
    if (PL.prog.empty()) PL.prog.assign(4, 0);
    PL.ok = true;
    return PL;
}

// Host execution of the plan (tests: k_csweep walks exactly these tables).  Returns the state
// in canonical order, C[ia * nB + ib] as interleaved (re, im).
static inline std::vector<double> mqc_cplan_forward_host(const MqcCPlan& PL, const std::vector<double>& cosv, const std::vector<double>& sinv) {
    const size_t nA = PL.nA, nB = PL.nB;
    std::vector<double> cur(2 * nA * nB, 0.0), nxt(2 * nA * nB, 0.0);
    cur[2 * ((size_t)PL.ref_row * nB + PL.ref_col)] = 1.0;
    std::vector<double> tr, ti;
    for (const MqcCPassDev& D : PL.pass) {
        for (uint32_t it = 0; it < D.n_tiles; ++it) {
            uint32_t ia, ib;
            if (D.tiles_explicit) { const uint32_t t = PL.tiles[D.tile_off + it]; ia = t & 0xffffu; ib = t >> 16; }
            else { ia = it / (uint32_t)D.n_clsB; ib = it - ia * (uint32_t)D.n_clsB; }
            const MqcCClass& cA = PL.cls[D.clsA_off + ia];
            const MqcCClass& cB = PL.cls[D.clsB_off + ib];
            const int nAl = cA.cnt, nBl = cB.cnt, sb = D.strideB;
            tr.assign((size_t)D.capA * sb, 0.0); ti.assign((size_t)D.capA * sb, 0.0);
            for (int ja = 0; ja < nAl; ++ja) for (int jb = 0; jb < nBl; ++jb) {
                const size_t g = (size_t)(cA.row0 + ja) * nB + cB.row0 + jb;
                tr[(size_t)ja * sb + jb] = cur[2 * g]; ti[(size_t)ja * sb + jb] = cur[2 * g + 1];
            }
            const MqcCLive& L = PL.live_tab[D.live_tab_off + (size_t)cA.k * MQC_C_KMAX + cB.k];
            for (uint32_t i = 0; i < L.cnt; ++i) {
                const MqcCProgHdr& H = PL.prog_hdr[L.off + i];
                const int f = (int)MQC_C_F(H.np_f);
                const uint32_t npairs = MQC_C_NP(H.np_f);
                const MqcCFactor& F = PL.fac[D.fac_off + f];
                const unsigned sg = ((unsigned)__builtin_popcount(cA.mask & F.zA_out) + (unsigned)__builtin_popcount(cB.mask & F.zB_out) + ((uint32_t)F.theta_sign >> 31)) & 1u;
                const double c = cosv[F.slot], sn = sg ? -sinv[F.slot] : sinv[F.slot];
                for (uint32_t u = 0; u < npairs; ++u) {
                    const uint32_t w = PL.prog[L.word_base + H.off + u];
                    const int e0 = (int)(w & 0x3fffu), e1 = (int)((w >> 14) & 0x3fffu);
                    const double ssn = ((w >> 28) & 1u) ? -sn : sn;
                    const double ax = tr[e0], ay = ti[e0], bx = tr[e1], by = ti[e1];
                    tr[e0] = c * ax - ssn * bx; ti[e0] = c * ay - ssn * by;
                    tr[e1] = c * bx + ssn * ax; ti[e1] = c * by + ssn * ay;
                }
            }
            for (int ja = 0; ja < nAl; ++ja) for (int jb = 0; jb < nBl; ++jb) {
                const size_t g = (size_t)PL.maps[D.map_row_off + cA.row0 + ja] * nB + PL.maps[D.map_col_off + cB.row0 + jb];
                nxt[2 * g] = tr[(size_t)ja * sb + jb]; nxt[2 * g + 1] = ti[(size_t)ja * sb + jb];
            }
        }
        cur.swap(nxt);
    }
    if (PL.pass.empty()) {
        // no passes: layout 0 is the canonical order already (ref_row / ref_col are canonical positions)
    }
    return cur;
}
