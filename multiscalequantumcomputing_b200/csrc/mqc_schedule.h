// Host-side sweep scheduler: turns the user's ordered factor list into fused tile passes.
//
// A pass keeps K "active" physical bits of the amplitude index in shared memory (a tile of
// 2^K amplitudes, all other bits fixed per tile) and applies every factor whose moved bits
// (its flip mask) lie inside the active set; Jordan-Wigner parity over the other bits is a
// per-tile constant.  The lowest L physical bits are always active so that every global
// access is a contiguous run of 2^L amplitudes.  Because a tile reads and writes the SAME
// address set, a pass may permute which logical bit sits at which active position while it
// stores ("re-layout") for free; the scheduler uses this to move the bits the NEXT pass
// needs into the low positions, so consecutive passes only have to share L bits.
//
// Sharded state: the top n_global physical positions are the rank id; each
// rank holds 2^(n_bits - n_global) amplitudes.  A global position may be ACTIVE like any
// other: the tile then straddles the 2^ga ranks that differ in the active global bits, its
// CTA reads and writes the peers' shards directly over NVLink, and the same free re-layout
// moves the logical bit that sat in the rank id into local memory while a bit that is not
// needed for the longest time takes its place (Belady's rule).  The global-qubit exchange is
// therefore fused into a compute pass: no separate swap pass, no staging buffer.
//
// Factor order is preserved exactly (excitation factors do not commute in general).
// Pure C++ (no CUDA): testable on a CPU-only box through mqc_schedule_* in the C ABI.
#pragma once
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mqc_common.h"

struct MqcSchedule {
    int n_bits = 0, k = 0, low = 0;
    int n_global = 0;              // top physical positions that live in the rank id
    std::vector<MqcPass> passes;
    std::vector<MqcTileFactor> tf;
    std::vector<std::vector<int>> pass_phys_before, pass_phys_after;   // layout around each pass
    std::vector<int> phys_init;    // logical bit -> physical position before the first pass
    std::vector<int> phys_final;   // logical bit -> physical position after the last pass
};

static inline u64 mqc_map_mask(u64 mask, const std::vector<int>& phys) {
    u64 out = 0;
    for (size_t l = 0; l < phys.size(); ++l)
        if ((mask >> l) & 1ull) out |= 1ull << phys[l];
    return out;
}

// Which tiles of a pass a rank sweeps, and where they start.  The 2^ga ranks that differ
// only in the pass's ACTIVE global bits share one set of 2^(n_local - K + ga) tiles; each
// takes the 2^(n_local - K) of them whose top ga tile-counter bits equal its own bits at the
// active global positions.  Inactive global bits of a tile are the rank's own.
struct MqcTileMap {
    u64 tau_hi;        // OR-ed into the tile counter (blockIdx) before it is deposited
    u64 fixed_bits;    // rank bits at inactive global positions, in physical coordinates
    u64 n_tiles;       // grid size
};
static inline MqcTileMap mqc_tile_map(const MqcPass& P, int n_bits, int n_global, int rank) {
    const int N = n_bits - n_global;
    const u64 gact = P.act_mask >> N;                    // active global positions (rank-bit coordinates)
    int ga = 0;
    u64 share = 0;
    for (int b = 0; b < n_global; ++b)
        if ((gact >> b) & 1ull) { share |= (u64)((rank >> b) & 1) << ga; ++ga; }
    MqcTileMap M;
    const int lg_tiles = N - P.k;                        // per rank, always
    M.n_tiles = 1ull << lg_tiles;
    M.tau_hi = share << lg_tiles;
    M.fixed_bits = ((u64)rank << N) & ~P.act_mask;
    return M;
}

// factors are given in LOGICAL bit coordinates
static inline MqcSchedule mqc_build_schedule(int n_bits, int k_req, int low_req,
                                             const std::vector<MqcFactor>& fac, int cap = MQC_MAX_PASS_FACTORS,
                                             int n_global = 0) {
    MqcSchedule S;
    const int NT = n_bits;                 // all physical bits
    const int N = n_bits - n_global;       // local physical bits: positions 0 .. N-1
    if (N < 1) throw std::runtime_error("too many global bits");
    S.n_global = n_global;
    const int K = std::min(k_req, N);
    const int L = std::min(low_req, K);
    if (K > 16) throw std::runtime_error("tile bits too large");
    if (K < NT && K < L + 4) throw std::runtime_error("tile bits must be >= low bits + 4");
    S.n_bits = NT; S.k = K; S.low = L;
    std::vector<int> phys(NT), inv(NT);
    for (int l = 0; l < NT; ++l) phys[l] = inv[l] = l;
    const size_t n = fac.size();
    // first factor at or after `from` that moves logical bit `bit` (n + 1: never again)
    auto next_use = [&](int bit, size_t from) -> size_t {
        for (size_t t = from; t < n; ++t) if (((fac[t].m | fac[t].ctrl) >> bit) & 1ull) return t;
        return n + 1;
    };
    auto swap_to = [&](int b, int target) {          // put logical bit b at physical position target
        const int p_old = phys[b], other = inv[target];
        phys[b] = target; inv[target] = b;
        phys[other] = p_old; inv[p_old] = other;
    };
    // The reference state is a single determinant, so the initial layout is free: put the
    // bits the first factors move into the low physical positions ...
    if (K < NT && L > 0) {
        int placed = 0;
        u64 seen = 0;
        for (size_t t = 0; t < n && placed < L; ++t) {
            u64 mm = (fac[t].m | fac[t].ctrl) & ~seen;
            while (mm && placed < L) {
                const int b = __builtin_ctzll(mm); mm &= mm - 1; seen |= 1ull << b;
                swap_to(b, placed);
                ++placed;
            }
        }
    }
    // ... and the bits whose first use is farthest away into the global (rank) positions
    if (n_global > 0) {
        std::vector<int> cand;
        for (int l = 0; l < NT; ++l) if (!(K < NT && phys[l] < L)) cand.push_back(l);
        std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return next_use(a, 0) > next_use(b, 0); });
        for (int t = 0; t < n_global && t < (int)cand.size(); ++t) swap_to(cand[t], N + t);
    }
    S.phys_init = phys;
    size_t i = 0;
    while (i < n) {
        u64 A = 0;                                   // logical active set
        for (int p = 0; p < L; ++p) A |= 1ull << inv[p];
        size_t j = i;
        while (j < n && (int)(j - i) < cap) {
            const u64 U = A | fac[j].m | fac[j].ctrl;
            if (MQC_POPC64(U) > K) break;
            A = U; ++j;
        }
        if (j == i) throw std::runtime_error("factor does not fit a tile");
        bool has_global = false;
        for (int p = N; p < NT; ++p) if ((A >> inv[p]) & 1ull) has_global = true;
        // look-ahead: bits the following factors move, in order of first use
        std::vector<int> ahead;
        u64 seen = 0;
        for (size_t t = j; t < n && (int)ahead.size() < NT; ++t) {
            u64 mm = (fac[t].m | fac[t].ctrl) & ~seen;
            while (mm) { int b = __builtin_ctzll(mm); mm &= mm - 1; ahead.push_back(b); seen |= 1ull << b; }
            if ((int)ahead.size() >= 2 * K) break;
        }
        // fill A up to K bits: first with bits needed soonest, then lowest physical positions.
        // A bit that sits in the rank id is only prefetched when the tile straddles ranks anyway.
        for (size_t t = 0; t < ahead.size() && MQC_POPC64(A) < K; ++t)
            if (phys[ahead[t]] < N || has_global) A |= 1ull << ahead[t];
        for (int p = 0; p < N && MQC_POPC64(A) < K; ++p) A |= 1ull << inv[p];

        MqcPass P{};
        P.k = K;
        P.f0 = (int)S.tf.size();
        S.pass_phys_before.push_back(phys);
        std::vector<int> pos;                       // physical positions of A, ascending
        for (int p = 0; p < NT; ++p) if ((A >> inv[p]) & 1ull) pos.push_back(p);
        u64 act = 0;
        std::vector<int> local_of_phys(NT, -1);
        for (int jj = 0; jj < K; ++jj) { P.pos[jj] = pos[jj]; act |= 1ull << pos[jj]; local_of_phys[pos[jj]] = jj; }
        P.act_mask = act;
        for (size_t t = i; t < j; ++t) {
            const MqcFactor& f = fac[t];
            MqcTileFactor q{};
            u64 zphys = mqc_map_mask(f.z, phys);
            q.z_out = zphys & ~act;
            u64 zin = zphys & act, mphys = mqc_map_mask(f.m, phys), vphys = mqc_map_mask(f.v, phys);
            uint32_t ml = 0, vl = 0, zl = 0;
            for (int jj = 0; jj < K; ++jj) {
                if ((mphys >> pos[jj]) & 1ull) ml |= 1u << jj;
                if ((vphys >> pos[jj]) & 1ull) vl |= 1u << jj;
                if ((zin >> pos[jj]) & 1ull) zl |= 1u << jj;
            }
            q.m = (uint16_t)ml; q.v = (uint16_t)vl; q.z = (uint16_t)zl;
            q.npos = (uint8_t)f.npos;
            int c = 0;
            for (int jj = 0; jj < K; ++jj) if ((ml >> jj) & 1u) q.pos[c++] = (uint8_t)jj;
            q.sign0 = (uint8_t)(f.sign0 & 1);
            q.slot = f.slot; q.theta = f.theta;
            S.tf.push_back(q);
        }
        P.f1 = (int)S.tf.size();
        // Re-layout among the active positions.  Active global positions receive the active
        // bits whose next use is farthest away; the low positions the bits needed soonest;
        // every other bit stays where it is whenever its position is still free.
        P.has_perm = 0;
        for (int jj = 0; jj < K; ++jj) P.perm[jj] = (uint8_t)jj;
        if (j < n && K < NT) {
            std::vector<int> bitsA;                  // logical bits of A
            for (int jj = 0; jj < K; ++jj) bitsA.push_back(inv[pos[jj]]);
            std::vector<int> gpos;                   // active global positions
            for (int jj = 0; jj < K; ++jj) if (pos[jj] >= N) gpos.push_back(pos[jj]);
            u64 chosen_g = 0, chosen_low = 0;
            std::vector<int> glist, lowlist;
            if (!gpos.empty()) {
                std::vector<int> order = bitsA;
                std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                    const size_t ua = next_use(a, j), ub = next_use(b, j);
                    if (ua != ub) return ua > ub;
                    return (phys[a] >= N) > (phys[b] >= N);          // ties: stay in the rank id
                });
                for (size_t t = 0; t < gpos.size(); ++t) { glist.push_back(order[t]); chosen_g |= 1ull << order[t]; }
            }
            if (L > 0) {
                for (size_t t = 0; t < ahead.size() && (int)lowlist.size() < L; ++t) {
                    const int b = ahead[t];
                    if (((A >> b) & 1ull) && !((chosen_g >> b) & 1ull) && !((chosen_low >> b) & 1ull)) {
                        lowlist.push_back(b); chosen_low |= 1ull << b;
                    }
                }
                // keep current low bits where possible (identity is cheapest), then anything in A
                for (int p = 0; p < L && (int)lowlist.size() < L; ++p) {
                    const int b = inv[p];
                    if (!((chosen_g >> b) & 1ull) && !((chosen_low >> b) & 1ull)) { lowlist.push_back(b); chosen_low |= 1ull << b; }
                }
                for (int b : bitsA) {
                    if ((int)lowlist.size() >= L) break;
                    if (!((chosen_g >> b) & 1ull) && !((chosen_low >> b) & 1ull)) { lowlist.push_back(b); chosen_low |= 1ull << b; }
                }
            }
            std::vector<int> newphys = phys;
            std::vector<char> taken(NT, 0);          // positions of pos[] already assigned
            std::vector<char> done(NT, 0);           // logical bits already assigned
            auto assign_class = [&](const std::vector<int>& bits, const std::vector<int>& places) {
                // bits already sitting in one of `places` stay; the others take the free ones
                std::vector<char> is_place(NT, 0);
                for (int p : places) is_place[p] = 1;
                for (int b : bits) if (is_place[phys[b]]) { newphys[b] = phys[b]; taken[phys[b]] = 1; done[b] = 1; }
                size_t pi = 0;
                for (int b : bits) {
                    if (done[b]) continue;
                    while (pi < places.size() && taken[places[pi]]) ++pi;
                    newphys[b] = places[pi]; taken[places[pi]] = 1; done[b] = 1;
                }
            };
            std::vector<int> lowpos;
            for (int p = 0; p < L; ++p) lowpos.push_back(p);
            assign_class(glist, gpos);
            assign_class(lowlist, lowpos);
            std::vector<int> rest_bits, rest_pos;
            for (int b : bitsA) if (!done[b]) rest_bits.push_back(b);
            for (int jj = 0; jj < K; ++jj) if (!taken[pos[jj]]) rest_pos.push_back(pos[jj]);
            assign_class(rest_bits, rest_pos);

            bool changed = false;
            for (int l = 0; l < NT; ++l) if (newphys[l] != phys[l]) changed = true;
            if (changed) {
                P.has_perm = 1;
                // new local bit j' (physical pos[j']) holds logical l' = newinv[pos[j']],
                // which sat at old local position local_of_phys[phys[l']]
                std::vector<int> newinv(NT);
                for (int l = 0; l < NT; ++l) newinv[newphys[l]] = l;
                for (int jj = 0; jj < K; ++jj) {
                    const int lg = newinv[pos[jj]];
                    P.perm[jj] = (uint8_t)local_of_phys[phys[lg]];
                }
                phys = newphys;
                for (int l = 0; l < NT; ++l) inv[phys[l]] = l;
            }
        }
        S.passes.push_back(P);
        S.pass_phys_after.push_back(phys);
        i = j;
    }
    S.phys_final = phys;
    return S;
}

// ---------------------------------------------------------------------------------------------
// compressed-sector tile plan (see MqcSpPass in mqc_common.h); needs every factor to conserve
// (N_alpha, N_beta).  Qubit q (alpha: even, beta: odd) <-> logical bit n_bits - 1 - q.
// ---------------------------------------------------------------------------------------------
struct MqcSparsePlan {
    std::vector<MqcSpPass> pass;
    std::vector<u64> dep, depp;          // physical offset of a pattern in the load / the other layout
    std::vector<uint16_t> pat;           // the pattern itself (tile-local mask), same indexing
    std::vector<uint32_t> lists;
    std::vector<uint32_t> hdr;           // quadruples {offset, count, float bits of 1 / count, 0}
    bool ok = false;
};

static inline MqcSparsePlan mqc_build_sparse_plan(const MqcSchedule& S, int na, int nb) {
    MqcSparsePlan PL;
    const int n = S.n_bits, K = S.k;
    const int n_orb_a = (n + 1) / 2, n_orb_b = n / 2;
    for (size_t p = 0; p < S.passes.size(); ++p) {
        const MqcPass& P = S.passes[p];
        const std::vector<int>& ph = S.pass_phys_before[p];
        MqcSpPass SP{};
        for (int q = 0; q < n; ++q) {
            const u64 bit = 1ull << ph[n - 1 - q];
            if (q & 1) SP.bmask |= bit; else SP.amask |= bit;
        }
        uint32_t loc[2] = {0, 0};                       // local masks of alpha / beta active bits
        for (int j = 0; j < K; ++j) {
            if ((SP.amask >> P.pos[j]) & 1ull) loc[0] |= 1u << j;
            if ((SP.bmask >> P.pos[j]) & 1ull) loc[1] |= 1u << j;
        }
        SP.a_act = __builtin_popcount(loc[0]); SP.b_act = __builtin_popcount(loc[1]);
        SP.na = na; SP.nb = nb;
        if (SP.a_act >= MQC_SP_KMAX || SP.b_act >= MQC_SP_KMAX) return PL;
        // other-layout position of local bit j: new local bit j' takes old local bit perm[j']
        int newpos[16];
        for (int j = 0; j < K; ++j) newpos[j] = P.pos[j];
        if (P.has_perm) for (int jp = 0; jp < K; ++jp) newpos[P.perm[jp]] = P.pos[jp];
        std::vector<uint32_t> pats[2][MQC_SP_KMAX];     // [spin][k] -> local masks, ascending
        std::vector<int> rank_of[2];
        for (int sp = 0; sp < 2; ++sp) {
            const int act = sp ? SP.b_act : SP.a_act;
            rank_of[sp].assign((size_t)1 << K, -1);
            // enumerate subsets of loc[sp] in ascending order of the mask value
            uint32_t sub = 0;
            while (true) {
                const int k = __builtin_popcount(sub);
                rank_of[sp][sub] = (int)pats[sp][k].size();
                pats[sp][k].push_back(sub);
                if (sub == loc[sp]) break;
                sub = (sub - loc[sp]) & loc[sp];
            }
            for (int k = 0; k <= act; ++k) {
                if (pats[sp][k].size() > 127) return PL;         // pair-list entries hold 7-bit pattern ranks
                (sp ? SP.patoffB : SP.patoffA)[k] = (uint32_t)PL.dep.size();
                (sp ? SP.cntB : SP.cntA)[k] = (uint16_t)pats[sp][k].size();
                for (uint32_t t : pats[sp][k]) {
                    u64 d = 0, dp = 0;
                    for (int j = 0; j < K; ++j) if ((t >> j) & 1u) { d |= 1ull << P.pos[j]; dp |= 1ull << newpos[j]; }
                    PL.dep.push_back(d); PL.depp.push_back(dp); PL.pat.push_back((uint16_t)t);
                }
            }
        }
        // feasible classes and the slot capacity
        double best = 1;
        {
            int ca = 0, cb = 0;
            for (int k = 0; k <= SP.a_act; ++k) if (k <= na && na - k <= n_orb_a - SP.a_act) ca = std::max(ca, (int)SP.cntA[k]);
            for (int k = 0; k <= SP.b_act; ++k) if (k <= nb && nb - k <= n_orb_b - SP.b_act) cb = std::max(cb, (int)SP.cntB[k]);
            best = (double)std::max(1, ca) * std::max(1, cb);
        }
        SP.cap = (int)best;
        const int nf = P.f1 - P.f0;
        int max_act_list[2] = {0, 0};
        for (int sp = 0; sp < 2; ++sp) {
            (sp ? SP.hdrB_off : SP.hdrA_off) = (uint32_t)(PL.hdr.size() / 4);
            const int act = sp ? SP.b_act : SP.a_act;
            for (int f = 0; f < nf; ++f) {
                const MqcTileFactor& F = S.tf[P.f0 + f];
                const uint32_t fm = F.m & loc[sp], fv = F.v & loc[sp], fz = F.z & loc[sp];
                for (int k = 0; k < MQC_SP_KMAX; ++k) {
                    const uint32_t off = (uint32_t)PL.lists.size();
                    uint32_t cnt = 0;
                    if (k <= act) for (uint32_t t : pats[sp][k]) {
                        if ((t & fm) != fv) continue;
                        const uint32_t lo = (uint32_t)rank_of[sp][t], hi = (uint32_t)rank_of[sp][t ^ fm];
                        const uint32_t par = (uint32_t)__builtin_popcount(t & fz) & 1u;
                        PL.lists.push_back(lo | (hi << 12) | (par << 24));
                        ++cnt;
                    }
                    const float rcp = cnt ? 1.0f / (float)cnt : 0.0f;
                    uint32_t rb;
                    std::memcpy(&rb, &rcp, sizeof(rb));
                    // a factor that does not act on this spin couples every pattern with itself: the
                    // kernel then needs no list, only the parity mask
                    const uint32_t ident = (fm == 0) ? 1u : 0u;
                    if (!ident) max_act_list[sp] = std::max(max_act_list[sp], (int)cnt);
                    PL.hdr.push_back(off); PL.hdr.push_back(cnt); PL.hdr.push_back(rb); PL.hdr.push_back(fz | (ident << 16));
                }
            }
        }
        SP.lcap = 32 * (max_act_list[0] + max_act_list[1]) + 32;
        PL.pass.push_back(SP);
    }
    if (PL.pat.empty()) PL.pat.push_back(0);
    if (PL.lists.empty()) PL.lists.push_back(0);
    if (PL.hdr.empty()) PL.hdr.assign(4, 0);
    if (PL.dep.empty()) { PL.dep.push_back(0); PL.depp.push_back(0); }
    PL.ok = true;
    return PL;
}

// Host execution of the plan on a dense 2^n_bits vector in physical order (tests: the kernel
// k_tile_sp2 walks exactly these tables).  cs[slot] = (cos, sin).
static inline void mqc_sparse_plan_forward_host(const MqcSchedule& S, const MqcSparsePlan& PL, std::vector<double>& re,
                                                std::vector<double>& im, const std::vector<double>& cosv,
                                                const std::vector<double>& sinv) {
    const int n = S.n_bits, K = S.k;
    for (size_t p = 0; p < S.passes.size(); ++p) {
        const MqcPass& P = S.passes[p];
        const MqcSpPass& SP = PL.pass[p];
        const int nf = P.f1 - P.f0;
        const u64 n_tiles = 1ull << (n - K);
        std::vector<double> tr, ti;
        for (u64 tile = 0; tile < n_tiles; ++tile) {
            const u64 base = mqc_deposit_complement(tile, P.act_mask, n);
            const int ka = SP.na - __builtin_popcountll(base & SP.amask), kb = SP.nb - __builtin_popcountll(base & SP.bmask);
            if (ka < 0 || kb < 0 || ka > SP.a_act || kb > SP.b_act) continue;
            const int nAl = SP.cntA[ka], nBl = SP.cntB[kb], cnt = nAl * nBl;
            tr.assign(cnt, 0.0); ti.assign(cnt, 0.0);
            for (int e = 0; e < cnt; ++e) {
                const u64 g = base | PL.dep[SP.patoffA[ka] + e / nBl] | PL.dep[SP.patoffB[kb] + e % nBl];
                tr[e] = re[g]; ti[e] = im[g];
            }
            for (int f = 0; f < nf; ++f) {
                const MqcTileFactor& F = S.tf[P.f0 + f];
                const uint32_t* hA = &PL.hdr[4 * (SP.hdrA_off + (size_t)f * MQC_SP_KMAX + ka)];
                const uint32_t* hB = &PL.hdr[4 * (SP.hdrB_off + (size_t)f * MQC_SP_KMAX + kb)];
                const unsigned sg = ((unsigned)__builtin_popcountll(base & F.z_out) + F.sign0) & 1u;
                const double c = cosv[F.slot], sn = sinv[F.slot];
                // mirrors k_tile_sp2: identity sides come from the pattern masks, acting sides from
                // the 7-bit packed pair lists
                const bool idA = (hA[3] >> 16) & 1u, idB = (hB[3] >> 16) & 1u;
                for (uint32_t ua = 0; ua < hA[1]; ++ua) for (uint32_t ub = 0; ub < hB[1]; ++ub) {
                    auto pack = [](uint32_t en) { return (en & 0x7fu) | (((en >> 12) & 0x7fu) << 7) | (((en >> 24) & 1u) << 14); };
                    const uint32_t la = idA ? (ua * 129u | (((uint32_t)__builtin_popcount(PL.pat[SP.patoffA[ka] + ua] & hA[3] & 0xffffu) & 1u) << 14)) : pack(PL.lists[hA[0] + ua]);
                    const uint32_t lb = idB ? (ub * 129u | (((uint32_t)__builtin_popcount(PL.pat[SP.patoffB[kb] + ub] & hB[3] & 0xffffu) & 1u) << 14)) : pack(PL.lists[hB[0] + ub]);
                    const int e0 = (int)(la & 0x7fu) * nBl + (int)(lb & 0x7fu), e1 = (int)((la >> 7) & 0x7fu) * nBl + (int)((lb >> 7) & 0x7fu);
                    const double ssn = ((((la ^ lb) >> 14) ^ sg) & 1u) ? -sn : sn;
                    const double ax = tr[e0], ay = ti[e0], bx = tr[e1], by = ti[e1];
                    tr[e0] = c * ax - ssn * bx; ti[e0] = c * ay - ssn * by;
                    tr[e1] = c * bx + ssn * ax; ti[e1] = c * by + ssn * ay;
                }
            }
            if (P.has_perm) for (int e = 0; e < cnt; ++e) {
                const u64 g = base | PL.dep[SP.patoffA[ka] + e / nBl] | PL.dep[SP.patoffB[kb] + e % nBl];
                re[g] = 0.0; im[g] = 0.0;
            }
            const std::vector<u64>& ds = P.has_perm ? PL.depp : PL.dep;
            for (int e = 0; e < cnt; ++e) {
                const u64 g = base | ds[SP.patoffA[ka] + e / nBl] | ds[SP.patoffB[kb] + e % nBl];
                re[g] = tr[e]; im[g] = ti[e];
            }
        }
    }
}

// Non-empty tiles of every pass for one rank, heaviest first, packed base | ka << 48 | kb << 56.
// off[p] .. off[p + 1] delimit pass p.  Empty result (ok = false) when the lists would be too
// large to be worth it: the kernel then decodes and tests every tile counter itself.
struct MqcTileLists {
    std::vector<u64> tiles;
    std::vector<u64> off;
    bool ok = false;
};
static inline MqcTileLists mqc_build_tile_lists(const MqcSchedule& S, const MqcSparsePlan& PL, int rank, u64 max_entries = (8ull << 20)) {
    MqcTileLists TL;
    if (!PL.ok) return TL;
    const int n_local = S.n_bits - S.n_global;
    u64 total = 0;
    for (const MqcPass& P : S.passes) total += 1ull << (n_local - P.k);
    if (total > max_entries || S.n_bits > 47) return TL;
    TL.off.push_back(0);
    std::vector<std::pair<int, u64>> tmp;
    for (size_t p = 0; p < S.passes.size(); ++p) {
        const MqcPass& P = S.passes[p];
        const MqcSpPass& SP = PL.pass[p];
        const MqcTileMap TM = mqc_tile_map(P, S.n_bits, S.n_global, rank);
        tmp.clear();
        for (u64 t = 0; t < TM.n_tiles; ++t) {
            const u64 base = mqc_deposit_complement(t | TM.tau_hi, P.act_mask, n_local) | TM.fixed_bits;
            const int ka = SP.na - __builtin_popcountll(base & SP.amask), kb = SP.nb - __builtin_popcountll(base & SP.bmask);
            if (ka < 0 || kb < 0 || ka > SP.a_act || kb > SP.b_act) continue;
            tmp.emplace_back((int)SP.cntA[ka] * (int)SP.cntB[kb], base | ((u64)ka << 48) | ((u64)kb << 56));
        }
        std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, u64>& a, const std::pair<int, u64>& b) { return a.first > b.first; });
        for (auto& kv : tmp) TL.tiles.push_back(kv.second);
        TL.off.push_back((u64)TL.tiles.size());
    }
    if (TL.tiles.empty()) TL.tiles.push_back(0);
    TL.ok = true;
    return TL;
}

// ---------------------------------------------------------------------------------------------
// pair-address tables of the dense tile kernel (k_tile_sweep), one 64-entry row per tile factor
// ---------------------------------------------------------------------------------------------
static inline uint32_t mqc_swz_host(uint32_t t) {
    return t ^ ((uint32_t)__builtin_popcount(t & MQC_SWZ_M0) & 1u) ^ (((uint32_t)__builtin_popcount(t & MQC_SWZ_M1) & 1u) << 1) ^
           (((uint32_t)__builtin_popcount(t & MQC_SWZ_M2) & 1u) << 2);
}
static inline std::vector<uint16_t> mqc_build_dense_luts(const MqcSchedule& S) {
    std::vector<uint16_t> L(std::max<size_t>(1, S.tf.size()) * MQC_LUT_STRIDE, 0);
    const uint32_t T = 1u << S.k;
    auto ins0 = [](uint32_t x, uint32_t p) { return ((x >> p) << (p + 1)) | (x & ((1u << p) - 1u)); };
    for (size_t i = 0; i < S.tf.size(); ++i) {
        const MqcTileFactor& F = S.tf[i];
        uint16_t* row = &L[i * MQC_LUT_STRIDE];
        for (int e = 0; e < 48; ++e) {
            const uint32_t up = (e < 32) ? (uint32_t)e : (e < 40 ? (uint32_t)(e - 32) << 5 : (uint32_t)(e - 40) << 8);
            uint32_t xl = up;
            for (int k = 0; k < F.npos; ++k) xl = ins0(xl, F.pos[k]);
            uint32_t val = 0;
            if (xl < T) val = mqc_swz_host(xl) | (((uint32_t)__builtin_popcount(xl & F.z) & 1u) << 15);
            if (e < 32) val ^= mqc_swz_host(F.v) | (((uint32_t)__builtin_popcount(F.v & F.z) & 1u) << 15);   // v rides on table A
            row[e] = (uint16_t)val;
        }
        row[48] = (uint16_t)mqc_swz_host(F.m);
        row[49] = (uint16_t)(S.k - F.npos);
    }
    return L;
}
