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
// Factor order is preserved exactly (excitation factors do not commute in general).
// Pure C++ (no CUDA): testable on a CPU-only box through mqc_schedule_* in the C ABI.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "mqc_common.h"

// One step of a (possibly sharded) sweep: a tile pass on the local shard, or the exchange of
// the amplitudes' global bit (physical position pg >= n_local) with local position pl between
// rank pairs r <-> r ^ (1 << (pg - n_local)).
struct MqcStep {
    int type;      // 0 = tile pass, 1 = exchange
    int a, b;      // pass index, 0   |   pg, pl
};

struct MqcSchedule {
    int n_bits = 0, k = 0, low = 0;
    int n_global = 0;              // top physical positions that live in the rank id
    std::vector<MqcStep> steps;
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
    if (K < N && K < L + 4) throw std::runtime_error("tile bits must be >= low bits + 4");
    S.n_bits = NT; S.k = K; S.low = L;
    std::vector<int> phys(NT), inv(NT);
    for (int l = 0; l < NT; ++l) phys[l] = inv[l] = l;
    const size_t n = fac.size();
    // next use of every logical bit after factor t (for the exchange victim choice)
    auto next_use = [&](int bit, size_t from) -> size_t {
        for (size_t t = from; t < n; ++t) if ((fac[t].m >> bit) & 1ull) return t;
        return n + 1;
    };
    // The reference state is a single determinant, so the initial layout is free: put the
    // bits the first factors move into the low physical positions.
    if (K < N && L > 0) {
        int placed = 0;
        u64 seen = 0;
        for (size_t t = 0; t < n && placed < L; ++t) {
            u64 mm = fac[t].m & ~seen;
            while (mm && placed < L) {
                const int b = __builtin_ctzll(mm); mm &= mm - 1; seen |= 1ull << b;
                const int p_old = phys[b], other = inv[placed];
                phys[b] = placed; inv[placed] = b;
                phys[other] = p_old; inv[p_old] = other;
                ++placed;
            }
        }
    }
    // ... and the bits whose first use is farthest away into the global (rank) positions
    if (n_global > 0) {
        std::vector<int> cand;
        for (int l = 0; l < NT; ++l) if (!(K < N && phys[l] < L)) cand.push_back(l);
        std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return next_use(a, 0) > next_use(b, 0); });
        for (int t = 0; t < n_global && t < (int)cand.size(); ++t) {
            const int b = cand[t], target = N + t;
            if (phys[b] == target) continue;
            const int other = inv[target], p_old = phys[b];
            phys[b] = target; inv[target] = b;
            phys[other] = p_old; inv[p_old] = other;
        }
    }
    S.phys_init = phys;
    size_t i = 0;
    while (i < n) {
        // bring the bits the next factor moves into the local shard
        {
            u64 need = fac[i].m;
            while (need) {
                const int b = __builtin_ctzll(need); need &= need - 1;
                if (phys[b] < N) continue;
                // victim: local bit, not moved by this factor, whose next use is farthest away
                int victim = -1; size_t far = 0;
                for (int p = N - 1; p >= 0; --p) {
                    const int lb = inv[p];
                    if ((fac[i].m >> lb) & 1ull) continue;
                    const size_t nu = next_use(lb, i);
                    if (victim < 0 || nu > far) { victim = lb; far = nu; }
                }
                if (victim < 0) throw std::runtime_error("no local bit to exchange with");
                const int pg = phys[b], pl = phys[victim];
                S.steps.push_back(MqcStep{1, pg, pl});
                phys[b] = pl; inv[pl] = b;
                phys[victim] = pg; inv[pg] = victim;
            }
        }
        u64 A = 0;                                   // logical active set
        for (int p = 0; p < L; ++p) A |= 1ull << inv[p];
        if (K == N) { A = 0; for (int p = 0; p < N; ++p) A |= 1ull << inv[p]; }
        size_t j = i;
        while (j < n && (int)(j - i) < cap) {
            bool local = true;
            { u64 mm = fac[j].m; while (mm) { const int b = __builtin_ctzll(mm); mm &= mm - 1; if (phys[b] >= N) local = false; } }
            if (!local) break;
            const u64 U = A | fac[j].m;
            if (MQC_POPC64(U) > K) break;
            A = U; ++j;
        }
        if (j == i) throw std::runtime_error("factor does not fit a tile");
        // look-ahead: bits the following factors move, in order of first use
        std::vector<int> ahead;
        u64 seen = 0;
        for (size_t t = j; t < n && (int)ahead.size() < N; ++t) {
            u64 mm = fac[t].m & ~seen;
            while (mm) { int b = __builtin_ctzll(mm); mm &= mm - 1; ahead.push_back(b); seen |= 1ull << b; }
            if ((int)ahead.size() >= 2 * K) break;
        }
        // fill A up to K bits: first with bits needed soonest, then lowest physical positions
        for (size_t t = 0; t < ahead.size() && MQC_POPC64(A) < K; ++t) if (phys[ahead[t]] < N) A |= 1ull << ahead[t];
        for (int p = 0; p < N && MQC_POPC64(A) < K; ++p) A |= 1ull << inv[p];

        MqcPass P{};
        P.k = K;
        P.f0 = (int)S.tf.size();
        S.pass_phys_before.push_back(phys);
        std::vector<int> pos;                       // physical positions of A, ascending
        for (int p = 0; p < N; ++p) if ((A >> inv[p]) & 1ull) pos.push_back(p);
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
        // choose the next low set among A: bits needed soonest first
        P.has_perm = 0;
        for (int jj = 0; jj < K; ++jj) P.perm[jj] = (uint8_t)jj;
        if (j < n && K < N && L > 0) {
            std::vector<int> newlow;
            u64 chosen = 0;
            for (size_t t = 0; t < ahead.size() && (int)newlow.size() < L; ++t)
                if (((A >> ahead[t]) & 1ull) && !((chosen >> ahead[t]) & 1ull)) { newlow.push_back(ahead[t]); chosen |= 1ull << ahead[t]; }
            // keep current low bits where possible (identity is cheapest), then anything in A
            for (int p = 0; p < L && (int)newlow.size() < L; ++p)
                if (!((chosen >> inv[p]) & 1ull)) { newlow.push_back(inv[p]); chosen |= 1ull << inv[p]; }
            // new layout inside the active positions: members of newlow that already sit in
            // a low position stay; the others swap with low-position bits that leave.
            std::vector<int> newphys = phys;
            std::vector<int> leaving;                // low positions whose bit is not in newlow
            for (int p = 0; p < L; ++p) if (!((chosen >> inv[p]) & 1ull)) leaving.push_back(p);
            size_t li = 0;
            for (int b : newlow) {
                if (phys[b] >= L) {
                    const int plow = leaving[li++];
                    const int other = inv[plow];
                    newphys[b] = plow;
                    newphys[other] = phys[b];
                }
            }
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
        S.steps.push_back(MqcStep{0, (int)S.passes.size(), 0});
        S.passes.push_back(P);
        S.pass_phys_after.push_back(phys);
        i = j;
    }
    S.phys_final = phys;
    return S;
}
