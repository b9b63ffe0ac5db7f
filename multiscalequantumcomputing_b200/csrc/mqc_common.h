// Shared host/device definitions for the B200 UCCSD state-vector path.
//
// Index conventions (see include/mqc_b200.h):
//   * "logical bit" l of an amplitude index is OpenFermion's: qubit q <-> bit N-1-q
//     (/root/reference/test/testlog_vqe:92-94 pins this order).
//   * the state lives in HBM in a *physical* bit order that the sweep scheduler permutes
//     from pass to pass (phys[l] = physical bit position of logical bit l); every kernel
//     works on physical indices with masks translated on the host.
//
// An excitation factor exp(theta (T - T^+)) is a Givens rotation on amplitude pairs:
//   x0 : (x0 & m) == v  (annihilated orbitals occupied, created orbitals empty)
//   x1 = x0 ^ m
//   s  = (-1)^(sign0 + popc(x0 & z))                     (Jordan-Wigner parity)
//   psi'[x0] = c psi[x0] - s sn psi[x1],   psi'[x1] = c psi[x1] + s sn psi[x0]
// with c = cos(theta), sn = sin(theta).  The generator acts as
//   (A psi)[x1] = s psi[x0],   (A psi)[x0] = -s psi[x1].
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MQC_HD __host__ __device__ __forceinline__
#else
#define MQC_HD inline
#endif

typedef unsigned long long u64;

#define MQC_MAX_TILE_BITS 13
#define MQC_MAX_PASS_FACTORS 320

struct MqcFactor {          // one excitation factor in some bit coordinate system
    u64 ctrl;               // control mask: the factor only acts on determinants with these bits set (number-dressed
                            // single n_c (a_a^+ a_i - h.c.)); 0 for plain singles / doubles.  Compact sector path only.
    u64 m;                  // flip mask (2 or 4 bits)
    u64 v;                  // pattern of the "lower" partner on the m bits
    u64 z;                  // parity mask (never overlaps m)
    int32_t npos;           // popc(m)
    int32_t pos[4];         // bit positions of m, ascending
    int32_t sign0;          // constant sign bit
    int32_t slot;           // index of the factor in the user's list (-> cos/sin table)
    int32_t theta;          // theta index (gradient accumulates here)
};

struct MqcTileFactor {      // factor in tile-local coordinates (k <= 16 bits)
    u64 z_out;              // physical parity mask over the tile's fixed (inactive) bits
    uint16_t m, v, z;       // local masks
    uint8_t npos;
    uint8_t sign0;
    uint8_t pos[4];         // local positions of m, ascending (4-byte aligned: read as one word)
    int32_t slot;
    int32_t theta;
};

struct MqcPass {
    int32_t k;              // active (tile) bits
    int32_t f0, f1;         // tile-factor range of this pass
    int32_t has_perm;       // store (forward) / load (reverse) re-layout present
    u64 act_mask;           // physical mask of active bits
    int32_t pos[16];        // physical position of local bit j (ascending)
    uint8_t perm[16];       // forward store: new local bit j' <- old local bit perm[j']
};

// Compressed-sector tile plan (second generation).  With a_act alpha and b_act beta active
// bits, the non-zero amplitudes of a tile are the product set  {alpha-local patterns with ka
// bits} x {beta-local patterns with kb bits}; slot = ja * cntB[kb] + jb with ja / jb the ranks
// of the patterns (ascending local-mask value).  Per pass and per (factor, ka) [resp. kb] the
// host tabulates the pattern pairs the factor couples on that spin species, so the kernel
// enumerates exactly the coupled slot pairs (pair list A x pair list B) instead of scanning.
#define MQC_SP_KMAX 17
struct MqcSpPass {
    uint32_t patoffA[MQC_SP_KMAX], patoffB[MQC_SP_KMAX];   // first pattern of class ka / kb in dep[] / depp[]
    uint16_t cntA[MQC_SP_KMAX], cntB[MQC_SP_KMAX];         // C(a_act, ka), C(b_act, kb)
    uint32_t hdrA_off, hdrB_off;                           // header (factor, k) -> {list offset, count}
    int32_t a_act, b_act;
    u64 amask, bmask;                                      // physical masks of alpha / beta qubits (load layout)
    int32_t na, nb, cap;                                   // sector, max slots of any tile of this pass
    int32_t lcap;                                          // pair-list entries a 32-factor chunk can need
};
// pair list entry: lo pattern rank (12) | hi pattern rank << 12 | in-tile parity << 24

// bank swizzle of the dense tile (see mqc_kernels.cuh): GF(2)-linear image of the high bits
#define MQC_SWZ_M0 0x74E8u
#define MQC_SWZ_M1 0x9D38u
#define MQC_SWZ_M2 0x3A70u
#define MQC_LUT_STRIDE 64       // u16 per factor: A[32] B[8] C[8] swz(m) log2(pairs) pad

// insert zero bits at the (ascending) positions pos[0..npos)
MQC_HD u64 mqc_insert_zeros(u64 u, const int32_t* pos, int npos) {
    for (int i = 0; i < npos; ++i) {
        const u64 low = u & ((1ull << pos[i]) - 1ull);
        u = ((u >> pos[i]) << (pos[i] + 1)) | low;
    }
    return u;
}
MQC_HD uint32_t mqc_insert_zeros8(uint32_t u, const uint8_t* pos, int npos) {
    for (int i = 0; i < npos; ++i) {
        const uint32_t low = u & ((1u << pos[i]) - 1u);
        u = ((u >> pos[i]) << (pos[i] + 1)) | low;
    }
    return u;
}

// deposit the low bits of t into the positions pos[0..k) (generic, slow path / LUT builder)
MQC_HD u64 mqc_deposit(u64 t, const int32_t* pos, int k) {
    u64 out = 0;
    for (int j = 0; j < k; ++j) out |= ((t >> j) & 1ull) << pos[j];
    return out;
}
// deposit tile number tau into the bits NOT in act_mask (ascending)
MQC_HD u64 mqc_deposit_complement(u64 tau, u64 act_mask, int n_bits) {
    u64 out = 0;
    int j = 0;
    for (int b = 0; b < n_bits; ++b) {
        if (!((act_mask >> b) & 1ull)) {
            out |= ((tau >> j) & 1ull) << b;
            ++j;
        }
    }
    return out;
}
MQC_HD uint32_t mqc_permute_bits(uint32_t t, const uint8_t* perm, int k) {
    uint32_t out = 0;
    for (int j = 0; j < k; ++j) out |= ((t >> j) & 1u) << perm[j];
    return out;
}

#if defined(__CUDA_ARCH__)
#define MQC_POPC64(x) __popcll(x)
#else
#define MQC_POPC64(x) __builtin_popcountll(x)
#endif
