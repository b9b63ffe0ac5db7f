// K1 / K3 on the COMPACT sector matrix C[alpha string][beta string] (plan: mqc_cplan.h).
//
//   k_csweep<false>  forward pass:  tile rows arrive by bulk (TMA) copies global -> shared,
//                    completion on an mbarrier; factors; scatter through the row / column map
//                    into the layout of the next pass.
//   k_csweep<true>   adjoint pass (reverse order, -theta) on psi and lambda together with the
//                    gradient partials: gather through the maps (cp.async), factors, rows
//                    leave by bulk copies shared -> global.
//
// One CTA owns a tile (one CTA per tile, or persistent over the pass's tile list, heaviest first).  The
// tile's class program (mqc_cplan.h) tells every thread which pair of slots to rotate for each live
// factor: the program headers arrive with the same mbarrier transaction as the rows, the 32-bit pair
// words stream through a shared-memory ring of four chunks that thread 0 refills with further bulk
// copies, so that between two barriers a thread only reads its two amplitudes, rotates and writes them
// back.  Amplitudes are complex (16 bytes) or, template parameter REAL, real (8 bytes: real Hamiltonian,
// DESIGN.md section 2).  Algorithmic bytes per launch: 2 * b * N_det (forward), 4 * b * N_det (adjoint),
// b = bytes per amplitude - the state is read once and written once, nothing else is touched.
#pragma once
#include <cuda_runtime.h>

#include "mqc_common.h"
#include "mqc_cplan.h"

struct MqcCArgs {
    const MqcCPassDev* pass;
    const MqcCClass* cls;
    const MqcCFactor* fac;
    const uint32_t* maps;
    const uint32_t* tiles;
    const MqcCLive* live_tab;
    const MqcCProgHdr* prog_hdr;
    const uint32_t* prog;
    const double2* cs;
    const void* psi_in;              // double2 (complex) or double (REAL) amplitudes
    void* psi_out;
    const void* lam_in;
    void* lam_out;
    double* grad;
    double* gpart;                   // [gridDim.x][nf] per-CTA gradient partials of this pass (summed in a fixed order
                                     // by k_grad_combine_parts: reproducible gradients); NULL: atomics into grad
    uint32_t nB;
};

__device__ __forceinline__ uint32_t mqc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mqc_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mqc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mqc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mqc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mqc_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = mqc_smem_u32(bar);
    uint32_t done = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
// bulk copy global -> shared (TMA engine, SASS UBLKCP), bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void mqc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(mqc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(mqc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mqc_bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(mqc_smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mqc_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mqc_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mqc_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mqc_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mqc_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mqc_cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(mqc_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mqc_cp_async_amp(double2* d, const double2* g) { mqc_cp_async16(d, g); }
__device__ __forceinline__ void mqc_cp_async_amp(double* d, const double* g) { mqc_cp_async8(d, g); }
// no-op unless the kernel was launched with cudaLaunchAttributeProgrammaticStreamSerialization
__device__ __forceinline__ void mqc_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mqc_cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// one pair rotation on psi (and lambda, with the gradient partial): w = slot0 | slot1 << 14 | parity << 28
template <bool ADJ>
__device__ __forceinline__ double mqc_rotate_pair(double2* __restrict__ sp, double2* __restrict__ sl, const uint32_t w,
                                                  const double c, const double s) {
    const int e0 = (int)(w & 0x3fffu), e1 = (int)((w >> 14) & 0x3fffu);
    const int sgn = (int)((w << 3) & 0x80000000u);                            // pair parity -> sign bit of a double
    const double ssn = __hiloint2double(__double2hiint(s) ^ sgn, __double2loint(s));
    const double2 a = sp[e0], b = sp[e1];
    double g = 0.0;
    if (ADJ) {
        const double2 xa = sl[e0], xb = sl[e1];
        const double gv = xb.x * a.x + xb.y * a.y - xa.x * b.x - xa.y * b.y;
        g = __hiloint2double(__double2hiint(gv) ^ sgn, __double2loint(gv));    // pair parity alone (the tile sign is applied at the flush)
        sl[e0] = make_double2(c * xa.x - ssn * xb.x, c * xa.y - ssn * xb.y);
        sl[e1] = make_double2(c * xb.x + ssn * xa.x, c * xb.y + ssn * xa.y);
    }
    sp[e0] = make_double2(c * a.x - ssn * b.x, c * a.y - ssn * b.y);
    sp[e1] = make_double2(c * b.x + ssn * a.x, c * b.y + ssn * a.y);
    return g;
}

// real amplitudes: the same rotation on 8-byte slots (half the shared-memory traffic and arithmetic)
template <bool ADJ>
__device__ __forceinline__ double mqc_rotate_pair(double* __restrict__ sp, double* __restrict__ sl, const uint32_t w,
                                                  const double c, const double s) {
    const int e0 = (int)(w & 0x3fffu), e1 = (int)((w >> 14) & 0x3fffu);
    const int sgn = (int)((w << 3) & 0x80000000u);
    const double ssn = __hiloint2double(__double2hiint(s) ^ sgn, __double2loint(s));
    const double a = sp[e0], b = sp[e1];
    double g = 0.0;
    if (ADJ) {
        const double xa = sl[e0], xb = sl[e1];
        const double gv = xb * a - xa * b;
        g = __hiloint2double(__double2hiint(gv) ^ sgn, __double2loint(gv));
        sl[e0] = c * xa - ssn * xb;
        sl[e1] = c * xb + ssn * xa;
    }
    sp[e0] = c * a - ssn * b;
    sp[e1] = c * b + ssn * a;
    return g;
}

// two independent pairs of one step at once (the pairs of a step are disjoint): all loads are issued before
// the first store, so the two rotations overlap instead of running back to back
template <bool ADJ>
__device__ __forceinline__ double mqc_rotate_pair2(double* __restrict__ sp, double* __restrict__ sl, const uint32_t w0, const uint32_t w1,
                                                   const double c, const double s) {
    const int a0 = (int)(w0 & 0x3fffu), a1 = (int)((w0 >> 14) & 0x3fffu);
    const int b0 = (int)(w1 & 0x3fffu), b1 = (int)((w1 >> 14) & 0x3fffu);
    const int sg0 = (int)((w0 << 3) & 0x80000000u), sg1 = (int)((w1 << 3) & 0x80000000u);
    const double s0 = __hiloint2double(__double2hiint(s) ^ sg0, __double2loint(s));
    const double s1 = __hiloint2double(__double2hiint(s) ^ sg1, __double2loint(s));
    const double pa0 = sp[a0], pa1 = sp[a1], pb0 = sp[b0], pb1 = sp[b1];
    double g = 0.0;
    if (ADJ) {
        const double xa0 = sl[a0], xa1 = sl[a1], xb0 = sl[b0], xb1 = sl[b1];
        const double g0 = xa1 * pa0 - xa0 * pa1, g1 = xb1 * pb0 - xb0 * pb1;
        g = __hiloint2double(__double2hiint(g0) ^ sg0, __double2loint(g0)) + __hiloint2double(__double2hiint(g1) ^ sg1, __double2loint(g1));
        sl[a0] = c * xa0 - s0 * xa1; sl[a1] = c * xa1 + s0 * xa0;
        sl[b0] = c * xb0 - s1 * xb1; sl[b1] = c * xb1 + s1 * xb0;
    }
    sp[a0] = c * pa0 - s0 * pa1; sp[a1] = c * pa1 + s0 * pa0;
    sp[b0] = c * pb0 - s1 * pb1; sp[b1] = c * pb1 + s1 * pb0;
    return g;
}

// transposed butterfly: eight per-lane partials -> one total per lane group (9 double shuffles instead of 40);
// lane L with (L & 3) == 0 ends up with the total of step j = 4 * bit4(L) + 2 * bit3(L) + bit2(L) in OUT
#define MQC_CS_BUTTERFLY(ACC, LANE, OUT) \
    double OUT; \
    { \
        double w4_[4], w2_[2]; \
        { const bool hi_ = ((LANE) & 16) != 0; \
          _Pragma("unroll") for (int j_ = 0; j_ < 4; ++j_) { \
              const double send_ = hi_ ? ACC[j_] : ACC[j_ + 4], keep_ = hi_ ? ACC[j_ + 4] : ACC[j_]; \
              w4_[j_] = keep_ + __shfl_xor_sync(0xffffffffu, send_, 16); } } \
        { const bool hi_ = ((LANE) & 8) != 0; \
          _Pragma("unroll") for (int j_ = 0; j_ < 2; ++j_) { \
              const double send_ = hi_ ? w4_[j_] : w4_[j_ + 2], keep_ = hi_ ? w4_[j_ + 2] : w4_[j_]; \
              w2_[j_] = keep_ + __shfl_xor_sync(0xffffffffu, send_, 8); } } \
        { const bool hi_ = ((LANE) & 4) != 0; \
          const double send_ = hi_ ? w2_[0] : w2_[1], keep_ = hi_ ? w2_[1] : w2_[0]; \
          OUT = keep_ + __shfl_xor_sync(0xffffffffu, send_, 4); } \
        OUT += __shfl_xor_sync(0xffffffffu, OUT, 2); \
        OUT += __shfl_xor_sync(0xffffffffu, OUT, 1); \
    }

#ifndef MQC_CS_FWD_UNROLL
#define MQC_CS_FWD_UNROLL 2
#endif
template <bool REAL> struct MqcAmp { typedef double2 T; };
template <> struct MqcAmp<true> { typedef double T; };

// REAL: 8-byte amplitudes.  Needs an even row pitch nB (host-checked): a tile row then starts on a 16-byte
// boundary or one element after it (`shift`), the same for all rows of the tile; bulk copies move the
// 16-byte units that cover the row (forward: the neighbours' elements that come along are never addressed;
// adjoint stores: the aligned interior goes by bulk copy, up to two edge elements by plain stores).
template <bool ADJ, int NT, int CH, bool REAL = false>
__global__ void __launch_bounds__(NT, (1024 / NT) > (ADJ ? 8 : 16) ? (ADJ ? 8 : 16) : (1024 / NT))   // adjoint tiles: shared memory allows <= 8 CTAs per SM anyway
k_csweep(const __grid_constant__ MqcCArgs A) {
    typedef typename MqcAmp<REAL>::T AT;
    constexpr int ES = (int)sizeof(AT);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const MqcCPassDev* __restrict__ P = A.pass;
    const int nf = P->nf, strideB = P->strideB, capA = P->capA, capB = P->capB;
    const uint32_t nB = A.nB;
    // ---- shared memory carve-up (mqc_c_smem_bytes) -------------------------------------------
    unsigned char* q = smem_raw;
    const size_t tile_bytes = ((size_t)capA * strideB * ES + 15) & ~(size_t)15;
    AT* sp0 = reinterpret_cast<AT*>(q);                       q += tile_bytes;
    AT* sl0 = reinterpret_cast<AT*>(q);                       if (ADJ) q += tile_bytes;
    uint2* shdr = reinterpret_cast<uint2*>(q) + 8;            q += (((size_t)P->max_live + 9) * 8 + 15) & ~(size_t)15;   // [-8 .. -1] and [n_live]: empty sentinel steps
    uint32_t* ring = reinterpret_cast<uint32_t*>(q);          q += (size_t)MQC_C_RING_SLOTS * CH * 4;
    double2* hcs = reinterpret_cast<double2*>(q);             q += (size_t)nf * 16;
    unsigned char* hsg = q;                                   q += ((size_t)nf + 15) & ~(size_t)15;
    uint32_t* smap = reinterpret_cast<uint32_t*>(q);          q += ((size_t)(capA + capB) * 4 + 15) & ~(size_t)15;
    const size_t gw_stride = (((size_t)nf * 8 + 15) & ~(size_t)15) / 8;
    double* gw = reinterpret_cast<double*>(q);                if (ADJ) q += (size_t)NW * gw_stride * 8;
    uint64_t* bar = reinterpret_cast<uint64_t*>(q);           // [0] tile rows + headers, [1 ..] ring slots
    uint64_t* rbar = bar + 1;
    constexpr uint32_t RMASK = MQC_C_RING_SLOTS * CH - 1;

    // Programmatic dependent launch (the host launches the passes of a sweep with the attribute): the next pass may
    // start its CTAs as soon as SM resources free up and set up everything that does not depend on the state (program
    // headers, pair-word ring, cos / sin table, maps); it reads the state only after mqc_grid_dependency_wait().
    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
        for (int i = 0; i <= MQC_C_RING_SLOTS; ++i) mqc_mbar_init(bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (ADJ) for (int i = tid; i < (int)(NW * gw_stride); i += NT) gw[i] = 0.0;
    __syncthreads();
    uint32_t phase = 0, phbits = 0;

    const uint32_t n_tiles = P->n_tiles;
    for (uint32_t it = blockIdx.x; it < n_tiles; it += gridDim.x) {
        uint32_t ia, ib;
        if (P->tiles_explicit) { const uint32_t t = __ldg(A.tiles + P->tile_off + it); ia = t & 0xffffu; ib = t >> 16; }
        else { ia = it / (uint32_t)P->n_clsB; ib = it - ia * (uint32_t)P->n_clsB; }
        const uint4 ca4 = __ldg(reinterpret_cast<const uint4*>(A.cls + P->clsA_off + ia));
        const uint4 cb4 = __ldg(reinterpret_cast<const uint4*>(A.cls + P->clsB_off + ib));
        const uint32_t row0 = ca4.x, umask = ca4.y, col0 = cb4.x, vmask = cb4.y;
        const int ka = (int)(ca4.z & 0xffffu), nAl = (int)(ca4.z >> 16), kb = (int)(cb4.z & 0xffffu), nBl = (int)(cb4.z >> 16);
        const uint4 lt4 = __ldg(reinterpret_cast<const uint4*>(A.live_tab + P->live_tab_off + (size_t)ka * MQC_C_KMAX + kb));
        const int n_live = (int)lt4.y;
        const uint32_t* __restrict__ prog = A.prog + lt4.z;          // this class's pair words
        const int NC = (int)((lt4.w + CH - 1) / CH);   // chunks of the program
        int issued = 0, have = 0;                                    // chunks (processing order) issued / arrived
        auto issue_chunk = [&](int k) {
            const int c = ADJ ? NC - 1 - k : k;
            const int slot = c & (MQC_C_RING_SLOTS - 1);
            const uint32_t words = min((uint32_t)CH, lt4.w - (uint32_t)c * CH);
            mqc_mbar_expect_tx(rbar + slot, words * 4u);
            mqc_bulk_g2s(ring + slot * CH, prog + (size_t)c * CH, words * 4u, rbar + slot);
        };
        // wait for the next `count` chunks in processing order (header flags, mqc_c_ring_flags)
        auto wait_chunks = [&](uint32_t count) {
            for (; count; --count) {
                const int c = ADJ ? NC - 1 - have : have;
                const int slot = c & (MQC_C_RING_SLOTS - 1);
                mqc_mbar_wait(rbar + slot, (phbits >> slot) & 1u);
                phbits ^= 1u << slot;
                ++have;
            }
        };
        const uint32_t magicB = mqc_c_magic((uint32_t)nBl);
        const int cnt = nAl * nBl;
        const uint32_t shift = REAL ? (col0 & 1u) : 0u;               // slot 0 of a row sits `shift` elements into the buffer row
        AT* const sp = sp0 + shift;
        AT* const sl = sl0 + shift;
        const uint32_t row_units = REAL ? ((shift + (uint32_t)nBl + 1u) & ~1u) : (uint32_t)nBl;   // elements a row load moves

        // ---- asynchronous fetch: program headers (and the rows, forward) -> one mbarrier phase ------
        if (warp == 0) {
            const uint32_t bH = ((uint32_t)n_live * 8u + 15u) & ~15u;
            if (lane == 0) {
                mqc_fence_async();              // generic-proxy accesses of the previous tile before async-proxy writes
                mqc_mbar_expect_tx(bar, bH + (ADJ ? 0u : (uint32_t)nAl * row_units * (uint32_t)ES));
            }
            __syncwarp();
            if (lane == 31 && bH) mqc_bulk_g2s(shdr, A.prog_hdr + lt4.x, bH, bar);
            if (lane == 0) { for (; issued < NC && issued < MQC_C_RING_SLOTS; ++issued) issue_chunk(issued); }
            if (!ADJ) mqc_grid_dependency_wait();               // the previous pass (or whatever wrote psi_in) is complete
            if (!ADJ)
                for (int ja = lane; ja < nAl; ja += 32)
                    mqc_bulk_g2s(sp0 + (size_t)ja * strideB, static_cast<const AT*>(A.psi_in) + (size_t)(row0 + ja) * nB + col0 - shift,
                                 row_units * (uint32_t)ES, bar);
        }
        // ---- per-tile factor constants and the tile's row / column map ------------------------------
        for (int f = tid; f < nf; f += NT) {
            const uint4 F = __ldg(reinterpret_cast<const uint4*>(A.fac + P->fac_off + f));
            const uint32_t sg = ((uint32_t)__popc(umask & F.x) + (uint32_t)__popc(vmask & F.y) + (F.w >> 31)) & 1u;
            const double2 v = A.cs[(int)F.z];
            const double sn = ADJ ? -v.y : v.y;
            hcs[f] = make_double2(v.x, sg ? -sn : sn);
            hsg[f] = (unsigned char)sg;
        }
        for (int i = tid; i < nAl + nBl; i += NT)
            smap[i] = i < nAl ? (P->row_identity ? row0 + i : __ldg(A.maps + P->map_row_off + row0 + i))
                              : (P->col_identity ? col0 + (i - nAl) : __ldg(A.maps + P->map_col_off + col0 + (i - nAl)));
        if (ADJ) {
            __syncthreads();
            mqc_grid_dependency_wait();
            for (int e = tid; e < cnt; e += NT) {
                const int ja = (nBl == 1) ? e : (int)__umulhi((uint32_t)e, magicB), jb = e - ja * nBl;
                const size_t g = (size_t)smap[ja] * nB + smap[nAl + jb];
                mqc_cp_async_amp(sp + ja * strideB + jb, static_cast<const AT*>(A.psi_in) + g);
                mqc_cp_async_amp(sl + ja * strideB + jb, static_cast<const AT*>(A.lam_in) + g);
            }
            mqc_cp_async_wait_all();
        }
        if (ADJ) { if (tid < 8) shdr[-1 - tid] = make_uint2(0u, 0u); }
        else if (tid == 0 && !(n_live & 1)) shdr[n_live] = make_uint2(0u, 0u);   // odd n_live: the bulk copy brings the list's zero pad entry
        mqc_mbar_wait(bar, phase);
        phase ^= 1u;
        __syncthreads();

        // ---- factors: step i of the class program = header (first word, pairs | factor << 16) + its
        //      pair words in the ring; the thread's first word of step i + 1 is read before the barrier
        //      that ends step i, so between two barriers a thread only rotates ----------------------------
        if (!ADJ && REAL && n_live > 0) {
            // forward sweep, lean step loop: the header after the last step is an empty sentinel (no pairs, no
            // flags), waits and refills sit behind one uniform flag test each
            uint2 hc = shdr[0];
            wait_chunks((hc.y >> 23) & 3u);
            uint32_t wc = ring[(hc.x + (uint32_t)tid) & RMASK];
            const uint2* ph = shdr + 1;
#pragma unroll 1
            for (int i = n_live; i > 0; --i, ++ph) {
                const uint2 hn = *ph;
                const uint32_t npairs = MQC_C_NP(hc.y);
                if (hn.y & (3u << 23)) wait_chunks((hn.y >> 23) & 3u);
                if ((uint32_t)(warp * 32) < npairs) {
                    const double2 csv = hcs[MQC_C_F(hc.y)];
                    uint32_t u = (uint32_t)tid;
                    uint32_t w0 = wc;
                    for (; u + NT < npairs; u += 2 * NT) {
                        mqc_rotate_pair2<false>(reinterpret_cast<double*>(sp), reinterpret_cast<double*>(sl), w0,
                                                ring[(hc.x + u + NT) & RMASK], csv.x, csv.y);
                        w0 = ring[(hc.x + u + 2 * NT) & RMASK];
                    }
                    if (u < npairs) mqc_rotate_pair<false>(sp, sl, w0, csv.x, csv.y);
                }
                const uint32_t wn = ring[(hn.x + (uint32_t)tid) & RMASK];
                if (NT == 32) __syncwarp(); else __syncthreads();
                if (hc.y & (3u << 27)) {                                // uniform: some ring slots are free now
                    if (tid == 0) {
                        mqc_fence_async();
                        for (uint32_t cnt_is = (hc.y >> 27) & 3u; cnt_is; --cnt_is) issue_chunk(issued++);
                    }
                }
                hc = hn; wc = wn;
            }
        } else if (ADJ && REAL && n_live > 0) {
            // adjoint sweep, same lean loop in groups of eight steps (the gradient partials of a group are reduced
            // together); the group may run past the first step into the empty sentinel headers below it
            const uint2* ph = shdr + (n_live - 1);
            uint2 hc = *ph;
            wait_chunks((hc.y >> 25) & 3u);
            uint32_t wc = ring[(hc.x + (uint32_t)tid) & RMASK];
#pragma unroll 1
            for (int i0 = 0; i0 < n_live; i0 += 8) {
                const uint2* pg = ph;                                   // header of step i0 (adjoint order: descending)
                double acc8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    --ph;
                    const uint2 hn = *ph;
                    const uint32_t npairs = MQC_C_NP(hc.y);
                    if (hn.y & (3u << 25)) wait_chunks((hn.y >> 25) & 3u);
                    double gacc = 0.0;
                    if ((uint32_t)(warp * 32) < npairs) {
                        const double2 csv = hcs[MQC_C_F(hc.y)];
                        uint32_t u = (uint32_t)tid;
                        uint32_t w0 = wc;
                        for (; u + NT < npairs; u += 2 * NT) {
                            gacc += mqc_rotate_pair2<true>(reinterpret_cast<double*>(sp), reinterpret_cast<double*>(sl), w0,
                                                           ring[(hc.x + u + NT) & RMASK], csv.x, csv.y);
                            w0 = ring[(hc.x + u + 2 * NT) & RMASK];
                        }
                        if (u < npairs) gacc += mqc_rotate_pair<true>(sp, sl, w0, csv.x, csv.y);
                    }
                    const uint32_t wn = ring[(hn.x + (uint32_t)tid) & RMASK];
                    acc8[j] = gacc;
                    if (NT == 32) __syncwarp(); else __syncthreads();
                    if (hc.y & (3u << 29)) {
                        if (tid == 0) {
                            mqc_fence_async();
                            for (uint32_t cnt_is = (hc.y >> 29) & 3u; cnt_is; --cnt_is) issue_chunk(issued++);
                        }
                    }
                    hc = hn; wc = wn;
                }
                MQC_CS_BUTTERFLY(acc8, lane, w1r)
                const int j = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                if ((lane & 3) == 0 && i0 + j < n_live && w1r != 0.0) {
                    const int f = (int)MQC_C_F(pg[-j].y);
                    gw[(size_t)warp * gw_stride + f] += hsg[f] ? -w1r : w1r;
                }
            }
        } else if (n_live > 0) {
            auto step = [&](int i) -> int { return ADJ ? n_live - 1 - i : i; };
            constexpr int WSH = ADJ ? 25 : 23, ISH = ADJ ? 29 : 27;
            uint2 hc = shdr[step(0)];
            wait_chunks((hc.y >> WSH) & 3u);
            uint32_t wc = ring[(hc.x + (uint32_t)tid) & RMASK];
            constexpr int UN = ADJ ? 8 : MQC_CS_FWD_UNROLL;            // the adjoint reduces the partials of eight steps together
            for (int i0 = 0; i0 < n_live; i0 += UN) {
                double acc8[8];
#pragma unroll
                for (int j = 0; j < UN; ++j) {
                    acc8[j] = 0.0;
                    const int i = i0 + j;
                    if (i < n_live) {                                   // uniform over the CTA
                        const uint32_t npairs = MQC_C_NP(hc.y);
                        uint2 hn = make_uint2(0u, 0u);
                        uint32_t wn = 0u;
                        if (i + 1 < n_live) {
                            hn = shdr[step(i + 1)];
                            wait_chunks((hn.y >> WSH) & 3u);
                        }
                        double gacc = 0.0;
                        if ((uint32_t)(warp * 32) < npairs) {           // warps without a pair in this step only keep step
                            const double2 csv = hcs[MQC_C_F(hc.y)];
                            if (REAL) {
                                uint32_t u = (uint32_t)tid;
                                uint32_t w0 = wc;
                                for (; u + NT < npairs; u += 2 * NT) {
                                    gacc += mqc_rotate_pair2<ADJ>(reinterpret_cast<double*>(sp), reinterpret_cast<double*>(sl), w0,
                                                                  ring[(hc.x + u + NT) & RMASK], csv.x, csv.y);
                                    w0 = ring[(hc.x + u + 2 * NT) & RMASK];               // stale words past the step's end are never used
                                }
                                if (u < npairs) gacc += mqc_rotate_pair<ADJ>(sp, sl, w0, csv.x, csv.y);
                            } else {
                                if ((uint32_t)tid < npairs) gacc = mqc_rotate_pair<ADJ>(sp, sl, wc, csv.x, csv.y);
                                for (uint32_t u = (uint32_t)tid + NT; u < npairs; u += NT)
                                    gacc += mqc_rotate_pair<ADJ>(sp, sl, ring[(hc.x + u) & RMASK], csv.x, csv.y);
                            }
                        }
                        if ((uint32_t)(warp * 32) < MQC_C_NP(hn.y)) wn = ring[(hn.x + (uint32_t)tid) & RMASK];
                        if (ADJ) acc8[j] = gacc;
                        if (NT == 32) __syncwarp(); else __syncthreads();       // one warp per tile: no CTA barrier between steps
                        if (tid == 0) {                                 // slots the next step no longer needs get new chunks
                            uint32_t cnt_is = (hc.y >> ISH) & 3u;
                            if (cnt_is) {
                                mqc_fence_async();
                                for (; cnt_is; --cnt_is) issue_chunk(issued++);
                            }
                        }
                        hc = hn; wc = wn;
                    }
                }
                if (ADJ) {
                    MQC_CS_BUTTERFLY(acc8, lane, w1r)
                    const int j = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                    if ((lane & 3) == 0 && i0 + j < n_live && w1r != 0.0) {
                        const int f = (int)MQC_C_F(shdr[step(i0 + j)].y);
                        gw[(size_t)warp * gw_stride + f] += hsg[f] ? -w1r : w1r;
                    }
                }
            }
        }

        // ---- store ------------------------------------------------------------------------------------
        if (!ADJ) {
            for (int e = tid; e < cnt; e += NT) {
                const int ja = (nBl == 1) ? e : (int)__umulhi((uint32_t)e, magicB), jb = e - ja * nBl;
                static_cast<AT*>(A.psi_out)[(size_t)smap[ja] * nB + smap[nAl + jb]] = sp[ja * strideB + jb];
            }
        } else {
            // rows leave by bulk copies (generic-proxy writes made visible to the async proxy first)
            mqc_fence_async();
            __syncthreads();
            if (warp == 0) {
                const uint32_t first = shift;                                   // REAL: leading element off the 16-byte grid
                const uint32_t ni = REAL ? (((uint32_t)nBl - first) & ~1u) : (uint32_t)nBl;
                for (int ja = lane; ja < nAl; ja += 32) {
                    AT* po = static_cast<AT*>(A.psi_out) + (size_t)(row0 + ja) * nB + col0;
                    AT* lo = static_cast<AT*>(A.lam_out) + (size_t)(row0 + ja) * nB + col0;
                    const AT* ps = sp + (size_t)ja * strideB;
                    const AT* ls = sl + (size_t)ja * strideB;
                    if (ni) {
                        mqc_bulk_s2g(po + first, ps + first, ni * (uint32_t)ES);
                        mqc_bulk_s2g(lo + first, ls + first, ni * (uint32_t)ES);
                    }
                    if (REAL) {
                        if (first) { po[0] = ps[0]; lo[0] = ls[0]; }
                        if (first + ni < (uint32_t)nBl) { po[nBl - 1] = ps[nBl - 1]; lo[nBl - 1] = ls[nBl - 1]; }
                    }
                }
                mqc_bulk_commit();
                mqc_bulk_wait_read();
            }
        }
        __syncthreads();            // the tile buffers, headers and maps are free for the next tile
    }
    if (ADJ) {
        // one global atomic per (CTA, factor)
        for (int f = tid; f < nf; f += NT) {
            double g = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) g += gw[(size_t)w * gw_stride + f];
            if (A.gpart) A.gpart[(size_t)blockIdx.x * nf + f] = 2.0 * g;
            else if (g != 0.0) atomicAdd(A.grad + (A.fac[P->fac_off + f].theta_sign & 0x7fffffff), 2.0 * g);
        }
    }
}
