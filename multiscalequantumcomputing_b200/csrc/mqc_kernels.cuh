// sm_100a kernels of the UCCSD state-vector path.  All HBM-bandwidth-bound complex128
// streaming work: no tensor cores (nothing here is a dense contraction).
//
//   k_tile_sweep<ADJ>   K1/K3  fused excitation sweep over a dense 2^K shared-memory tile
//                              (ADJ: psi and lambda together + per-factor gradient partials)
//   k_tile_sparse<ADJ>  K1/K3  same sweep on the compressed (N_alpha, N_beta)-sector tile
//   k_excite_direct     K1     one factor per HBM pass (unfused baseline / fallback)
//   k_adjoint_direct    K3     one factor per pass on (psi, lambda) + gradient
//   k_happly            K2     lambda = H psi and E = <psi|H|psi>, X-mask-grouped terms,
//                              enumerated over the (N_alpha, N_beta) sector or densely
//   k_rdm1 / k_rdm2     K4     1-/2-RDM directly on determinant bit strings
#pragma once
#include <cuda_runtime.h>

#include "mqc_common.h"

struct MqcSectorDev {
    int enabled;
    int na, nb;
    u64 n_a, n_b;              // number of alpha / beta strings
    const u64* astr;           // physical-bit images of the alpha strings
    const u64* bstr;
    u64 amask, bmask;          // physical masks of alpha / beta qubits
};

struct MqcHamDev {
    int n_groups;
    const u64* gmask;          // [G] physical x-mask
    const int* goff;           // [G+1]
    const u64* tz;             // [T] physical z-mask
    const double2* tc;         // [T] coefficient
};

// Shards of the state.  The top physical index bits select the rank; every
// rank maps all peers' shards (same process: peer access; one process per GPU: CUDA IPC), so
// a physical index resolves to (peer pointer, local offset) and a tile whose active bits
// include rank bits is loaded from / stored to the peers' HBM directly over NVLink.
#define MQC_MAX_RANKS 8
struct MqcPeers {
    double2* psi[MQC_MAX_RANKS];
    double2* lam[MQC_MAX_RANKS];
    int n_local;               // address bits inside one shard
    u64 tau_hi;                // this rank's share of the pass's tiles (mqc_tile_map)
    u64 fixed_bits;            // rank bits at inactive global positions
};
__device__ __forceinline__ double2* peer_psi(const MqcPeers& R, u64 g) {
    return R.psi[g >> R.n_local] + (g & ((1ull << R.n_local) - 1ull));
}
__device__ __forceinline__ double2* peer_lam(const MqcPeers& R, u64 g) {
    return R.lam[g >> R.n_local] + (g & ((1ull << R.n_local) - 1ull));
}

__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(double2* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------
// angle of factor f = scale[f] * theta[theta_index[f]]  (scale: coefficient of the factor inside a generator that is
// a linear combination of excitations, e.g. a spin-adapted pool operator; 1 for plain UCCSD)
__global__ void k_sincos(const double* __restrict__ theta, const int* __restrict__ theta_index, const double* __restrict__ scale,
                         double2* __restrict__ cs, int n) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) {
        double s, c;
        sincos(scale[f] * theta[theta_index[f]], &s, &c);
        cs[f] = make_double2(c, s);
    }
}

// The sweep kernels accumulate dE/d(angle of factor f) per FACTOR; chain rule onto the parameters:
// g_theta[theta_index[f]] += scale[f] * g_factor[f]
__global__ void k_grad_combine(const double* __restrict__ g_factor, const int* __restrict__ theta_index, const double* __restrict__ scale,
                               double* __restrict__ g_theta, int n) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) {
        const double g = g_factor[f] * scale[f];
        if (g != 0.0) atomicAdd(g_theta + theta_index[f], g);
    }
}

// Reproducible variant: the adjoint sweep left one partial per (CTA, factor); a warp per PARAMETER walks the
// parameter's factors (th_ptr / th_fac: CSR lists in factor order) and adds each factor's partials in a fixed order
// (lane-strided, then a fixed shuffle tree): no atomics, the same bits every time, tied parameters included.
// where[f] = {offset of the factor's first partial (lo, hi), stride between CTAs, number of CTAs}.
__global__ void k_grad_combine_parts(const double* __restrict__ gpart, const uint4* __restrict__ where, const int* __restrict__ th_ptr,
                                     const int* __restrict__ th_fac, const double* __restrict__ scale, double* __restrict__ g_theta, int n_theta) {
    const int t = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= n_theta) return;
    double gt = 0.0;
    for (int i = th_ptr[t]; i < th_ptr[t + 1]; ++i) {
        const int f = th_fac[i];
        const uint4 w = where[f];
        const double* __restrict__ src = gpart + (((unsigned long long)w.y << 32) | w.x);
        double g = 0.0;
        for (unsigned c = lane; c < w.w; c += 32) g += src[(size_t)c * w.z];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        gt += scale[f] * g;
    }
    if (lane == 0) g_theta[t] = gt;
}

__global__ void k_set_amplitude(double2* psi, u64 index, double re, double im) {
    if (blockIdx.x == 0 && threadIdx.x == 0) psi[index] = make_double2(re, im);
}

__global__ void k_set_real(double* psi, u64 index, double v) {
    if (blockIdx.x == 0 && threadIdx.x == 0) psi[index] = v;
}
// real compact state -> complex copy
__global__ void k_real_to_complex(const double* __restrict__ in, double2* __restrict__ out, const u64 n, const unsigned n_b, const unsigned pitch) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 r = i / n_b;
        out[i] = make_double2(in[r * pitch + (i - r * n_b)], 0.0);          // rows of the real matrix are `pitch` apart
    }
}

// out[logical index] = psi[physical index]; phys_pos[l] = physical position of logical bit l
__global__ void k_gather_logical(const double2* __restrict__ psi, double2* __restrict__ out,
                                 const int* __restrict__ phys_pos, int n_bits, u64 dim) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < dim; i += stride) {
        u64 p = 0;
        for (int l = 0; l < n_bits; ++l) p |= ((i >> l) & 1ull) << phys_pos[l];
        out[i] = psi[p];
    }
}

// ------------------------------------------------------------------------------------------
// K1 / K3 : fused tile sweep
// ------------------------------------------------------------------------------------------
// Shared memory layout (dynamic):
//   double2 sp[2^K]                 psi tile
//   double2 sl[2^K]                 lambda tile                   (ADJ only)
//   u64     dep_lo[128], dep_hi[128]  local index -> physical offset LUTs
//   uint16  prm_lo[128], prm_hi[128]  re-layout LUTs
//   MqcTileFactor sf[nf]; double2 scs[nf]
//   double  gpart[nwarps][nf]                                     (ADJ only)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// Bank swizzle of the dense tile: a 16-byte element lives in bank group (index & 7).  Factors
// move arbitrary tile bits, so a quarter-warp's eight lanes may differ in ANY three index bits;
// the three low address bits are XORed with a GF(2)-linear image of the high bits whose
// columns are successive powers of a primitive element of GF(8) (1,2,4,3,6,7,5,...): any
// three consecutive bit positions - and most other triples - map to independent vectors,
// i.e. eight distinct bank groups.  Linear, so swz(x ^ m) = swz(x) ^ swz(m).
__device__ __forceinline__ uint32_t swz(uint32_t t) {
    return t ^ ((uint32_t)__popc(t & MQC_SWZ_M0) & 1u) ^ (((uint32_t)__popc(t & MQC_SWZ_M1) & 1u) << 1) ^
           (((uint32_t)__popc(t & MQC_SWZ_M2) & 1u) << 2);
}
__device__ __forceinline__ uint32_t ins0(uint32_t x, uint32_t p) {
    return ((x >> p) << (p + 1)) | (x & ((1u << p) - 1u));
}

// ------------------------------------------------------------------------------------------
// K1 / K3, sector form : compressed tile
// ------------------------------------------------------------------------------------------
// When every ansatz factor conserves (N_alpha, N_beta), a tile with a_act alpha and b_act beta
// active bits can only hold C(a_act, ka) * C(b_act, kb) non-zero amplitudes (ka, kb fixed by the
// tile's inactive bits): <= 400 of 4096 for a 6+6-bit tile.  This kernel keeps only those in
// shared memory (<= 6.4 KB per state instead of 64 KB), so ~10 CTAs per SM overlap their
// per-factor latency chains instead of 3, touches only those amplitudes in HBM/L2 (the zeros
// outside the sector are never read or written), and skips tiles that cannot hold any.
//
//   slist[e]  : slot -> tile-local index (ascending)         smap[t] : tile-local index -> slot
// A factor scans the list; element xl is the lower partner iff (xl & m) == v; its partner
// xl ^ m is in the same sector, slot = smap[xl ^ m].
struct MqcSparse {
    u64 amask, bmask;          // physical masks of the alpha / beta qubits
    int na, nb;
    int cap;                   // slots reserved per state (max elements of any tile of this pass)
};

#define MQC_FCHUNK 32

template <bool ADJ>
__global__ void __launch_bounds__(256)
k_tile_sparse(const __grid_constant__ MqcPeers PR, const __grid_constant__ MqcPass P,
              const MqcTileFactor* __restrict__ tf, const double2* __restrict__ cs,
              double* __restrict__ grad, const int reverse, const MqcSparse SP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = P.k;
    const uint32_t T = 1u << K;
    const int nf = P.f1 - P.f0;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int h0 = K >> 1, h1 = K - h0;
    const uint32_t lo_mask = (1u << h0) - 1u;
    const int cap = SP.cap;

    double2* sp = reinterpret_cast<double2*>(smem_raw);
    double2* sl = sp + (ADJ ? cap : 0);
    unsigned char* q = smem_raw + (size_t)(ADJ ? 2 : 1) * cap * sizeof(double2);
    u64* dep_lo = reinterpret_cast<u64*>(q);            q += 128 * sizeof(u64);
    u64* dep_hi = reinterpret_cast<u64*>(q);            q += 128 * sizeof(u64);
    double2* scs = reinterpret_cast<double2*>(q);       q += MQC_FCHUNK * sizeof(double2);
    double* gpart = reinterpret_cast<double*>(q);       q += (ADJ ? 8 * MQC_FCHUNK : 0) * sizeof(double);
    int* stheta = reinterpret_cast<int*>(q);            q += MQC_FCHUNK * sizeof(int);
    uint32_t* smz = reinterpret_cast<uint32_t*>(q);     q += MQC_FCHUNK * sizeof(uint32_t);     // m | z << 16
    uint32_t* svs = reinterpret_cast<uint32_t*>(q);     q += MQC_FCHUNK * sizeof(uint32_t);     // v | sign << 16
    int* spos = reinterpret_cast<int*>(q);              q += 16 * sizeof(int);
    int* sperm = reinterpret_cast<int*>(q);             q += 16 * sizeof(int);
    int* scount = reinterpret_cast<int*>(q);            q += 4 * sizeof(int);
    uint16_t* inv_lo = reinterpret_cast<uint16_t*>(q);  q += 128 * sizeof(uint16_t);
    uint16_t* inv_hi = reinterpret_cast<uint16_t*>(q);  q += 128 * sizeof(uint16_t);
    uint16_t* slist = reinterpret_cast<uint16_t*>(q);   q += (size_t)cap * sizeof(uint16_t);
    uint16_t* smap = reinterpret_cast<uint16_t*>(q);

    if (tid < 16) {
        int pv = 0, qv = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) if (j == tid) { pv = P.pos[j]; qv = P.perm[j]; }
        spos[tid] = pv; sperm[tid] = qv;
        if (tid == 0) scount[0] = 0;
    }
    __syncthreads();
    // tile counter -> inactive LOCAL bits; inactive rank bits are this rank's own
    const u64 base = mqc_deposit_complement((u64)blockIdx.x | PR.tau_hi, P.act_mask, PR.n_local) | PR.fixed_bits;
    const u64 sbase = base;
    uint32_t aloc = 0, bloc = 0;
    for (int j = 0; j < K; ++j) {
        aloc |= (uint32_t)((SP.amask >> spos[j]) & 1ull) << j;
        bloc |= (uint32_t)((SP.bmask >> spos[j]) & 1ull) << j;
    }
    const int ka = SP.na - __popcll(sbase & SP.amask);
    const int kb = SP.nb - __popcll(sbase & SP.bmask);
    if (ka < 0 || kb < 0 || ka > __popc(aloc) || kb > __popc(bloc)) return;      // tile holds only zeros

    // physical offset of a tile-local index, and the local index an element has in the OTHER
    // layout of a re-layout pass: t' with bit j' = bit perm[j'] of t
    for (int t = tid; t < (1 << h0); t += nthr) {
        u64 d = 0;
        for (int j = 0; j < h0; ++j) d |= (((u64)t >> j) & 1ull) << spos[j];
        dep_lo[t] = d;
    }
    for (int t = tid; t < (1 << h1); t += nthr) {
        u64 d = 0;
        for (int j = 0; j < h1; ++j) d |= (((u64)t >> j) & 1ull) << spos[h0 + j];
        dep_hi[t] = d;
    }
    // inverse permutation as two gathers: contribution of the low / high half of t to t'
    for (int t = tid; t < (1 << h0); t += nthr) {
        uint32_t o = 0;
        for (int jp = 0; jp < K; ++jp) { const int src = sperm[jp]; if (src < h0) o |= (((uint32_t)t >> src) & 1u) << jp; }
        inv_lo[t] = (uint16_t)o;
    }
    for (int t = tid; t < (1 << h1); t += nthr) {
        uint32_t o = 0;
        for (int jp = 0; jp < K; ++jp) { const int src = sperm[jp]; if (src >= h0) o |= (((uint32_t)t >> (src - h0)) & 1u) << jp; }
        inv_hi[t] = (uint16_t)o;
    }
    // sector elements of this tile, ascending (block-wide ordered compaction)
    {
        __shared__ int wcount[8];
        const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
        int running = 0;
        for (uint32_t t0 = 0; t0 < T; t0 += nthr) {
            const uint32_t t = t0 + tid;
            const bool ok = (t < T) && (__popc(t & aloc) == ka) && (__popc(t & bloc) == kb);
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) wcount[warp] = __popc(bal);
            __syncthreads();
            int off = running, tot = 0;
            for (int w = 0; w < nw; ++w) { const int c = wcount[w]; if (w < warp) off += c; tot += c; }
            if (ok) {
                const int o = off + __popc(bal & ((1u << lane) - 1u));
                if (o < cap) { slist[o] = (uint16_t)t; smap[t] = (uint16_t)o; }
            }
            running += tot;
            __syncthreads();
        }
        if (tid == 0) scount[0] = min(running, cap);
    }
    __syncthreads();
    const int cnt = scount[0];

    // ---- load -----------------------------------------------------------------------
    const bool other_layout_load = reverse && P.has_perm;
    const bool other_layout_store = (!reverse) && P.has_perm;
    for (int e = tid; e < cnt; e += nthr) {
        uint32_t t = slist[e];
        if (other_layout_load) t = (uint32_t)(inv_lo[t & lo_mask] | inv_hi[t >> h0]);
        const u64 g = base | dep_lo[t & lo_mask] | dep_hi[t >> h0];
        cp_async16(sp + e, peer_psi(PR, g));
        if (ADJ) cp_async16(sl + e, peer_lam(PR, g));
    }

    // ---- factors, in chunks ---------------------------------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    const int nw = nthr >> 5;
    for (int c0 = 0; c0 < nf; c0 += MQC_FCHUNK) {
        const int nc = min(MQC_FCHUNK, nf - c0);
        for (int fc = tid; fc < nc; fc += nthr) {
            const int f = reverse ? (nf - 1 - (c0 + fc)) : (c0 + fc);
            const MqcTileFactor* Fp = tf + P.f0 + f;
            const double2 v = cs[Fp->slot];
            scs[fc] = make_double2(v.x, reverse ? -v.y : v.y);
            stheta[fc] = Fp->theta;
            const uint32_t sg = ((uint32_t)__popcll(sbase & Fp->z_out) + Fp->sign0) & 1u;
            smz[fc] = (uint32_t)Fp->m | ((uint32_t)Fp->z << 16);
            svs[fc] = (uint32_t)Fp->v | (sg << 16);
        }
        if (c0 == 0) cp_async_wait_all();
        __syncthreads();
        for (int fc = 0; fc < nc; ++fc) {
            const uint32_t mz = smz[fc], vs = svs[fc];
            const uint32_t fm = mz & 0xffffu, fz = mz >> 16, fv = vs & 0xffffu, sg = vs >> 16;
            const double2 csv = scs[fc];
            const double c = csv.x, sn = csv.y;
            double gacc = 0.0;
            for (int e = tid; e < cnt; e += nthr) {
                const uint32_t xl = slist[e];
                if ((xl & fm) != fv) continue;
                const uint32_t e1 = smap[xl ^ fm];
                const bool neg = (((uint32_t)__popc(xl & fz) + sg) & 1u) != 0u;
                const double ssn = neg ? -sn : sn;
                const double2 a = sp[e], b = sp[e1];
                if (ADJ) {
                    const double2 la = sl[e], lb = sl[e1];
                    const double gv = lb.x * a.x + lb.y * a.y - la.x * b.x - la.y * b.y;
                    gacc += neg ? -gv : gv;
                    sl[e] = make_double2(c * la.x - ssn * lb.x, c * la.y - ssn * lb.y);
                    sl[e1] = make_double2(c * lb.x + ssn * la.x, c * lb.y + ssn * la.y);
                }
                sp[e] = make_double2(c * a.x - ssn * b.x, c * a.y - ssn * b.y);
                sp[e1] = make_double2(c * b.x + ssn * a.x, c * b.y + ssn * a.y);
            }
            if (ADJ) {
                gacc = warp_sum(gacc);
                if (lane == 0) gpart[warp * MQC_FCHUNK + fc] = gacc;
            }
            __syncthreads();
        }
        if (ADJ) {
            for (int fc = tid; fc < nc; fc += nthr) {
                double g = 0.0;
                for (int w = 0; w < nw; ++w) g += gpart[w * MQC_FCHUNK + fc];
                if (g != 0.0) atomicAdd(grad + stheta[fc], 2.0 * g);
            }
            __syncthreads();
        }
    }
    if (nf == 0) { cp_async_wait_all(); __syncthreads(); }

    // ---- store ----------------------------------------------------------------------
    // A re-layout pass moves the sector elements to other addresses of the same tile: clear
    // the addresses they were loaded from first, so that everything outside the sector stays
    // exactly zero in HBM (the dense kernels and mqc_state read it).
    if (P.has_perm) {
        for (int e = tid; e < cnt; e += nthr) {
            uint32_t t = slist[e];
            if (other_layout_load) t = (uint32_t)(inv_lo[t & lo_mask] | inv_hi[t >> h0]);
            const u64 g = base | dep_lo[t & lo_mask] | dep_hi[t >> h0];
            *peer_psi(PR, g) = make_double2(0.0, 0.0);
            if (ADJ) *peer_lam(PR, g) = make_double2(0.0, 0.0);
        }
        __syncthreads();
    }
    for (int e = tid; e < cnt; e += nthr) {
        uint32_t t = slist[e];
        if (other_layout_store) t = (uint32_t)(inv_lo[t & lo_mask] | inv_hi[t >> h0]);
        const u64 g = base | dep_lo[t & lo_mask] | dep_hi[t >> h0];
        *peer_psi(PR, g) = sp[e];
        if (ADJ) *peer_lam(PR, g) = sl[e];
    }
}

// ------------------------------------------------------------------------------------------
// K1 / K3, sector form, second generation: one WARP per compressed tile, pair lists
// ------------------------------------------------------------------------------------------
// The non-zero amplitudes of a tile are the product set {alpha-local patterns with ka bits} x
// {beta-local patterns with kb bits} (MqcSpPass): slot = ja * nBl + jb.  The host tabulates,
// per (factor, ka) and (factor, kb), the pattern pairs the factor couples on each spin species
// (identity pairs where it does not act), so the lanes of a warp enumerate exactly the coupled
// slot pairs  listA x listB  - no scan over the tile, no index->slot map, no block barrier: a
// warp owns its tile and factors are separated by __syncwarp().  Warps are persistent (grid =
// resident warps, tiles strided), which also lets a CTA sum its gradient partials in shared
// memory and issue one global atomic per (CTA, factor) instead of one per (tile, factor).
struct MqcSpTables {
    const u64* dep;            // pattern -> physical offset, load layout
    const u64* depp;           // pattern -> physical offset, other layout of a re-layout pass
    const uint32_t* lists;     // pair lists
    const uint16_t* pat;       // pattern -> tile-local mask
    const uint4* hdr;          // (factor, k) -> {first entry, count, float bits of 1 / count, parity mask | identity << 16}
    const u64* tiles;          // optional: non-empty tiles of this rank, heaviest first:
    u64 n_list;                //   base (48 bits) | ka << 48 | kb << 56
};
#define MQC_SP2_HDR_BYTES (32 * (16 + 16 + 16 + 4 + 4))
static inline __host__ __device__ size_t mqc_sp2_dep_bytes(const MqcSpPass& SP) {     // [2 layouts][maxA + maxB] offsets per warp
    int ma = 1, mb = 1;
    for (int k = 0; k <= SP.a_act; ++k) ma = SP.cntA[k] > ma ? SP.cntA[k] : ma;
    for (int k = 0; k <= SP.b_act; ++k) mb = SP.cntB[k] > mb ? SP.cntB[k] : mb;
    return (size_t)2 * (ma + mb) * sizeof(u64) + (((size_t)(ma + mb) * sizeof(uint16_t) + 15) & ~(size_t)15)
           + (((size_t)SP.lcap * sizeof(uint16_t) + 15) & ~(size_t)15);
}
static inline __host__ __device__ int mqc_sp2_max_patterns(const MqcSpPass& SP) {
    int ma = 1, mb = 1;
    for (int k = 0; k <= SP.a_act; ++k) ma = SP.cntA[k] > ma ? SP.cntA[k] : ma;
    for (int k = 0; k <= SP.b_act; ++k) mb = SP.cntB[k] > mb ? SP.cntB[k] : mb;
    return ma + mb;
}

// LIST: tiles come from a host-built list (only the non-empty ones, sorted by size and dealt
// round-robin over the CTAs so that every SM gets the same mix of large and small tiles);
// otherwise every tile counter is decoded and tested on the fly (huge sharded registers).
// TW: warps per TEAM.  A team owns a tile; with TW == 1 factors are separated by __syncwarp(),
// with TW > 1 by a named barrier over the team's TW * 32 threads (14-bit tiles hold up to 1225
// amplitudes and 100-350 pairs per factor: enough work for several warps, and four times
// as many warps in flight as tiles).
template <int TW>
__device__ __forceinline__ void team_sync(const int team) {
    if (TW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(TW * 32) : "memory");
}

template <bool ADJ, bool LIST, int TW>
__global__ void __launch_bounds__(256)
k_tile_sp2(const __grid_constant__ MqcPeers PR, const __grid_constant__ MqcPass P, const __grid_constant__ MqcSpPass SP,
           const MqcSpTables TB, const MqcTileFactor* __restrict__ tf, const double2* __restrict__ cs,
           double* __restrict__ grad, const int reverse, const u64 n_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_teams = W / TW, team = warp / TW, wt = warp - team * TW;      // warp within its team
    const int tl = wt * 32 + lane;                                            // thread within its team
    constexpr int TT = TW * 32;
    const int nf = P.f1 - P.f0;
    const int cap = SP.cap;
    const size_t gw_bytes = ADJ ? (((size_t)nf * sizeof(double) + 15) & ~(size_t)15) : 0;
    const size_t dep_bytes = mqc_sp2_dep_bytes(SP);
    const size_t per_team = (size_t)(ADJ ? 2 : 1) * cap * sizeof(double2) + MQC_SP2_HDR_BYTES + dep_bytes;
    unsigned char* wb = smem_raw + (size_t)team * per_team;
    double2* sp = reinterpret_cast<double2*>(wb);
    double2* sl = sp + (ADJ ? cap : 0);
    unsigned char* q = wb + (size_t)(ADJ ? 2 : 1) * cap * sizeof(double2);
    const int maxpat = mqc_sp2_max_patterns(SP);
    u64* sdl = reinterpret_cast<u64*>(q);                                 // offsets of this tile's patterns: load layout,
    u64* sds = sdl + maxpat;                                              // store layout  ([alpha patterns][beta patterns])
    uint16_t* spat = reinterpret_cast<uint16_t*>(sds + maxpat);           // the patterns themselves
    uint16_t* slist = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(spat) + (((size_t)maxpat * sizeof(uint16_t) + 15) & ~(size_t)15));
    q += dep_bytes;
    double2* hcs = reinterpret_cast<double2*>(q);   q += 32 * sizeof(double2);
    uint4* hA = reinterpret_cast<uint4*>(q);        q += 32 * sizeof(uint4);
    uint4* hB = reinterpret_cast<uint4*>(q);        q += 32 * sizeof(uint4);
    uint32_t* hsg = reinterpret_cast<uint32_t*>(q); q += 32 * sizeof(uint32_t);
    int* hfi = reinterpret_cast<int*>(q);           q += 32 * sizeof(int);
    // [warp][nf] gradient partials (ADJ), after all team regions; the last word of the team's
    // header block carries the chunk's live-factor mask from warp 0 to the other warps
    double* gw = reinterpret_cast<double*>(smem_raw + (size_t)n_teams * per_team + (size_t)warp * gw_bytes);
    __shared__ unsigned s_todo[8];
    if (ADJ) {
        for (int i = lane; i < nf; i += 32) gw[i] = 0.0;
        __syncwarp();
    }
    const bool alt_load = reverse && P.has_perm, alt_store = (!reverse) && P.has_perm;
    const u64* __restrict__ dl = alt_load ? TB.depp : TB.dep;
    const u64* __restrict__ ds = alt_store ? TB.depp : TB.dep;
    const u64 n_work = LIST ? TB.n_list : n_tiles;

    for (u64 it = (u64)blockIdx.x + (u64)gridDim.x * team; it < n_work; it += (u64)gridDim.x * n_teams) {
        u64 base;
        int ka, kb;
        if (LIST) {
            const u64 pk = __ldg(TB.tiles + it);
            base = pk & 0xffffffffffffull; ka = (int)((pk >> 48) & 0xffu); kb = (int)(pk >> 56);
        } else {
            base = mqc_deposit_complement(it | PR.tau_hi, P.act_mask, PR.n_local) | PR.fixed_bits;
            ka = SP.na - __popcll(base & SP.amask);
            kb = SP.nb - __popcll(base & SP.bmask);
            if (ka < 0 || kb < 0 || ka > SP.a_act || kb > SP.b_act) continue;   // tile holds only zeros (uniform over the team)
        }
        const int nAl = SP.cntA[ka], nBl = SP.cntB[kb];
        const int cnt = nAl * nBl;
        const uint32_t offA = SP.patoffA[ka], offB = SP.patoffB[kb];
        const float rcpB = __frcp_rn((float)nBl);
        team_sync<TW>(team);                                      // previous tile's stores have read sp / sl / sd*
        for (int i = tl; i < nAl + nBl; i += TT) {
            const uint32_t src = i < nAl ? offA + i : offB + (i - nAl);
            sdl[i] = base | __ldg(dl + src);
            sds[i] = base | __ldg(ds + src);
            spat[i] = __ldg(TB.pat + src);
        }
        team_sync<TW>(team);
        // ---- load ---------------------------------------------------------------------------
        for (int e = tl; e < cnt; e += TT) {
            const int ja = (int)(((float)e + 0.5f) * rcpB), jb = e - ja * nBl;
            const u64 g = sdl[ja] | sdl[nAl + jb];
            cp_async16(sp + e, peer_psi(PR, g));
            if (ADJ) cp_async16(sl + e, peer_lam(PR, g));
        }
        // ---- factors ------------------------------------------------------------------------
        for (int c0 = 0; c0 < nf; c0 += 32) {
            const int nc = min(32, nf - c0);
            team_sync<TW>(team);                                  // previous chunk's headers are consumed
            if (wt == 0) {
                bool live = false;
                uint4 a4 = make_uint4(0, 0, 0, 0), b4 = make_uint4(0, 0, 0, 0);
                int nstage = 0;
                if (lane < nc) {
                    const int f = reverse ? (nf - 1 - (c0 + lane)) : (c0 + lane);
                    a4 = __ldg(TB.hdr + SP.hdrA_off + (size_t)f * MQC_SP_KMAX + ka);
                    b4 = __ldg(TB.hdr + SP.hdrB_off + (size_t)f * MQC_SP_KMAX + kb);
                    live = (a4.y != 0u) && (b4.y != 0u);
                    if (live) {
                        const MqcTileFactor* Fp = tf + P.f0 + f;
                        const double2 v = cs[Fp->slot];
                        hcs[lane] = make_double2(v.x, reverse ? -v.y : v.y);
                        hsg[lane] = ((uint32_t)__popcll(base & Fp->z_out) + Fp->sign0) & 1u;
                        hfi[lane] = f;
                        nstage = (int)(((a4.w >> 16) & 1u) ? 0u : a4.y) + (int)(((b4.w >> 16) & 1u) ? 0u : b4.y);
                    }
                }
                // the short pair lists of the acting sides go to shared memory (one batch of loads
                // per chunk instead of dependent global loads inside the pair loop)
                int soff = nstage;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, soff, o); if (lane >= o) soff += t; }
                soff -= nstage;
                if (live) {
                    int w = soff;
                    if (!((a4.w >> 16) & 1u)) {
                        for (uint32_t j = 0; j < a4.y; ++j) { const uint32_t en = __ldg(TB.lists + a4.x + j); slist[w + j] = (uint16_t)((en & 0x7fu) | (((en >> 12) & 0x7fu) << 7) | (((en >> 24) & 1u) << 14)); }
                        a4.x = (uint32_t)w; w += (int)a4.y;
                    }
                    if (!((b4.w >> 16) & 1u)) {
                        for (uint32_t j = 0; j < b4.y; ++j) { const uint32_t en = __ldg(TB.lists + b4.x + j); slist[w + j] = (uint16_t)((en & 0x7fu) | (((en >> 12) & 0x7fu) << 7) | (((en >> 24) & 1u) << 14)); }
                        b4.x = (uint32_t)w;
                    }
                    hA[lane] = a4; hB[lane] = b4;
                }
                const unsigned m = __ballot_sync(0xffffffffu, live);   // factors that couple anything in this tile
                if (lane == 0) s_todo[team] = m;
            }
            if (c0 == 0) cp_async_wait_all();
            team_sync<TW>(team);
            unsigned todo = s_todo[team];
            // Factors are applied one after the other (barrier in between), but the warp
            // reductions of their gradient partials are deferred: eight factors keep their
            // per-lane partials in registers and are reduced together, as independent shuffle
            // chains off the critical path (one reduction per factor cost ~200 cycles of pure
            // latency between two barriers).
            constexpr int GR = ADJ ? 8 : 1;
            while (todo) {
                double acc8[GR];
                int fi8[GR];
#pragma unroll
                for (int j = 0; j < GR; ++j) {
                    acc8[j] = 0.0; fi8[j] = -1;
                    if (todo) {                                   // uniform over the team
                        const int fc = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const uint4 A = hA[fc], B = hB[fc];
                        const int nLB = (int)B.y;
                        const int npairs = (int)A.y * nLB;
                        const double2 csv = hcs[fc];
                        const double c = csv.x, sn = csv.y;
                        const uint32_t sg = hsg[fc];
                        const float rcp = __uint_as_float(B.z);
                        const bool idA = (A.w >> 16) & 1u, idB = (B.w >> 16) & 1u;
                        double gacc = 0.0;
                        for (int u = tl; u < npairs; u += TT) {
                            const int ua = (int)(((float)u + 0.5f) * rcp), ub = u - ua * nLB;
                            const uint32_t la = idA ? ((uint32_t)ua * 129u | (((uint32_t)__popc(spat[ua] & A.w) & 1u) << 14)) : (uint32_t)slist[A.x + ua];
                            const uint32_t lb = idB ? ((uint32_t)ub * 129u | (((uint32_t)__popc(spat[nAl + ub] & B.w) & 1u) << 14)) : (uint32_t)slist[B.x + ub];
                            const int e0 = (int)(la & 0x7fu) * nBl + (int)(lb & 0x7fu);
                            const int e1 = (int)((la >> 7) & 0x7fu) * nBl + (int)((lb >> 7) & 0x7fu);
                            const bool neg = ((((la ^ lb) >> 14) ^ sg) & 1u) != 0u;
                            const double ssn = neg ? -sn : sn;
                            const double2 a = sp[e0], b = sp[e1];
                            if (ADJ) {
                                const double2 xa = sl[e0], xb = sl[e1];
                                const double gv = xb.x * a.x + xb.y * a.y - xa.x * b.x - xa.y * b.y;
                                gacc += neg ? -gv : gv;
                                sl[e0] = make_double2(c * xa.x - ssn * xb.x, c * xa.y - ssn * xb.y);
                                sl[e1] = make_double2(c * xb.x + ssn * xa.x, c * xb.y + ssn * xa.y);
                            }
                            sp[e0] = make_double2(c * a.x - ssn * b.x, c * a.y - ssn * b.y);
                            sp[e1] = make_double2(c * b.x + ssn * a.x, c * b.y + ssn * a.y);
                        }
                        if (ADJ) { acc8[j] = gacc; fi8[j] = hfi[fc]; }
                        team_sync<TW>(team);
                    }
                }
                if (ADJ) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int j = 0; j < GR; ++j) acc8[j] += __shfl_xor_sync(0xffffffffu, acc8[j], o);
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < GR; ++j) if (fi8[j] >= 0 && acc8[j] != 0.0) gw[fi8[j]] += acc8[j];
                    }
                }
            }
        }
        if (nf == 0) { cp_async_wait_all(); team_sync<TW>(team); }
        // ---- store --------------------------------------------------------------------------
        // a re-layout pass moves the elements to other addresses of the same tile: clear the
        // addresses they came from first (everything outside the sector stays exactly zero)
        if (P.has_perm) {
            for (int e = tl; e < cnt; e += TT) {
                const int ja = (int)(((float)e + 0.5f) * rcpB), jb = e - ja * nBl;
                const u64 g = sdl[ja] | sdl[nAl + jb];
                *peer_psi(PR, g) = make_double2(0.0, 0.0);
                if (ADJ) *peer_lam(PR, g) = make_double2(0.0, 0.0);
            }
            team_sync<TW>(team);
        }
        for (int e = tl; e < cnt; e += TT) {
            const int ja = (int)(((float)e + 0.5f) * rcpB), jb = e - ja * nBl;
            const u64 g = sds[ja] | sds[nAl + jb];
            *peer_psi(PR, g) = sp[e];
            if (ADJ) *peer_lam(PR, g) = sl[e];
        }
    }
    if (ADJ) {
        // CTA-level sum of the warps' partials, then one global atomic per (CTA, factor)
        __syncthreads();
        const double* gall = reinterpret_cast<const double*>(smem_raw + (size_t)n_teams * per_team);
        for (int i = threadIdx.x; i < nf; i += blockDim.x) {
            double g = 0.0;
            for (int w = 0; w < W; ++w) g += gall[(size_t)w * (gw_bytes / sizeof(double)) + i];
            if (g != 0.0) atomicAdd(grad + tf[P.f0 + i].theta, 2.0 * g);
        }
    }
}

static inline size_t mqc_sp2_smem_bytes(const MqcSpPass& SP, int nf, bool adj, int warps, int team_warps) {
    const size_t gw_bytes = adj ? (((size_t)nf * sizeof(double) + 15) & ~(size_t)15) : 0;
    const size_t per_team = (size_t)(adj ? 2 : 1) * SP.cap * sizeof(double2) + MQC_SP2_HDR_BYTES + mqc_sp2_dep_bytes(SP);
    return (size_t)(warps / team_warps) * per_team + (size_t)warps * gw_bytes;
}

static inline size_t mqc_sparse_smem_bytes(int K, int cap, bool adj) {
    size_t b = (size_t)(adj ? 2 : 1) * cap * sizeof(double2);
    b += 2 * 128 * sizeof(u64);
    b += MQC_FCHUNK * sizeof(double2);
    if (adj) b += 8 * MQC_FCHUNK * sizeof(double);
    b += MQC_FCHUNK * sizeof(int) + 2 * MQC_FCHUNK * sizeof(uint32_t);
    b += 32 * sizeof(int) + 4 * sizeof(int);
    b += 2 * 128 * sizeof(uint16_t);
    b += (size_t)cap * sizeof(uint16_t);
    b += ((size_t)1 << K) * sizeof(uint16_t);
    return b;
}

// Per-factor pair addressing.  For a factor with flip mask m the lower partner of pair u is
//   xl(u) = ins(u) | v      (u with zero bits inserted at the positions of m)
// and both ins() and the bank swizzle are GF(2)-linear, as is the parity popc(xl & z).  So
//   [ swz(xl(u)) | parity(u) << 15 ] = A[u & 31] ^ B[(u >> 5) & 7] ^ C[u >> 8]
// with three tiny tables per factor (48 x u16; v folded into A).  They do not depend on the
// tile, so the HOST builds them once per schedule (mqc_build_dense_luts, 128 bytes per factor,
// entries 48 / 49: swz(m) and log2(#pairs)) and a CTA only copies the chunk's tables; the
// per-tile Jordan-Wigner sign goes into the sign of sin(theta) instead.  A thread's A and B
// entries are loop invariants (u & 31 = lane, (u >> 5) & 7 = warp & 7 for 256 / 512 threads).
template <bool ADJ>
__global__ void __launch_bounds__(512)
k_tile_sweep(const __grid_constant__ MqcPeers PR, const __grid_constant__ MqcPass P,
             const MqcTileFactor* __restrict__ tf, const uint16_t* __restrict__ lutg, const double2* __restrict__ cs,
             double* __restrict__ grad, const int reverse) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = P.k;
    const uint32_t T = 1u << K;
    const int nf = P.f1 - P.f0;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int h0 = K >> 1, h1 = K - h0;
    const uint32_t lo_mask = (1u << h0) - 1u;

    double2* sp = reinterpret_cast<double2*>(smem_raw);
    double2* sl = sp + (ADJ ? T : 0);
    unsigned char* q = smem_raw + (size_t)(ADJ ? 2 : 1) * T * sizeof(double2);
    u64* dep_lo = reinterpret_cast<u64*>(q);            q += 128 * sizeof(u64);
    u64* dep_hi = reinterpret_cast<u64*>(q);            q += 128 * sizeof(u64);
    double2* scs = reinterpret_cast<double2*>(q);       q += MQC_FCHUNK * sizeof(double2);
    double* gpart = reinterpret_cast<double*>(q);       q += (ADJ ? 16 * MQC_FCHUNK : 0) * sizeof(double);
    int* stheta = reinterpret_cast<int*>(q);            q += MQC_FCHUNK * sizeof(int);
    uint16_t* slut = reinterpret_cast<uint16_t*>(q);    q += MQC_FCHUNK * MQC_LUT_STRIDE * sizeof(uint16_t);
    uint16_t* swl_lo = reinterpret_cast<uint16_t*>(q);  q += 128 * sizeof(uint16_t);     // swizzled slot of an element
    uint16_t* swl_hi = reinterpret_cast<uint16_t*>(q);  q += 128 * sizeof(uint16_t);     //   in the load layout,
    uint16_t* sws_lo = reinterpret_cast<uint16_t*>(q);  q += 128 * sizeof(uint16_t);     //   in the store layout
    uint16_t* sws_hi = reinterpret_cast<uint16_t*>(q);  q += 128 * sizeof(uint16_t);
    int* spos = reinterpret_cast<int*>(q);              q += 16 * sizeof(int);
    int* sperm = reinterpret_cast<int*>(q);             q += 16 * sizeof(int);

    // ---- per-CTA tables -------------------------------------------------------------
    if (tid < 16) {                      // static indices only: no local-memory copy of P
        int pv = 0, qv = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) if (j == tid) { pv = P.pos[j]; qv = P.perm[j]; }
        spos[tid] = pv; sperm[tid] = qv;
    }
    __syncthreads();
    // the tile is held in the forward pass's LOAD layout in both directions; the other layout's
    // index of element t is perm(t); the swizzle is linear, so both split into low / high halves
    const bool perm_load = reverse && P.has_perm;
    const bool perm_store = (!reverse) && P.has_perm;
    for (int t = tid; t < (1 << h0); t += nthr) {
        u64 d = 0; uint32_t pm = 0;
        for (int j = 0; j < h0; ++j) { const u64 bb = ((u64)t >> j) & 1ull; d |= bb << spos[j]; pm |= (uint32_t)bb << sperm[j]; }
        dep_lo[t] = d;
        swl_lo[t] = (uint16_t)swz(perm_load ? pm : (uint32_t)t);
        sws_lo[t] = (uint16_t)swz(perm_store ? pm : (uint32_t)t);
    }
    for (int t = tid; t < (1 << h1); t += nthr) {
        u64 d = 0; uint32_t pm = 0;
        for (int j = 0; j < h1; ++j) { const u64 bb = ((u64)t >> j) & 1ull; d |= bb << spos[h0 + j]; pm |= (uint32_t)bb << sperm[h0 + j]; }
        dep_hi[t] = d;
        swl_hi[t] = (uint16_t)swz(perm_load ? pm : ((uint32_t)t << h0));
        sws_hi[t] = (uint16_t)swz(perm_store ? pm : ((uint32_t)t << h0));
    }
    const u64 base = mqc_deposit_complement((u64)blockIdx.x | PR.tau_hi, P.act_mask, PR.n_local) | PR.fixed_bits;
    __syncthreads();

    // ---- load (cp.async: whole tile in flight, no register staging) -------------------
    for (uint32_t t = tid; t < T; t += nthr) {
        const u64 g = base | dep_lo[t & lo_mask] | dep_hi[t >> h0];
        const uint32_t sw = (uint32_t)swl_lo[t & lo_mask] ^ (uint32_t)swl_hi[t >> h0];
        cp_async16(sp + sw, peer_psi(PR, g));
        if (ADJ) cp_async16(sl + sw, peer_lam(PR, g));
    }

    // ---- factors, in chunks ---------------------------------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    const int nw = nthr >> 5;
    const bool fixed_b = (nthr & 255) == 0;                 // (u >> 5) & 7 does not change along a thread's pairs
    for (int c0 = 0; c0 < nf; c0 += MQC_FCHUNK) {
        const int nc = min(MQC_FCHUNK, nf - c0);
        // chunk tables: 64 u16 per factor, copied as 32 words
        for (int w = tid; w < nc * (MQC_LUT_STRIDE / 2); w += nthr) {
            const int fc = w >> 5, e = w & 31;
            const int f = reverse ? (nf - 1 - (c0 + fc)) : (c0 + fc);
            reinterpret_cast<uint32_t*>(slut)[w] = __ldg(reinterpret_cast<const uint32_t*>(lutg) + (size_t)(P.f0 + f) * (MQC_LUT_STRIDE / 2) + e);
        }
        for (int fc = tid; fc < nc; fc += nthr) {
            const int f = reverse ? (nf - 1 - (c0 + fc)) : (c0 + fc);
            const MqcTileFactor* Fp = tf + P.f0 + f;
            const double2 v = cs[Fp->slot];
            const uint32_t sg = ((uint32_t)__popcll(base & Fp->z_out) + Fp->sign0) & 1u;   // per-tile JW sign
            const double sn = reverse ? -v.y : v.y;
            scs[fc] = make_double2(v.x, sg ? -sn : sn);
            stheta[fc] = Fp->theta | (int)(sg << 30);       // the gradient needs the sign once more
        }
        if (c0 == 0) cp_async_wait_all();
        __syncthreads();

        for (int fc = 0; fc < nc; ++fc) {
            const uint16_t* lut = slut + fc * MQC_LUT_STRIDE;
            const uint32_t swm = lut[48];
            const uint32_t npairs = 1u << lut[49];
            const double2 csv = scs[fc];
            const double c = csv.x, sn = csv.y;
            const uint32_t vab = (uint32_t)lut[lane] ^ (fixed_b ? (uint32_t)lut[32 + (warp & 7)] : 0u);
            double gacc = 0.0;
            for (uint32_t u = tid; u < npairs; u += nthr) {
                uint32_t val = vab ^ (uint32_t)lut[40u + (u >> 8)];
                if (!fixed_b) val ^= (uint32_t)lut[32u + ((u >> 5) & 7u)];
                const uint32_t x0 = val & 0x7fffu;
                const uint32_t x1 = x0 ^ swm;
                const bool neg = (val & 0x8000u) != 0u;
                const double ssn = neg ? -sn : sn;
                const double2 a = sp[x0], b = sp[x1];
                if (ADJ) {
                    const double2 la = sl[x0], lb = sl[x1];
                    const double gv = lb.x * a.x + lb.y * a.y - la.x * b.x - la.y * b.y;
                    gacc += neg ? -gv : gv;
                    sl[x0] = make_double2(c * la.x - ssn * lb.x, c * la.y - ssn * lb.y);
                    sl[x1] = make_double2(c * lb.x + ssn * la.x, c * lb.y + ssn * la.y);
                }
                sp[x0] = make_double2(c * a.x - ssn * b.x, c * a.y - ssn * b.y);
                sp[x1] = make_double2(c * b.x + ssn * a.x, c * b.y + ssn * a.y);
            }
            if (ADJ) {
                gacc = warp_sum(gacc);
                if (lane == 0) gpart[warp * MQC_FCHUNK + fc] = gacc;
            }
            __syncthreads();
        }
        if (ADJ) {
            for (int fc = tid; fc < nc; fc += nthr) {
                double g = 0.0;
                for (int w = 0; w < nw; ++w) g += gpart[w * MQC_FCHUNK + fc];
                const int th = stheta[fc];
                if (g != 0.0) atomicAdd(grad + (th & 0x3fffffff), (th >> 30) & 1 ? -2.0 * g : 2.0 * g);
            }
            __syncthreads();
        }
    }
    if (nf == 0) { cp_async_wait_all(); __syncthreads(); }

    // ---- store ----------------------------------------------------------------------
    for (uint32_t t = tid; t < T; t += nthr) {
        const u64 g = base | dep_lo[t & lo_mask] | dep_hi[t >> h0];
        const uint32_t sw = (uint32_t)sws_lo[t & lo_mask] ^ (uint32_t)sws_hi[t >> h0];
        st_stream(peer_psi(PR, g), sp[sw]);
        if (ADJ) st_stream(peer_lam(PR, g), sl[sw]);
    }
}

static inline size_t mqc_tile_smem_bytes(int K, int nf, bool adj, int nthreads) {
    (void)nf; (void)nthreads;
    size_t b = (size_t)(adj ? 2 : 1) * ((size_t)1 << K) * sizeof(double2);
    b += 2 * 128 * sizeof(u64);
    b += MQC_FCHUNK * sizeof(double2);
    if (adj) b += 16 * MQC_FCHUNK * sizeof(double);
    b += MQC_FCHUNK * sizeof(int);
    b += MQC_FCHUNK * MQC_LUT_STRIDE * sizeof(uint16_t);
    b += 4 * 128 * sizeof(uint16_t) + 32 * sizeof(int);
    return b;
}

// ------------------------------------------------------------------------------------------
// K1 / K3 : one factor per pass (physical coordinates)
// ------------------------------------------------------------------------------------------
__global__ void k_excite_direct(double2* __restrict__ psi, const MqcFactor F, const double2* __restrict__ cs,
                                const int reverse, const u64 npairs) {
    const double c = cs[F.slot].x;
    const double sn = reverse ? -cs[F.slot].y : cs[F.slot].y;
    u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; u < npairs; u += stride) {
        const u64 x0 = mqc_insert_zeros(u, F.pos, F.npos) | F.v;
        const u64 x1 = x0 ^ F.m;
        const double s = ((__popcll(x0 & F.z) + F.sign0) & 1) ? -1.0 : 1.0;
        const double ssn = s * sn;
        const double2 a = psi[x0], b = psi[x1];
        psi[x0] = make_double2(c * a.x - ssn * b.x, c * a.y - ssn * b.y);
        psi[x1] = make_double2(c * b.x + ssn * a.x, c * b.y + ssn * a.y);
    }
}

__global__ void k_adjoint_direct(double2* __restrict__ psi, double2* __restrict__ lam, const MqcFactor F,
                                 const double2* __restrict__ cs, double* __restrict__ grad, const u64 npairs) {
    __shared__ double red[32];
    const double c = cs[F.slot].x;
    const double sn = -cs[F.slot].y;
    u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double gacc = 0.0;
    for (; u < npairs; u += stride) {
        const u64 x0 = mqc_insert_zeros(u, F.pos, F.npos) | F.v;
        const u64 x1 = x0 ^ F.m;
        const double s = ((__popcll(x0 & F.z) + F.sign0) & 1) ? -1.0 : 1.0;
        const double ssn = s * sn;
        const double2 a = psi[x0], b = psi[x1], la = lam[x0], lb = lam[x1];
        gacc += s * (lb.x * a.x + lb.y * a.y - la.x * b.x - la.y * b.y);
        psi[x0] = make_double2(c * a.x - ssn * b.x, c * a.y - ssn * b.y);
        psi[x1] = make_double2(c * b.x + ssn * a.x, c * b.y + ssn * a.y);
        lam[x0] = make_double2(c * la.x - ssn * lb.x, c * la.y - ssn * lb.y);
        lam[x1] = make_double2(c * lb.x + ssn * la.x, c * lb.y + ssn * la.y);
    }
    gacc = warp_sum(gacc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gacc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0 && v != 0.0) atomicAdd(grad + F.theta, 2.0 * v);
    }
}

// ------------------------------------------------------------------------------------------
// K2 : lambda = H psi,  E = <psi|H|psi>
// ------------------------------------------------------------------------------------------
// (H psi)[y] = sum_g psi[y ^ m_g] * sum_{t in g} c_t (-1)^popc((y ^ m_g) & z_t)
// One thread per output amplitude y; with a sector, y runs over alpha-string x beta-string
// and partners outside the sector are skipped (they hold exact zeros).
__global__ void __launch_bounds__(256)
k_happly(const __grid_constant__ MqcPeers PR, double2* __restrict__ lam, const MqcHamDev H,
         const MqcSectorDev S, const u64 count, const u64 y_rank_bits, double* __restrict__ e_out) {
    __shared__ double red[2][8];
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    double er = 0.0, ei = 0.0;
    if (tid < count) {
        u64 y;
        if (S.enabled) {
            const u64 ia = tid / S.n_b, ib = tid - ia * S.n_b;
            y = S.astr[ia] | S.bstr[ib];
        } else {
            y = tid | y_rank_bits;          // this rank's shard
        }
        double ar = 0.0, ai = 0.0;
        for (int g = 0; g < H.n_groups; ++g) {
            const u64 x = y ^ __ldg(H.gmask + g);
            if (S.enabled && (__popcll(x & S.amask) != S.na || __popcll(x & S.bmask) != S.nb)) continue;
            const double2 p = *peer_psi(PR, x);
            if (p.x == 0.0 && p.y == 0.0) continue;
            const int t0 = __ldg(H.goff + g), t1 = __ldg(H.goff + g + 1);
            double cr = 0.0, ci = 0.0;
            for (int t = t0; t < t1; ++t) {
                const double2 cc = __ldg(H.tc + t);
                if (__popcll(x & __ldg(H.tz + t)) & 1) { cr -= cc.x; ci -= cc.y; }
                else { cr += cc.x; ci += cc.y; }
            }
            ar += cr * p.x - ci * p.y;
            ai += cr * p.y + ci * p.x;
        }
        if (lam) lam[y & ((1ull << PR.n_local) - 1ull)] = make_double2(ar, ai);
        const double2 py = *peer_psi(PR, y);
        er = py.x * ar + py.y * ai;       // conj(psi[y]) * acc
        ei = py.x * ai - py.y * ar;
    }
    er = warp_sum(er);
    ei = warp_sum(ei);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = er; red[1][warp] = ei; }
    __syncthreads();
    if (warp == 0) {
        double vr = (lane < (blockDim.x >> 5)) ? red[0][lane] : 0.0;
        double vi = (lane < (blockDim.x >> 5)) ? red[1][lane] : 0.0;
        vr = warp_sum(vr);
        vi = warp_sum(vi);
        if (lane == 0) {
            if (vr != 0.0) atomicAdd(e_out, vr);
            if (vi != 0.0) atomicAdd(e_out + 1, vi);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2 (sector form): C = psi restricted to alpha-string x beta-string, Lambda = H C
// ------------------------------------------------------------------------------------------
// The UCCSD state only populates one (N_alpha, N_beta) sector.  It is gathered into a dense
// matrix C[ia][ib]; X-mask groups are sorted by their alpha part so that a CTA, which owns one
// output alpha string, tests alpha validity once per distinct alpha mask (uniform branch) and
// reads whole rows of C; threads own output beta strings.
struct MqcSecHamDev {
    int n_ma;                  // distinct alpha x-masks (orbital space)
    const unsigned* ma_list;   // [n_ma]
    const int* ma_goff;        // [n_ma+1] group range per alpha mask
    const unsigned* gmb;       // [G] beta x-mask (orbital space)
    const int* goff;           // [G+1] term range per group
    const u64* tz;             // [T] physical z-mask
    const double2* tc;         // [T]
    // "simple" groups (<= 4 flipped bits, all terms share the parity mask outside them):
    //   coefficient(x) = (-1)^popc(x & zc) * W[pattern of x on the flipped bits]
    const unsigned* ginfo;     // [G] bit0 simple, bits 4..7 = # flipped alpha bits
    const unsigned* gzca;      // [G] common parity mask, alpha orbitals
    const unsigned* gzcb;      // [G] common parity mask, beta orbitals
    const double2* gw;         // [G][16] pattern = alpha bits (ascending) then beta bits
    const unsigned* aorb;      // [nA] alpha strings, orbital space
    const unsigned* borb;      // [nB]
    const u64* astr;           // [nA] physical images
    const u64* bstr;           // [nB]
    const unsigned* rank_a;    // [2^n_orb] string -> index (0xffffffff: not in sector)
    const unsigned* rank_b;
    unsigned n_a, n_b;
};

// rows [ia0, ia1) of the sector matrix; the state is read through the peer table, so on a
// sharded state every rank can assemble whatever rows it needs
// (C / L hold rows starting at row `buf_row0`)
__global__ void k_sector_gather(const __grid_constant__ MqcPeers PR, double2* __restrict__ C,
                                const u64* __restrict__ astr, const u64* __restrict__ bstr,
                                const unsigned ia0, const unsigned ia1, const unsigned n_b, const unsigned buf_row0) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (u64)(ia1 - ia0) * n_b) return;
    const unsigned ia = ia0 + (unsigned)(tid / n_b), ib = (unsigned)(tid % n_b);
    C[(u64)(ia - buf_row0) * n_b + ib] = *peer_psi(PR, astr[ia] | bstr[ib]);
}

__global__ void k_sector_scatter(const __grid_constant__ MqcPeers PR, const double2* __restrict__ L,
                                 const u64* __restrict__ astr, const u64* __restrict__ bstr,
                                 const unsigned ia0, const unsigned ia1, const unsigned n_b, const unsigned buf_row0) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (u64)(ia1 - ia0) * n_b) return;
    const unsigned ia = ia0 + (unsigned)(tid / n_b), ib = (unsigned)(tid % n_b);
    *peer_lam(PR, astr[ia] | bstr[ib]) = L[(u64)(ia - buf_row0) * n_b + ib];
}

// ------------------------------------------------------------------------------------------
// cross-rank plumbing of a sharded state (one process per GPU): barrier and sum over peers
// ------------------------------------------------------------------------------------------
struct MqcFlags {
    unsigned long long* flags[MQC_MAX_RANKS];   // flags[r][q]: rank q's arrival counter as seen by rank r
    int n_ranks, rank;
};
// All ranks launch this with the same epoch.  System-scope fence + release store publish
// everything this GPU wrote (to its own and to peers' HBM) before the arrival flag; the
// acquire spin returns once every peer has arrived.  Spins are bounded (~60 s) so that a dead
// peer becomes an error flag, not a hung GPU.
__global__ void k_rank_barrier(const __grid_constant__ MqcFlags F, const unsigned long long epoch,
                               int* __restrict__ err) {
    const int q = threadIdx.x;
    if (q >= F.n_ranks) return;
    __threadfence_system();
    unsigned long long* dst = F.flags[q] + F.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
    const unsigned long long* src = F.flags[F.rank] + q;
    const long long t0 = clock64();
    unsigned long long seen = 0;
    while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
        if (seen >= epoch) break;
        if (clock64() - t0 > 120000000000ll) { *err = 1; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

struct MqcSumPeers {
    const double* src[MQC_MAX_RANKS];
    int n_ranks;
};
__global__ void k_sum_peers(const __grid_constant__ MqcSumPeers S, double* __restrict__ out, const int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 0.0;
    for (int r = 0; r < S.n_ranks; ++r) v += S.src[r][i];
    out[i] = v;
}

#define MQC_HS_R 4
__device__ __forceinline__ unsigned pext_small(unsigned x, unsigned m) {   // gather the bits of x under m (<= 4 set bits)
    unsigned out = 0, k = 0;
    while (m) { const unsigned b = __ffs(m) - 1; m &= m - 1; out |= ((x >> b) & 1u) << k; ++k; }
    return out;
}

__global__ void __launch_bounds__(256)
k_happly_sector(const double2* __restrict__ C, double2* __restrict__ L, const MqcSecHamDev H,
                double* __restrict__ e_out, const unsigned ia0, const int accumulate) {
    __shared__ double red[2][8];
    const unsigned iap = ia0 + blockIdx.x;
    const unsigned ap = H.aorb[iap];
    unsigned bp[MQC_HS_R];
    unsigned ibp[MQC_HS_R];
    double ar[MQC_HS_R], ai[MQC_HS_R];
#pragma unroll
    for (int r = 0; r < MQC_HS_R; ++r) {
        ibp[r] = (blockIdx.y * MQC_HS_R + r) * blockDim.x + threadIdx.x;
        bp[r] = (ibp[r] < H.n_b) ? H.borb[ibp[r]] : 0xffffffffu;
        ar[r] = 0.0; ai[r] = 0.0;
    }
    for (int im = 0; im < H.n_ma; ++im) {
        const unsigned ma = __ldg(H.ma_list + im);
        if (2 * __popc(ap & ma) != __popc(ma)) continue;          // uniform over the CTA
        const unsigned a = ap ^ ma;                               // source alpha string
        const unsigned ia = __ldg(H.rank_a + a);
        const unsigned pata = pext_small(a, ma);
        const int nbits_a = __popc(ma);
        const double2* __restrict__ Crow = C + (size_t)ia * H.n_b;
        const int g0 = __ldg(H.ma_goff + im), g1 = __ldg(H.ma_goff + im + 1);
        for (int g = g0; g < g1; ++g) {
            const unsigned mb = __ldg(H.gmb + g);
            const int pmb = __popc(mb);
            const unsigned info = __ldg(H.ginfo + g);
            if (info & 1u) {
                // ---- LUT path ---------------------------------------------------------
                const unsigned sa = __popc(a & __ldg(H.gzca + g)) & 1u;
                const unsigned zcb = __ldg(H.gzcb + g);
                const double2* __restrict__ Wg = H.gw + (size_t)g * 16 + pata;
                if (pmb == 2) {
                    // exactly one of the two beta bits is set in a valid source string
                    const unsigned q0 = __ffs(mb) - 1;
                    const double2 w1 = __ldg(Wg + (1u << nbits_a)), w2 = __ldg(Wg + (2u << nbits_a));
#pragma unroll
                    for (int r = 0; r < MQC_HS_R; ++r) {
                        if (bp[r] != 0xffffffffu && __popc(bp[r] & mb) == 1) {
                            const unsigned b = bp[r] ^ mb;
                            const double2 amp = Crow[__ldg(H.rank_b + b)];
                            const bool first = (b >> q0) & 1u;
                            double wr = first ? w1.x : w2.x, wi = first ? w1.y : w2.y;
                            if ((__popc(b & zcb) + sa) & 1u) { wr = -wr; wi = -wi; }
                            ar[r] += wr * amp.x - wi * amp.y;
                            ai[r] += wr * amp.y + wi * amp.x;
                        }
                    }
                } else if (pmb == 0) {
                    const double2 w = __ldg(Wg);
#pragma unroll
                    for (int r = 0; r < MQC_HS_R; ++r) {
                        if (bp[r] != 0xffffffffu) {
                            const double2 amp = Crow[ibp[r]];
                            double wr = w.x, wi = w.y;
                            if ((__popc(bp[r] & zcb) + sa) & 1u) { wr = -wr; wi = -wi; }
                            ar[r] += wr * amp.x - wi * amp.y;
                            ai[r] += wr * amp.y + wi * amp.x;
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < MQC_HS_R; ++r) {
                        if (bp[r] != 0xffffffffu && 2 * __popc(bp[r] & mb) == pmb) {
                            const unsigned b = bp[r] ^ mb;
                            const double2 amp = Crow[__ldg(H.rank_b + b)];
                            const double2 w = __ldg(Wg + (pext_small(b, mb) << nbits_a));
                            double wr = w.x, wi = w.y;
                            if ((__popc(b & zcb) + sa) & 1u) { wr = -wr; wi = -wi; }
                            ar[r] += wr * amp.x - wi * amp.y;
                            ai[r] += wr * amp.y + wi * amp.x;
                        }
                    }
                }
                continue;
            }
            // ---- general path: explicit term loop -------------------------------------
            const u64 xa = __ldg(H.astr + ia);
            u64 x[MQC_HS_R];
            double2 amp[MQC_HS_R];
            double cr[MQC_HS_R], ci[MQC_HS_R];
            bool ok[MQC_HS_R];
            bool any = false;
#pragma unroll
            for (int r = 0; r < MQC_HS_R; ++r) {
                ok[r] = (bp[r] != 0xffffffffu) && (2 * __popc(bp[r] & mb) == pmb);
                cr[r] = 0.0; ci[r] = 0.0;
                if (ok[r]) {
                    const unsigned ib = __ldg(H.rank_b + (bp[r] ^ mb));
                    amp[r] = Crow[ib];
                    x[r] = xa | __ldg(H.bstr + ib);
                    ok[r] = (amp[r].x != 0.0) || (amp[r].y != 0.0);
                }
                any |= ok[r];
            }
            if (!any) continue;
            const int t0 = __ldg(H.goff + g), t1 = __ldg(H.goff + g + 1);
            for (int t = t0; t < t1; ++t) {
                const u64 z = __ldg(H.tz + t);
                const double2 cc = __ldg(H.tc + t);
#pragma unroll
                for (int r = 0; r < MQC_HS_R; ++r) {
                    if (ok[r]) {
                        if (__popcll(x[r] & z) & 1) { cr[r] -= cc.x; ci[r] -= cc.y; }
                        else { cr[r] += cc.x; ci[r] += cc.y; }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < MQC_HS_R; ++r) {
                if (ok[r]) {
                    ar[r] += cr[r] * amp[r].x - ci[r] * amp[r].y;
                    ai[r] += cr[r] * amp[r].y + ci[r] * amp[r].x;
                }
            }
        }
    }
    double er = 0.0, ei = 0.0;
#pragma unroll
    for (int r = 0; r < MQC_HS_R; ++r) {
        if (ibp[r] < H.n_b) {
            const size_t o = (size_t)iap * H.n_b + ibp[r];
            if (L) {
                if (accumulate) { const double2 o0 = L[o]; L[o] = make_double2(o0.x + ar[r], o0.y + ai[r]); }
                else L[o] = make_double2(ar[r], ai[r]);
            }
            const double2 py = C[o];
            er += py.x * ar[r] + py.y * ai[r];
            ei += py.x * ai[r] - py.y * ar[r];
        }
    }
    er = warp_sum(er);
    ei = warp_sum(ei);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = er; red[1][warp] = ei; }
    __syncthreads();
    if (warp == 0) {
        double vr = (lane < (blockDim.x >> 5)) ? red[0][lane] : 0.0;
        double vi = (lane < (blockDim.x >> 5)) ? red[1][lane] : 0.0;
        vr = warp_sum(vr);
        vi = warp_sum(vi);
        if (lane == 0) {
            if (vr != 0.0) atomicAdd(e_out, vr);
            if (vi != 0.0) atomicAdd(e_out + 1, vi);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4 : RDMs on determinant bit strings (physical indices)
// ------------------------------------------------------------------------------------------
struct MqcRdmDev {
    int n_orb;                 // spatial orbitals
    const u64* qbit;           // [2n] physical bit of qubit q
    const u64* below;          // [2n] physical mask of qubits < q
    int n_copies;              // privatised accumulator copies
};

// rdm1[a/2][i/2] += sign * Re(conj(psi[D]) psi[D'])   with D' = a^+ i D   (rdm.py:131-138)
__global__ void k_rdm1(const __grid_constant__ MqcPeers PR, const MqcSectorDev S, const MqcRdmDev R,
                       const u64 det0, const u64 det1, double* __restrict__ rdm1) {
    const u64 tid = det0 + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= det1) return;
    const u64 ia = tid / S.n_b, ib = tid - ia * S.n_b;
    const u64 y = S.astr[ia] | S.bstr[ib];
    const double2 c0 = *peer_psi(PR, y);
    if (c0.x == 0.0 && c0.y == 0.0) return;
    const int n = R.n_orb, nq = 2 * n;
    double* out = rdm1 + (size_t)(blockIdx.x % R.n_copies) * n * n;
    for (int i = 0; i < nq; ++i) {
        const u64 bi = R.qbit[i];
        if (!(y & bi)) continue;
        atomicAdd(out + (i >> 1) * n + (i >> 1), c0.x * c0.x + c0.y * c0.y);
        const u64 y1 = y ^ bi;
        const int par_i = __popcll(y & R.below[i]);
        for (int a = (i & 1); a < nq; a += 2) {
            const u64 ba = R.qbit[a];
            if (y1 & ba) continue;
            if (a == i) continue;
            const u64 y2 = y1 | ba;
            const double2 c1 = *peer_psi(PR, y2);
            if (c1.x == 0.0 && c1.y == 0.0) continue;
            const int par = par_i + __popcll(y1 & R.below[a]);
            const double v = c0.x * c1.x + c0.y * c1.y;
            atomicAdd(out + (a >> 1) * n + (i >> 1), (par & 1) ? -v : v);
        }
    }
}

// dm2[p,q,r,s] = sum_{sigma,tau} <p_sigma^+ r_tau^+ s_tau q_sigma>   (pyscf order, what
// make_rdm12 + reorder_rdm return, rdm.py:157-172).  One thread per (determinant, q_sigma).
__global__ void k_rdm2(const __grid_constant__ MqcPeers PR, const MqcSectorDev S, const MqcRdmDev R,
                       const u64 det0, const u64 det1, double* __restrict__ rdm2) {
    const int n = R.n_orb, nq = 2 * n;
    const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 det = det0 + gid / nq;
    const int Q = (int)(gid % nq);
    if (det >= det1) return;
    const u64 ia = det / S.n_b, ib = det - ia * S.n_b;
    const u64 y = S.astr[ia] | S.bstr[ib];
    const u64 bQ = R.qbit[Q];
    if (!(y & bQ)) return;
    const double2 c0 = *peer_psi(PR, y);
    if (c0.x == 0.0 && c0.y == 0.0) return;
    const size_t n2 = (size_t)n * n, n4 = n2 * n2;
    double* out = rdm2 + (size_t)(blockIdx.x % R.n_copies) * n4;
    const int par_q = __popcll(y & R.below[Q]);
    const u64 y1 = y ^ bQ;
    for (int Sx = 0; Sx < nq; ++Sx) {
        const u64 bS = R.qbit[Sx];
        if (!(y1 & bS)) continue;
        const int par_s = par_q + __popcll(y1 & R.below[Sx]);
        const u64 y2 = y1 ^ bS;
        for (int Rr = (Sx & 1); Rr < nq; Rr += 2) {
            const u64 bR = R.qbit[Rr];
            if (y2 & bR) continue;
            const int par_r = par_s + __popcll(y2 & R.below[Rr]);
            const u64 y3 = y2 | bR;
            for (int Pp = (Q & 1); Pp < nq; Pp += 2) {
                const u64 bP = R.qbit[Pp];
                if (y3 & bP) continue;
                const u64 y4 = y3 | bP;
                const double2 c1 = *peer_psi(PR, y4);
                if (c1.x == 0.0 && c1.y == 0.0) continue;
                const int par = par_r + __popcll(y3 & R.below[Pp]);
                const double v = c1.x * c0.x + c1.y * c0.y;       // Re(conj(psi[D']) psi[D])
                atomicAdd(out + (((size_t)(Pp >> 1) * n + (Q >> 1)) * n + (Rr >> 1)) * n + (Sx >> 1),
                          (par & 1) ? -v : v);
            }
        }
    }
}

__global__ void k_reduce_copies(double* __restrict__ buf, const size_t len, const int n_copies) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double v = 0.0;
    for (int c = 0; c < n_copies; ++c) v += buf[(size_t)c * len + i];
    buf[i] = v;
}
