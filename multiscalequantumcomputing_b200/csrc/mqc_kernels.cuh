// sm_100a kernels of the UCCSD state-vector path.  All HBM-bandwidth-bound complex128
// streaming work: no tensor cores (nothing here is a dense contraction).
//
//   k_tile_sweep<ADJ>   K1/K3  fused excitation sweep over a 2^K shared-memory tile
//                              (ADJ: psi and lambda together + per-factor gradient partials)
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
__global__ void k_sincos(const double* __restrict__ theta, const int* __restrict__ theta_index,
                         double2* __restrict__ cs, int n) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) {
        double s, c;
        sincos(theta[theta_index[f]], &s, &c);
        cs[f] = make_double2(c, s);
    }
}

__global__ void k_set_amplitude(double2* psi, u64 index, double re, double im) {
    if (blockIdx.x == 0 && threadIdx.x == 0) psi[index] = make_double2(re, im);
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
template <bool ADJ>
__global__ void __launch_bounds__(512)
k_tile_sweep(double2* __restrict__ psi, double2* __restrict__ lam, const MqcPass P,
             const MqcTileFactor* __restrict__ tf, const double2* __restrict__ cs,
             double* __restrict__ grad, const int n_bits, const int reverse) {
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
    uint16_t* prm_lo = reinterpret_cast<uint16_t*>(q);  q += 128 * sizeof(uint16_t);
    uint16_t* prm_hi = reinterpret_cast<uint16_t*>(q);  q += 128 * sizeof(uint16_t);
    MqcTileFactor* sf = reinterpret_cast<MqcTileFactor*>(q);  q += (size_t)nf * sizeof(MqcTileFactor);
    double2* scs = reinterpret_cast<double2*>(q);       q += (size_t)nf * sizeof(double2);
    double* gpart = reinterpret_cast<double*>(q);

    // ---- per-CTA tables -------------------------------------------------------------
    for (int t = tid; t < (1 << h0); t += nthr) {
        dep_lo[t] = mqc_deposit((u64)t, P.pos, h0);
        prm_lo[t] = (uint16_t)mqc_permute_bits((uint32_t)t, P.perm, h0);
    }
    for (int t = tid; t < (1 << h1); t += nthr) {
        dep_hi[t] = mqc_deposit((u64)t, P.pos + h0, h1);
        prm_hi[t] = (uint16_t)mqc_permute_bits((uint32_t)t, P.perm + h0, h1);
    }
    for (int f = tid; f < nf; f += nthr) {
        const MqcTileFactor F = tf[P.f0 + f];
        sf[f] = F;
        scs[f] = cs[F.slot];
    }
    const u64 base = mqc_deposit_complement((u64)blockIdx.x, P.act_mask, n_bits);
    __syncthreads();

    // ---- load -----------------------------------------------------------------------
    const bool perm_load = reverse && P.has_perm;
    const bool perm_store = (!reverse) && P.has_perm;
    for (uint32_t t = tid; t < T; t += nthr) {
        const u64 g = base | dep_lo[t & lo_mask] | dep_hi[t >> h0];
        const uint32_t st = perm_load ? (uint32_t)(prm_lo[t & lo_mask] | prm_hi[t >> h0]) : t;
        sp[st] = ld_stream(psi + g);
        if (ADJ) sl[st] = ld_stream(lam + g);
    }
    __syncthreads();

    // ---- factors --------------------------------------------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    for (int fi = 0; fi < nf; ++fi) {
        const int f = reverse ? (nf - 1 - fi) : fi;
        const MqcTileFactor F = sf[f];
        const double c = scs[f].x;
        const double sn = reverse ? -scs[f].y : scs[f].y;
        const int sgn_tile = (__popcll(base & F.z_out) + F.sign0) & 1;
        const uint32_t npairs = T >> F.npos;
        double gacc = 0.0;
        for (uint32_t u = tid; u < npairs; u += nthr) {
            const uint32_t x0 = mqc_insert_zeros8(u, F.pos, F.npos) | F.v;
            const uint32_t x1 = x0 ^ F.m;
            const double s = ((__popc(x0 & F.z) + sgn_tile) & 1) ? -1.0 : 1.0;
            const double ssn = s * sn;
            const double2 a = sp[x0], b = sp[x1];
            if (ADJ) {
                const double2 la = sl[x0], lb = sl[x1];
                gacc += s * (lb.x * a.x + lb.y * a.y - la.x * b.x - la.y * b.y);
                sl[x0] = make_double2(c * la.x - ssn * lb.x, c * la.y - ssn * lb.y);
                sl[x1] = make_double2(c * lb.x + ssn * la.x, c * lb.y + ssn * la.y);
            }
            sp[x0] = make_double2(c * a.x - ssn * b.x, c * a.y - ssn * b.y);
            sp[x1] = make_double2(c * b.x + ssn * a.x, c * b.y + ssn * a.y);
        }
        if (ADJ) {
            gacc = warp_sum(gacc);
            if (lane == 0) gpart[warp * nf + f] = gacc;
        }
        __syncthreads();
    }
    if (ADJ) {
        const int nw = nthr >> 5;
        for (int f = tid; f < nf; f += nthr) {
            double g = 0.0;
            for (int w = 0; w < nw; ++w) g += gpart[w * nf + f];
            if (g != 0.0) atomicAdd(grad + sf[f].theta, 2.0 * g);
        }
    }

    // ---- store ----------------------------------------------------------------------
    for (uint32_t t = tid; t < T; t += nthr) {
        const u64 g = base | dep_lo[t & lo_mask] | dep_hi[t >> h0];
        const uint32_t st = perm_store ? (uint32_t)(prm_lo[t & lo_mask] | prm_hi[t >> h0]) : t;
        st_stream(psi + g, sp[st]);
        if (ADJ) st_stream(lam + g, sl[st]);
    }
}

static inline size_t mqc_tile_smem_bytes(int K, int nf, bool adj, int nthreads) {
    size_t b = (size_t)(adj ? 2 : 1) * ((size_t)1 << K) * sizeof(double2);
    b += 2 * 128 * sizeof(u64) + 2 * 128 * sizeof(uint16_t);
    b += (size_t)nf * (sizeof(MqcTileFactor) + sizeof(double2));
    if (adj) b += (size_t)(nthreads / 32) * nf * sizeof(double);
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
k_happly(const double2* __restrict__ psi, double2* __restrict__ lam, const MqcHamDev H,
         const MqcSectorDev S, const u64 count, double* __restrict__ e_out) {
    __shared__ double red[2][8];
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    double er = 0.0, ei = 0.0;
    if (tid < count) {
        u64 y;
        if (S.enabled) {
            const u64 ia = tid / S.n_b, ib = tid - ia * S.n_b;
            y = S.astr[ia] | S.bstr[ib];
        } else {
            y = tid;
        }
        double ar = 0.0, ai = 0.0;
        for (int g = 0; g < H.n_groups; ++g) {
            const u64 x = y ^ __ldg(H.gmask + g);
            if (S.enabled && (__popcll(x & S.amask) != S.na || __popcll(x & S.bmask) != S.nb)) continue;
            const double2 p = psi[x];
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
        if (lam) lam[y] = make_double2(ar, ai);
        const double2 py = psi[y];
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
// K4 : RDMs on determinant bit strings (physical indices)
// ------------------------------------------------------------------------------------------
struct MqcRdmDev {
    int n_orb;                 // spatial orbitals
    const u64* qbit;           // [2n] physical bit of qubit q
    const u64* below;          // [2n] physical mask of qubits < q
    int n_copies;              // privatised accumulator copies
};

// rdm1[a/2][i/2] += sign * Re(conj(psi[D]) psi[D'])   with D' = a^+ i D   (rdm.py:131-138)
__global__ void k_rdm1(const double2* __restrict__ psi, const MqcSectorDev S, const MqcRdmDev R,
                       const u64 count, double* __restrict__ rdm1) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= count) return;
    const u64 ia = tid / S.n_b, ib = tid - ia * S.n_b;
    const u64 y = S.astr[ia] | S.bstr[ib];
    const double2 c0 = psi[y];
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
            const double2 c1 = psi[y2];
            if (c1.x == 0.0 && c1.y == 0.0) continue;
            const int par = par_i + __popcll(y1 & R.below[a]);
            const double v = c0.x * c1.x + c0.y * c1.y;
            atomicAdd(out + (a >> 1) * n + (i >> 1), (par & 1) ? -v : v);
        }
    }
}

// dm2[p,q,r,s] = sum_{sigma,tau} <p_sigma^+ r_tau^+ s_tau q_sigma>   (pyscf order, what
// make_rdm12 + reorder_rdm return, rdm.py:157-172).  One thread per (determinant, q_sigma).
__global__ void k_rdm2(const double2* __restrict__ psi, const MqcSectorDev S, const MqcRdmDev R,
                       const u64 count, double* __restrict__ rdm2) {
    const int n = R.n_orb, nq = 2 * n;
    const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 det = gid / nq;
    const int Q = (int)(gid - det * nq);
    if (det >= count) return;
    const u64 ia = det / S.n_b, ib = det - ia * S.n_b;
    const u64 y = S.astr[ia] | S.bstr[ib];
    const u64 bQ = R.qbit[Q];
    if (!(y & bQ)) return;
    const double2 c0 = psi[y];
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
                const double2 c1 = psi[y4];
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
