// C ABI (include/mqc_b200.h) on top of the sm_100a kernels.  Host logic only in this file:
// factor/mask construction, schedule building, device table upload, launch sequencing.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/mqc_b200.h"
#include "mqc_common.h"
#include "mqc_kernels.cuh"
#include "mqc_schedule.h"

static thread_local std::string g_err;

#define MQC_TRY try {
#define MQC_CATCH                                                   \
    }                                                               \
    catch (const std::exception& e) { g_err = e.what(); return 1; } \
    catch (...) { g_err = "unknown error"; return 1; }              \
    return 0;

struct MqcError : std::runtime_error { using std::runtime_error::runtime_error; };

static void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw MqcError(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
#define CK(x) cuda_check((x), #x)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        if (count) CK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void upload(const std::vector<T>& h, cudaStream_t st) {
        if (h.size() != n || !p) alloc(h.size());
        if (n) CK(cudaMemcpyAsync(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
};

struct HamHost {
    std::vector<u64> x, z;          // logical masks
    std::vector<double2> c;
};

struct HamTables {                  // one Hamiltonian mapped to one physical layout
    DevBuf<u64> gmask, tz;
    DevBuf<int> goff;
    DevBuf<double2> tc;
    int n_groups = 0;
    MqcHamDev dev() const { return MqcHamDev{n_groups, gmask.p, goff.p, tz.p, tc.p}; }
    void reset() { gmask.release(); tz.release(); goff.release(); tc.release(); n_groups = 0; }
};

struct SecHamTables {               // Hamiltonian in sector form for one layout and (na, nb)
    DevBuf<unsigned> ma_list, gmb, ginfo, gzca, gzcb;
    DevBuf<int> ma_goff, goff;
    DevBuf<u64> tz;
    DevBuf<double2> tc, gw;
    int n_ma = 0;
    void reset() {
        ma_list.release(); gmb.release(); ma_goff.release(); goff.release(); tz.release(); tc.release();
        ginfo.release(); gzca.release(); gzcb.release(); gw.release(); n_ma = 0;
    }
};

struct LayoutTables {
    std::vector<int> phys;          // logical bit -> physical position
    HamTables ham;
    SecHamTables sham;
    bool sham_ready = false;
    DevBuf<u64> astr, bstr, qbit, below;
    DevBuf<unsigned> aorb, borb, rank_a, rank_b;
    DevBuf<int> phys_dev;
    u64 amask = 0, bmask = 0;
    int na = -1, nb = -1;           // sector the strings were built for
    bool ham_ready = false;
    void reset() {
        phys.clear(); ham.reset(); sham.reset(); astr.release(); bstr.release(); qbit.release(); below.release();
        aorb.release(); borb.release(); rank_a.release(); rank_b.release();
        phys_dev.release(); amask = bmask = 0; na = nb = -1; ham_ready = false; sham_ready = false;
    }
};

struct mqc_solver {
    int n = 0, device = 0;
    u64 dim = 0;                    // amplitudes held by THIS handle (2^n_local)
    int n_global = 0, rank = 0;     // sharded state: top n_global physical bits = rank id
    int n_local = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t user_ev[2] = {nullptr, nullptr};
    bool gpu_ready = false;

    // options
    int tile_bits = 12, low_bits = 3, grad_tile_bits = 12, mode = 1, use_sector = 1, tile_threads = 256, sparse_tiles = 1, sparse_threads = 128;

    DevBuf<double2> psi, lam, cs, scratch_state, csec, lsec;
    DevBuf<double> theta, grad, scalars, rdm_acc;
    DevBuf<int> theta_index;
    DevBuf<MqcTileFactor> tf[2];

    HamHost ham;
    bool have_ham = false, have_ansatz = false;

    std::vector<MqcFactor> fac;     // logical coordinates
    std::vector<int> fac_theta;
    int n_theta = 0;
    u64 ref_index = 0;
    bool conserving = false;
    int na = 0, nb = 0;

    MqcSchedule sched[2];
    LayoutTables lay[2];

    std::vector<double> theta_host;
    int state_layout = -1;          // which layout psi currently holds (-1: invalid)
    double ms[4] = {0, 0, 0, 0};
    int64_t counters[3] = {0, 0, 0};
};

// ---------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------
static int ladder(u64& D, int p, int create) {
    const int par = __builtin_popcountll(D & ((1ull << p) - 1ull)) & 1;
    const bool occ = (D >> p) & 1ull;
    if (create) { if (occ) return -1; D |= 1ull << p; }
    else { if (!occ) return -1; D &= ~(1ull << p); }
    return par;
}

static u64 between_mask_qubits(int p, int q) {      // qubit-number bits strictly between
    const int lo = std::min(p, q), hi = std::max(p, q);
    if (hi - lo < 2) return 0;
    return ((1ull << hi) - 1ull) & ~((1ull << (lo + 1)) - 1ull);
}

static u64 qubit_to_logical(u64 qmask, int n) {
    u64 out = 0;
    for (int q = 0; q < n; ++q)
        if ((qmask >> q) & 1ull) out |= 1ull << (n - 1 - q);
    return out;
}

static MqcFactor make_factor(int n, int kind, const int32_t* ix, int slot, int theta) {
    MqcFactor f{};
    u64 mq, vq, zq;
    int sign0 = 0;
    auto chk = [&](int q) { if (q < 0 || q >= n) throw MqcError("ansatz index out of range"); };
    if (kind == 1) {
        const int i = ix[0], a = ix[1];
        chk(i); chk(a);
        if (i == a) throw MqcError("single excitation with i == a");
        mq = (1ull << i) | (1ull << a);
        vq = 1ull << i;
        zq = between_mask_qubits(i, a);
        u64 D = 1ull << i;
        int p1 = ladder(D, i, 0), p2 = ladder(D, a, 1);
        sign0 = (p1 + p2) & 1;
    } else if (kind == 2) {
        const int i = ix[0], j = ix[1], a = ix[2], b = ix[3];
        chk(i); chk(j); chk(a); chk(b);
        mq = (1ull << i) | (1ull << j) | (1ull << a) | (1ull << b);
        if (__builtin_popcountll(mq) != 4) throw MqcError("double excitation needs four distinct orbitals");
        vq = (1ull << i) | (1ull << j);
        zq = (between_mask_qubits(i, j) ^ between_mask_qubits(a, b)) & ~mq;
        u64 D = vq;
        int p1 = ladder(D, i, 0), p2 = ladder(D, j, 0), p3 = ladder(D, b, 1), p4 = ladder(D, a, 1);
        sign0 = (p1 + p2 + p3 + p4) & 1;
    } else {
        throw MqcError("unknown factor kind (1 = single, 2 = double)");
    }
    f.m = qubit_to_logical(mq, n);
    f.v = qubit_to_logical(vq, n);
    f.z = qubit_to_logical(zq, n);
    f.npos = __builtin_popcountll(f.m);
    int c = 0;
    for (int b = 0; b < n; ++b) if ((f.m >> b) & 1ull) f.pos[c++] = b;
    f.sign0 = sign0;
    f.slot = slot;
    f.theta = theta;
    return f;
}

static MqcFactor map_factor(const MqcFactor& f, const std::vector<int>& phys) {
    MqcFactor g = f;
    g.m = mqc_map_mask(f.m, phys);
    g.v = mqc_map_mask(f.v, phys);
    g.z = mqc_map_mask(f.z, phys);
    int c = 0;
    for (int b = 0; b < (int)phys.size(); ++b) if ((g.m >> b) & 1ull) g.pos[c++] = b;
    return g;
}

static void require_gpu(mqc_solver* s) {
    if (!s) throw MqcError("null solver handle");
    if (s->gpu_ready) { CK(cudaSetDevice(s->device)); return; }
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt <= 0)
        throw MqcError("no CUDA device available: the B200 kernels are required (there is no CPU fallback)");
    if (s->device >= cnt) throw MqcError("device index out of range");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto& e2 : s->ev) CK(cudaEventCreate(&e2));
    CK(cudaFuncSetAttribute(k_tile_sweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CK(cudaFuncSetAttribute(k_tile_sweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CK(cudaFuncSetAttribute(k_tile_sparse<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_tile_sparse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    s->gpu_ready = true;
}

static void build_ham_tables(mqc_solver* s, const HamHost& H, const std::vector<int>& phys, HamTables& T) {
    const size_t nt = H.x.size();
    std::vector<size_t> order(nt);
    std::iota(order.begin(), order.end(), 0);
    std::vector<u64> px(nt), pz(nt);
    for (size_t t = 0; t < nt; ++t) { px[t] = mqc_map_mask(H.x[t], phys); pz[t] = mqc_map_mask(H.z[t], phys); }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return px[a] < px[b]; });
    std::vector<u64> gmask, tz(nt);
    std::vector<int> goff;
    std::vector<double2> tc(nt);
    for (size_t k = 0; k < nt; ++k) {
        const size_t t = order[k];
        if (k == 0 || px[t] != gmask.back()) { gmask.push_back(px[t]); goff.push_back((int)k); }
        tz[k] = pz[t];
        tc[k] = H.c[t];
    }
    goff.push_back((int)nt);
    T.n_groups = (int)gmask.size();
    T.gmask.upload(gmask, s->stream);
    T.goff.upload(goff, s->stream);
    T.tz.upload(tz, s->stream);
    T.tc.upload(tc, s->stream);
    CK(cudaStreamSynchronize(s->stream));
}

static void enum_strings(int n_orb, int k, int spin, int n, const std::vector<int>& phys, std::vector<u64>& out,
                         std::vector<unsigned>& orb, std::vector<unsigned>& rank) {
    out.clear(); orb.clear();
    rank.assign((size_t)1 << n_orb, 0xffffffffu);
    if (k < 0 || k > n_orb) return;
    if (k == 0) { out.push_back(0); orb.push_back(0); rank[0] = 0; return; }
    u64 v = (1ull << k) - 1ull;
    const u64 limit = 1ull << n_orb;
    while (v < limit) {
        u64 img = 0;
        for (int p = 0; p < n_orb; ++p)
            if ((v >> p) & 1ull) img |= 1ull << phys[n - 1 - (2 * p + spin)];
        rank[v] = (unsigned)out.size();
        out.push_back(img);
        orb.push_back((unsigned)v);
        const u64 c = v & (~v + 1ull), r = v + c;
        v = (((r ^ v) >> 2) / c) | r;
    }
}

static void build_sector_tables(mqc_solver* s, LayoutTables& L, int na, int nb) {
    if (L.na == na && L.nb == nb && L.astr.p) return;
    const int n = s->n, n_orb = n / 2;
    std::vector<u64> a, b;
    std::vector<unsigned> ao, bo, ra, rb;
    enum_strings(n_orb, na, 0, n, L.phys, a, ao, ra);
    enum_strings(n_orb, nb, 1, n, L.phys, b, bo, rb);
    L.astr.upload(a, s->stream);
    L.bstr.upload(b, s->stream);
    L.aorb.upload(ao, s->stream);
    L.borb.upload(bo, s->stream);
    L.rank_a.upload(ra, s->stream);
    L.rank_b.upload(rb, s->stream);
    L.sham_ready = false;
    L.amask = L.bmask = 0;
    std::vector<u64> qbit(n), below(n);
    u64 acc = 0;
    for (int q = 0; q < n; ++q) {
        const u64 bit = 1ull << L.phys[n - 1 - q];
        qbit[q] = bit;
        below[q] = acc;
        acc |= bit;
        if (q & 1) L.bmask |= bit; else L.amask |= bit;
    }
    L.qbit.upload(qbit, s->stream);
    L.below.upload(below, s->stream);
    CK(cudaStreamSynchronize(s->stream));
    L.na = na; L.nb = nb;
}

static void build_sec_ham(mqc_solver* s, const HamHost& H, const std::vector<int>& phys, SecHamTables& T) {
    const int n = s->n;
    std::vector<int> orb_of(n), spin_of(n);
    for (int q = 0; q < n; ++q) { orb_of[phys[n - 1 - q]] = q / 2; spin_of[phys[n - 1 - q]] = q & 1; }
    struct Row { unsigned ma, mb; u64 z; double2 c; u64 px; };
    std::vector<u64> abit(n / 2 + 1, 0), bbit(n / 2 + 1, 0);           // physical bit of alpha / beta orbital p
    for (int q = 0; q < n; ++q) { if (q & 1) bbit[q / 2] = 1ull << phys[n - 1 - q]; else abit[q / 2] = 1ull << phys[n - 1 - q]; }
    std::vector<Row> rows;
    rows.reserve(H.x.size());
    for (size_t t = 0; t < H.x.size(); ++t) {
        u64 px = mqc_map_mask(H.x[t], phys);
        unsigned ma = 0, mb = 0;
        while (px) {
            const int b = __builtin_ctzll(px); px &= px - 1;
            if (spin_of[b]) mb |= 1u << orb_of[b]; else ma |= 1u << orb_of[b];
        }
        if ((__builtin_popcount(ma) | __builtin_popcount(mb)) & 1) continue;   // leaves the sector
        rows.push_back(Row{ma, mb, mqc_map_mask(H.z[t], phys), H.c[t], mqc_map_mask(H.x[t], phys)});
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
        return a.ma != b.ma ? a.ma < b.ma : a.mb < b.mb; });
    std::vector<unsigned> ma_list, gmb;
    std::vector<int> ma_goff, goff;
    std::vector<u64> tz(rows.size());
    std::vector<double2> tc(rows.size());
    for (size_t k = 0; k < rows.size(); ++k) {
        const bool new_ma = (k == 0) || rows[k].ma != rows[k - 1].ma;
        const bool new_g = new_ma || rows[k].mb != rows[k - 1].mb;
        if (new_g) {
            if (new_ma) { ma_list.push_back(rows[k].ma); ma_goff.push_back((int)gmb.size()); }
            gmb.push_back(rows[k].mb);
            goff.push_back((int)k);
        }
        tz[k] = rows[k].z;
        tc[k] = rows[k].c;
    }
    ma_goff.push_back((int)gmb.size());
    goff.push_back((int)rows.size());
    // LUT form of the "simple" groups
    const size_t G = gmb.size();
    std::vector<unsigned> ginfo(std::max<size_t>(G, 1), 0), gzca(std::max<size_t>(G, 1), 0), gzcb(std::max<size_t>(G, 1), 0);
    std::vector<double2> gw(std::max<size_t>(G, 1) * 16, make_double2(0, 0));
    for (size_t g = 0; g < G; ++g) {
        const int t0 = goff[g], t1 = goff[g + 1];
        const unsigned ma = rows[t0].ma, mb = rows[t0].mb;
        const int na_bits = __builtin_popcount(ma), nb_bits = __builtin_popcount(mb);
        if (na_bits + nb_bits > 4) continue;
        const u64 mphys = rows[t0].px;
        const u64 zc = rows[t0].z & ~mphys;
        bool simple = true;
        for (int t = t0; t < t1; ++t) if ((rows[t].z & ~mphys) != zc) { simple = false; break; }
        if (!simple) continue;
        // physical bit of every pattern position: alpha orbitals ascending, then beta
        u64 pbit[4]; int np = 0;
        for (int p = 0; p < n / 2 + 1; ++p) if ((ma >> p) & 1u) pbit[np++] = abit[p];
        for (int p = 0; p < n / 2 + 1; ++p) if ((mb >> p) & 1u) pbit[np++] = bbit[p];
        for (int pat = 0; pat < (1 << np); ++pat) {
            u64 xm = 0;
            for (int k = 0; k < np; ++k) if ((pat >> k) & 1) xm |= pbit[k];
            double wr = 0, wi = 0;
            for (int t = t0; t < t1; ++t) {
                const double sgn = (__builtin_popcountll(xm & rows[t].z) & 1) ? -1.0 : 1.0;
                wr += sgn * rows[t].c.x; wi += sgn * rows[t].c.y;
            }
            gw[g * 16 + pat] = make_double2(wr, wi);
        }
        unsigned za = 0, zb = 0;
        for (int p = 0; p < n / 2 + 1; ++p) { if (zc & abit[p]) za |= 1u << p; if (zc & bbit[p]) zb |= 1u << p; }
        gzca[g] = za; gzcb[g] = zb;
        ginfo[g] = 1u | ((unsigned)na_bits << 4);
    }
    T.ginfo.upload(ginfo, s->stream);
    T.gzca.upload(gzca, s->stream);
    T.gzcb.upload(gzcb, s->stream);
    T.gw.upload(gw, s->stream);
    if (gmb.empty()) gmb.push_back(0);
    if (ma_list.empty()) ma_list.push_back(0);
    if (tz.empty()) { tz.push_back(0); tc.push_back(make_double2(0, 0)); }
    T.n_ma = (int)ma_goff.size() - 1;
    T.ma_list.upload(ma_list, s->stream);
    T.ma_goff.upload(ma_goff, s->stream);
    T.gmb.upload(gmb, s->stream);
    T.goff.upload(goff, s->stream);
    T.tz.upload(tz, s->stream);
    T.tc.upload(tc, s->stream);
    CK(cudaStreamSynchronize(s->stream));
}

static MqcSecHamDev sec_ham_dev(const LayoutTables& L, const SecHamTables& T) {
    MqcSecHamDev D{};
    D.n_ma = T.n_ma;
    D.ma_list = T.ma_list.p; D.ma_goff = T.ma_goff.p; D.gmb = T.gmb.p; D.goff = T.goff.p;
    D.tz = T.tz.p; D.tc = T.tc.p;
    D.ginfo = T.ginfo.p; D.gzca = T.gzca.p; D.gzcb = T.gzcb.p; D.gw = T.gw.p;
    D.aorb = L.aorb.p; D.borb = L.borb.p; D.astr = L.astr.p; D.bstr = L.bstr.p;
    D.rank_a = L.rank_a.p; D.rank_b = L.rank_b.p;
    D.n_a = (unsigned)L.astr.n; D.n_b = (unsigned)L.bstr.n;
    return D;
}

static MqcSectorDev sector_dev(const LayoutTables& L, bool enabled) {
    MqcSectorDev S{};
    S.enabled = enabled ? 1 : 0;
    S.na = L.na; S.nb = L.nb;
    S.n_a = L.astr.n; S.n_b = L.bstr.n;
    S.astr = L.astr.p; S.bstr = L.bstr.p;
    S.amask = L.amask; S.bmask = L.bmask;
    return S;
}

static void ensure_layout_ready(mqc_solver* s, int which) {
    LayoutTables& L = s->lay[which];
    if (!s->have_ham) throw MqcError("mqc_set_hamiltonian has not been called");
    if (!L.ham_ready && !(s->conserving && s->use_sector)) {
        build_ham_tables(s, s->ham, L.phys, L.ham);
        L.ham_ready = true;
    }
    if (s->conserving && s->use_sector) {
        build_sector_tables(s, L, s->na, s->nb);
        if (!L.sham_ready) { build_sec_ham(s, s->ham, L.phys, L.sham); L.sham_ready = true; }
    }
}

// ---------------------------------------------------------------------------------------
// launch sequences
// ---------------------------------------------------------------------------------------
static double binom(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

static void launch_tile(mqc_solver* s, int which, int pass, bool adj, int reverse) {
    const MqcSchedule& S = s->sched[which];
    const MqcPass& P = S.passes[pass];
    const int nf = P.f1 - P.f0;
    const unsigned grid = (unsigned)(s->dim >> P.k);
    const u64 rank_bits = (u64)s->rank << s->n_local;
    const bool sparse = s->conserving && s->use_sector && s->sparse_tiles;
    if (sparse) {
        // alpha = even qubits, beta = odd qubits.  Tile-local coordinates are those of the
        // forward pass's LOAD layout in both directions (the reverse pass un-permutes while
        // loading), and a pass never moves the inactive bits, so one set of masks serves both.
        MqcSparse SP{};
        const std::vector<int>& ph = S.pass_phys_before[pass];
        for (int q = 0; q < s->n; ++q) {
            const u64 bit = 1ull << ph[s->n - 1 - q];
            if (q & 1) SP.bmask |= bit; else SP.amask |= bit;
        }
        SP.na = s->na; SP.nb = s->nb;
        const int n_orb_a = (s->n + 1) / 2, n_orb_b = s->n / 2;
        const int a_act = __builtin_popcountll(SP.amask & P.act_mask), b_act = __builtin_popcountll(SP.bmask & P.act_mask);
        double ca = 0, cb = 0;
        for (int k = 0; k <= a_act; ++k) if (k <= s->na && s->na - k <= n_orb_a - a_act) ca = std::max(ca, binom(a_act, k));
        for (int k = 0; k <= b_act; ++k) if (k <= s->nb && s->nb - k <= n_orb_b - b_act) cb = std::max(cb, binom(b_act, k));
        SP.cap = std::max(1, (int)(ca * cb + 0.5));
        int nthr = s->sparse_threads;
        while (nthr > 32 && nthr / 2 >= SP.cap) nthr >>= 1;
        const size_t smem = mqc_sparse_smem_bytes(P.k, SP.cap, adj);
        if (smem <= 200000) {
            if (adj) k_tile_sparse<true><<<grid, nthr, smem, s->stream>>>(s->psi.p, s->lam.p, P, s->tf[which].p, s->cs.p, s->grad.p, s->n_local, reverse, SP, rank_bits);
            else k_tile_sparse<false><<<grid, nthr, smem, s->stream>>>(s->psi.p, nullptr, P, s->tf[which].p, s->cs.p, nullptr, s->n_local, reverse, SP, rank_bits);
            s->counters[0]++;
            return;
        }
    }
    int nthr = s->tile_threads;
    const int T = 1 << P.k;
    while (nthr > 32 && nthr > T / 2) nthr >>= 1;
    const size_t smem = mqc_tile_smem_bytes(P.k, nf, adj, nthr);
    if (smem > 232448) throw MqcError("tile does not fit shared memory; lower tile_bits / grad_tile_bits");
    if (adj)
        k_tile_sweep<true><<<grid, nthr, smem, s->stream>>>(s->psi.p, s->lam.p, P, s->tf[which].p, s->cs.p, s->grad.p, s->n_local, reverse, rank_bits);
    else
        k_tile_sweep<false><<<grid, nthr, smem, s->stream>>>(s->psi.p, nullptr, P, s->tf[which].p, s->cs.p, nullptr, s->n_local, reverse, rank_bits);
    s->counters[0]++;
}

static void forward(mqc_solver* s, int which, const double* theta) {
    if (!s->have_ansatz) throw MqcError("mqc_set_ansatz has not been called");
    if (s->n_global) throw MqcError("sharded handle: drive it through the mqc_shard_* entry points");
    if (theta) {
        s->theta_host.assign(theta, theta + s->n_theta);
        CK(cudaMemcpyAsync(s->theta.p, theta, sizeof(double) * s->n_theta, cudaMemcpyHostToDevice, s->stream));
    }
    const int nfac = (int)s->fac.size();
    if (nfac) { k_sincos<<<(nfac + 255) / 256, 256, 0, s->stream>>>(s->theta.p, s->theta_index.p, s->cs.p, nfac); s->counters[0]++; }
    CK(cudaMemsetAsync(s->psi.p, 0, sizeof(double2) * s->dim, s->stream));
    const u64 ref_phys = (s->mode == 1) ? mqc_map_mask(s->ref_index, s->sched[which].phys_init) : s->ref_index;
    k_set_amplitude<<<1, 32, 0, s->stream>>>(s->psi.p, ref_phys, 1.0, 0.0);
    s->counters[0]++;
    if (s->mode == 1) {
        const int np = (int)s->sched[which].passes.size();
        for (int p = 0; p < np; ++p) launch_tile(s, which, p, false, 0);
        s->counters[1] += np;
    } else {
        for (int k = 0; k < nfac; ++k) {
            const MqcFactor& F = s->fac[k];
            const u64 npairs = s->dim >> F.npos;
            const unsigned grid = (unsigned)std::min<u64>((npairs + 255) / 256, 148ull * 16);
            k_excite_direct<<<grid, 256, 0, s->stream>>>(s->psi.p, F, s->cs.p, 0, npairs);
            s->counters[0]++;
        }
        s->counters[1] += nfac;
    }
    s->state_layout = which;
}

static void happly(mqc_solver* s, int which, const HamTables& T, const SecHamTables* ST, double2* lam_out,
                   double* host_val2) {
    LayoutTables& L = s->lay[which];
    const bool sec = s->conserving && s->use_sector && ST;
    CK(cudaMemsetAsync(s->scalars.p, 0, 2 * sizeof(double), s->stream));
    if (sec) {
        const unsigned nA = (unsigned)L.astr.n, nB = (unsigned)L.bstr.n;
        const u64 cnt = (u64)nA * nB;
        if (s->csec.n < cnt) { s->csec.alloc(cnt); s->lsec.alloc(cnt); }
        const unsigned gl = (unsigned)((cnt + 255) / 256);
        k_sector_gather<<<gl, 256, 0, s->stream>>>(s->psi.p, s->csec.p, L.astr.p, L.bstr.p, nA, nB);
        unsigned bd = 256;
        while (bd > 32 && (u64)(bd / 2) * MQC_HS_R >= nB) bd >>= 1;
        dim3 grid(nA, (nB + bd * MQC_HS_R - 1) / (bd * MQC_HS_R));
        k_happly_sector<<<grid, bd, 0, s->stream>>>(s->csec.p, lam_out ? s->lsec.p : nullptr, sec_ham_dev(L, *ST), s->scalars.p);
        s->counters[0] += 2;
        if (lam_out) {
            CK(cudaMemsetAsync(lam_out, 0, sizeof(double2) * s->dim, s->stream));
            k_sector_scatter<<<gl, 256, 0, s->stream>>>(lam_out, s->lsec.p, L.astr.p, L.bstr.p, nA, nB);
            s->counters[0]++;
        }
    } else {
        MqcSectorDev S = sector_dev(L, false);
        const u64 count = s->dim;
        const unsigned grid = (unsigned)((count + 255) / 256);
        k_happly<<<grid, 256, 0, s->stream>>>(s->psi.p, lam_out, T.dev(), S, count, s->scalars.p);
        s->counters[0]++;
    }
    if (host_val2) {
        CK(cudaMemcpyAsync(host_val2, s->scalars.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
}

static void reverse_sweep(mqc_solver* s, int which) {
    CK(cudaMemsetAsync(s->grad.p, 0, sizeof(double) * std::max(1, s->n_theta), s->stream));
    const int nfac = (int)s->fac.size();
    if (s->mode == 1) {
        const int np = (int)s->sched[which].passes.size();
        for (int p = np - 1; p >= 0; --p) launch_tile(s, which, p, true, 1);
        s->counters[2] += np;
    } else {
        for (int k = nfac - 1; k >= 0; --k) {
            const MqcFactor& F = s->fac[k];
            const u64 npairs = s->dim >> F.npos;
            const unsigned grid = (unsigned)std::min<u64>((npairs + 255) / 256, 148ull * 16);
            k_adjoint_direct<<<grid, 256, 0, s->stream>>>(s->psi.p, s->lam.p, F, s->cs.p, s->grad.p, npairs);
            s->counters[0]++;
        }
        s->counters[2] += nfac;
    }
    s->state_layout = -1;
}

static void reset_counters(mqc_solver* s) {
    s->counters[0] = s->counters[1] = s->counters[2] = 0;
    s->ms[0] = s->ms[1] = s->ms[2] = s->ms[3] = 0;
}

static void finish_timing(mqc_solver* s, bool with_reverse) {
    float t = 0;
    CK(cudaEventElapsedTime(&t, s->ev[0], s->ev[1])); s->ms[0] = t;
    CK(cudaEventElapsedTime(&t, s->ev[1], s->ev[2])); s->ms[1] = t;
    if (with_reverse) { CK(cudaEventElapsedTime(&t, s->ev[2], s->ev[3])); s->ms[2] = t; }
    CK(cudaEventElapsedTime(&t, s->ev[0], with_reverse ? s->ev[3] : s->ev[2])); s->ms[3] = t;
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

const char* mqc_last_error(void) { return g_err.c_str(); }
int mqc_version(void) { return 100; }

int mqc_device_count(int* count) {
    MQC_TRY
    if (!count) throw MqcError("null argument");
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) c = 0;
    *count = c;
    MQC_CATCH
}

int mqc_create(int n_qubits, int device, mqc_solver_t** out) {
    MQC_TRY
    if (!out) throw MqcError("null argument");
    if (n_qubits < 2 || n_qubits > 40) throw MqcError("n_qubits must be in [2, 40]");
    mqc_solver* s = new mqc_solver();
    s->n = n_qubits;
    s->n_local = n_qubits;
    s->device = device;
    s->dim = 1ull << n_qubits;
    *out = s;
    MQC_CATCH
}

int mqc_destroy(mqc_solver_t* s) {
    MQC_TRY
    if (!s) return 0;
    if (s->gpu_ready) {
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
        for (auto& e : s->ev) if (e) cudaEventDestroy(e);
    }
    cudaStream_t st = s->stream;
    bool ready = s->gpu_ready;
    delete s;
    if (ready && st) cudaStreamDestroy(st);
    MQC_CATCH
}

int mqc_set_option(mqc_solver_t* s, const char* key, int64_t value) {
    MQC_TRY
    if (!s || !key) throw MqcError("null argument");
    const std::string k(key);
    if (k == "tile_bits") s->tile_bits = (int)value;
    else if (k == "low_bits") s->low_bits = (int)value;
    else if (k == "grad_tile_bits") s->grad_tile_bits = (int)value;
    else if (k == "mode") s->mode = (int)value;
    else if (k == "sector") s->use_sector = (int)value;
    else if (k == "tile_threads") s->tile_threads = (int)value;
    else if (k == "sparse_tiles") s->sparse_tiles = (int)value;
    else if (k == "sparse_threads") s->sparse_threads = (int)value;
    else throw MqcError("unknown option: " + k);
    if (s->tile_bits < 5 || s->tile_bits > MQC_MAX_TILE_BITS || s->grad_tile_bits < 5 || s->grad_tile_bits > MQC_MAX_TILE_BITS - 1)
        throw MqcError("tile_bits must be in [5,13], grad_tile_bits in [5,12]");
    if (s->low_bits < 0 || s->low_bits > 6) throw MqcError("low_bits must be in [0,6]");
    if (s->tile_threads < 32 || s->tile_threads > 512 || (s->tile_threads & (s->tile_threads - 1)))
        throw MqcError("tile_threads must be a power of two in [32,512]");
    if (s->sparse_threads < 32 || s->sparse_threads > 256 || (s->sparse_threads & (s->sparse_threads - 1)))
        throw MqcError("sparse_threads must be a power of two in [32,256]");
    MQC_CATCH
}

int mqc_set_hamiltonian(mqc_solver_t* s, int64_t n_terms, const uint64_t* xmask, const uint64_t* zmask,
                        const double* coeff_re, const double* coeff_im) {
    MQC_TRY
    if (!s || n_terms < 0 || (n_terms && (!xmask || !zmask || !coeff_re))) throw MqcError("bad arguments");
    const u64 lim = (s->n >= 64) ? ~0ull : ((1ull << s->n) - 1ull);
    s->ham.x.assign(xmask, xmask + n_terms);
    s->ham.z.assign(zmask, zmask + n_terms);
    s->ham.c.resize(n_terms);
    for (int64_t t = 0; t < n_terms; ++t) {
        if ((xmask[t] | zmask[t]) & ~lim) throw MqcError("Pauli mask outside the register");
        s->ham.c[t] = make_double2(coeff_re[t], coeff_im ? coeff_im[t] : 0.0);
    }
    s->have_ham = true;
    s->lay[0].ham_ready = s->lay[1].ham_ready = false;
    s->lay[0].sham_ready = s->lay[1].sham_ready = false;
    MQC_CATCH
}

int mqc_set_ansatz(mqc_solver_t* s, int32_t n_factors, const int32_t* kind, const int32_t* idx4,
                   const int32_t* theta_index, int32_t n_theta, uint64_t ref_index) {
    MQC_TRY
    if (!s || n_factors < 0 || n_theta < 0 || (n_factors && (!kind || !idx4 || !theta_index)))
        throw MqcError("bad arguments");
    if (ref_index >> s->n) throw MqcError("reference index outside the register");
    if (s->n_global && s->mode != 1) throw MqcError("a sharded state needs mode = 1 (tile sweep)");
    const int n = s->n;
    std::vector<MqcFactor> fac;
    fac.reserve(n_factors);
    bool conserving = (n % 2 == 0);
    for (int k = 0; k < n_factors; ++k) {
        if (theta_index[k] < 0 || theta_index[k] >= n_theta) throw MqcError("theta_index out of range");
        fac.push_back(make_factor(n, kind[k], idx4 + 4 * k, k, theta_index[k]));
        const int32_t* ix = idx4 + 4 * k;
        if (kind[k] == 1) { if ((ix[0] ^ ix[1]) & 1) conserving = false; }
        else {
            const int sa = (ix[0] & 1) + (ix[1] & 1), sc = (ix[2] & 1) + (ix[3] & 1);
            if (sa != sc) conserving = false;
        }
    }
    s->fac = fac;
    s->n_theta = n_theta;
    s->ref_index = ref_index;
    s->conserving = conserving;
    s->na = s->nb = 0;
    for (int q = 0; q < n; ++q)
        if ((ref_index >> (n - 1 - q)) & 1ull) { if (q & 1) s->nb++; else s->na++; }
    // schedules: 0 = energy-only, 1 = energy + gradient
    for (int w = 0; w < 2; ++w) {
        const int kb = (w == 0) ? s->tile_bits : s->grad_tile_bits;
        if (s->mode == 1) {
            s->sched[w] = mqc_build_schedule(n, kb, s->low_bits, fac, MQC_MAX_PASS_FACTORS, s->n_global);
        } else {
            s->sched[w] = MqcSchedule();
            s->sched[w].n_bits = n;
            s->sched[w].phys_final.resize(n);
            std::iota(s->sched[w].phys_final.begin(), s->sched[w].phys_final.end(), 0);
            s->sched[w].phys_init = s->sched[w].phys_final;
        }
        s->lay[w].reset();
        s->lay[w].phys = s->sched[w].phys_final;
    }
    s->have_ansatz = true;
    s->state_layout = -1;
    // device side is set up lazily so that schedules can be inspected without a GPU
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) == cudaSuccess && cnt > 0) {
        require_gpu(s);
        std::vector<int> ti(n_factors);
        for (int k = 0; k < n_factors; ++k) ti[k] = theta_index[k];
        s->theta_index.upload(ti, s->stream);
        s->theta.alloc(std::max(1, n_theta));
        s->grad.alloc(std::max(1, n_theta));
        s->cs.alloc(std::max(1, n_factors));
        s->scalars.alloc(4);
        for (int w = 0; w < 2; ++w) {
            s->tf[w].upload(s->sched[w].tf, s->stream);
            std::vector<int> ph = s->lay[w].phys;
            s->lay[w].phys_dev.upload(ph, s->stream);
        }
        if (!s->psi.p) s->psi.alloc(s->dim);
        CK(cudaStreamSynchronize(s->stream));
    } else {
        cudaGetLastError();
    }
    MQC_CATCH
}

int mqc_prepare(mqc_solver_t* s, const double* theta) {
    MQC_TRY
    require_gpu(s);
    if (!theta) throw MqcError("null theta");
    reset_counters(s);
    forward(s, 0, theta);
    CK(cudaStreamSynchronize(s->stream));
    MQC_CATCH
}

int mqc_energy(mqc_solver_t* s, const double* theta, double* energy) {
    MQC_TRY
    require_gpu(s);
    if (!theta || !energy) throw MqcError("null argument");
    reset_counters(s);
    ensure_layout_ready(s, 0);
    double val[2] = {0, 0};
    CK(cudaEventRecord(s->ev[0], s->stream));
    forward(s, 0, theta);
    CK(cudaEventRecord(s->ev[1], s->stream));
    happly(s, 0, s->lay[0].ham, &s->lay[0].sham, nullptr, val);
    CK(cudaEventRecord(s->ev[2], s->stream));
    CK(cudaStreamSynchronize(s->stream));
    finish_timing(s, false);
    *energy = val[0];
    MQC_CATCH
}

static void run_energy_grad(mqc_solver* s, const double* theta, double* energy, double* grad) {
    reset_counters(s);
    ensure_layout_ready(s, 1);
    if (!s->lam.p) s->lam.alloc(s->dim);
    double val[2] = {0, 0};
    CK(cudaEventRecord(s->ev[0], s->stream));
    forward(s, 1, theta);
    CK(cudaEventRecord(s->ev[1], s->stream));
    happly(s, 1, s->lay[1].ham, &s->lay[1].sham, s->lam.p, val);
    CK(cudaEventRecord(s->ev[2], s->stream));
    reverse_sweep(s, 1);
    CK(cudaEventRecord(s->ev[3], s->stream));
    if (grad && s->n_theta)
        CK(cudaMemcpyAsync(grad, s->grad.p, sizeof(double) * s->n_theta, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    finish_timing(s, true);
    *energy = val[0];
}

int mqc_energy_grad(mqc_solver_t* s, const double* theta, double* energy, double* grad) {
    MQC_TRY
    require_gpu(s);
    if (!theta || !energy || !grad) throw MqcError("null argument");
    run_energy_grad(s, theta, energy, grad);
    MQC_CATCH
}

int mqc_energy_grad_resident(mqc_solver_t* s, double* energy) {
    MQC_TRY
    require_gpu(s);
    if (!energy) throw MqcError("null argument");
    if ((int)s->theta_host.size() != s->n_theta || !s->have_ansatz)
        throw MqcError("no resident parameters: call mqc_energy_grad once first");
    run_energy_grad(s, nullptr, energy, nullptr);
    MQC_CATCH
}

int mqc_event_record(mqc_solver_t* s, int32_t which) {
    MQC_TRY
    require_gpu(s);
    if (which < 0 || which > 1) throw MqcError("event index must be 0 or 1");
    if (!s->user_ev[which]) CK(cudaEventCreate(&s->user_ev[which]));
    CK(cudaEventRecord(s->user_ev[which], s->stream));
    MQC_CATCH
}

int mqc_event_elapsed_ms(mqc_solver_t* s, double* ms) {
    MQC_TRY
    require_gpu(s);
    if (!ms || !s->user_ev[0] || !s->user_ev[1]) throw MqcError("record both events first");
    CK(cudaEventSynchronize(s->user_ev[1]));
    float t = 0;
    CK(cudaEventElapsedTime(&t, s->user_ev[0], s->user_ev[1]));
    *ms = t;
    MQC_CATCH
}

static void ensure_state(mqc_solver* s, const double* theta) {
    if (theta) { forward(s, 0, theta); return; }
    if (s->state_layout >= 0) return;
    if ((int)s->theta_host.size() != s->n_theta || !s->have_ansatz) throw MqcError("no prepared state: pass theta");
    std::vector<double> th = s->theta_host;
    forward(s, 0, th.data());
}

int mqc_state(mqc_solver_t* s, const double* theta, double* re_im, int64_t n_amplitudes) {
    MQC_TRY
    require_gpu(s);
    if (!re_im || (u64)n_amplitudes != s->dim) throw MqcError("state buffer must hold 2^N amplitudes");
    reset_counters(s);
    ensure_state(s, theta);
    const int w = s->state_layout;
    if (!s->scratch_state.p) s->scratch_state.alloc(s->dim);
    const unsigned grid = (unsigned)std::min<u64>((s->dim + 255) / 256, 148ull * 32);
    k_gather_logical<<<grid, 256, 0, s->stream>>>(s->psi.p, s->scratch_state.p, s->lay[w].phys_dev.p, s->n, s->dim);
    CK(cudaMemcpyAsync(re_im, s->scratch_state.p, sizeof(double2) * s->dim, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    MQC_CATCH
}

int mqc_expectation(mqc_solver_t* s, int64_t n_terms, const uint64_t* xmask, const uint64_t* zmask,
                    const double* coeff_re, const double* coeff_im, double* value_re, double* value_im) {
    MQC_TRY
    require_gpu(s);
    if (n_terms < 0 || (n_terms && (!xmask || !zmask || !coeff_re)) || !value_re) throw MqcError("bad arguments");
    ensure_state(s, nullptr);
    const int w = s->state_layout;
    HamHost H;
    H.x.assign(xmask, xmask + n_terms);
    H.z.assign(zmask, zmask + n_terms);
    H.c.resize(n_terms);
    for (int64_t t = 0; t < n_terms; ++t) H.c[t] = make_double2(coeff_re[t], coeff_im ? coeff_im[t] : 0.0);
    HamTables T;
    SecHamTables ST;
    const bool sec = s->conserving && s->use_sector;
    if (sec) { build_sector_tables(s, s->lay[w], s->na, s->nb); build_sec_ham(s, H, s->lay[w].phys, ST); }
    else build_ham_tables(s, H, s->lay[w].phys, T);
    double val[2] = {0, 0};
    happly(s, w, T, sec ? &ST : nullptr, nullptr, val);
    CK(cudaStreamSynchronize(s->stream));
    *value_re = val[0];
    if (value_im) *value_im = val[1];
    MQC_CATCH
}

int mqc_rdm12(mqc_solver_t* s, int32_t n_orb, int32_t n_alpha, int32_t n_beta, double* rdm1, double* rdm2) {
    MQC_TRY
    require_gpu(s);
    if (!rdm1) throw MqcError("null rdm1");
    if (2 * n_orb != s->n) throw MqcError("n_orb must be n_qubits / 2");
    ensure_state(s, nullptr);
    const int w = s->state_layout;
    LayoutTables& L = s->lay[w];
    build_sector_tables(s, L, n_alpha, n_beta);
    MqcSectorDev S = sector_dev(L, true);
    const u64 count = (u64)L.astr.n * (u64)L.bstr.n;
    const size_t n2 = (size_t)n_orb * n_orb, n4 = n2 * n2;
    const int copies = 8;
    MqcRdmDev R{n_orb, L.qbit.p, L.below.p, copies};
    const size_t need = (size_t)copies * (rdm2 ? n4 : n2);
    if (s->rdm_acc.n < need) s->rdm_acc.alloc(need);
    CK(cudaMemsetAsync(s->rdm_acc.p, 0, sizeof(double) * (size_t)copies * n2, s->stream));
    k_rdm1<<<(unsigned)((count + 127) / 128), 128, 0, s->stream>>>(s->psi.p, S, R, count, s->rdm_acc.p);
    k_reduce_copies<<<(unsigned)((n2 + 255) / 256), 256, 0, s->stream>>>(s->rdm_acc.p, n2, copies);
    CK(cudaMemcpyAsync(rdm1, s->rdm_acc.p, sizeof(double) * n2, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (rdm2) {
        CK(cudaMemsetAsync(s->rdm_acc.p, 0, sizeof(double) * (size_t)copies * n4, s->stream));
        const u64 work = count * (u64)(2 * n_orb);
        k_rdm2<<<(unsigned)((work + 127) / 128), 128, 0, s->stream>>>(s->psi.p, S, R, count, s->rdm_acc.p);
        k_reduce_copies<<<(unsigned)((n4 + 255) / 256), 256, 0, s->stream>>>(s->rdm_acc.p, n4, copies);
        CK(cudaMemcpyAsync(rdm2, s->rdm_acc.p, sizeof(double) * n4, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    // the sector tables of this layout may now be for a different (na, nb) than the ansatz's
    if (s->conserving && s->use_sector && (n_alpha != s->na || n_beta != s->nb)) build_sector_tables(s, L, s->na, s->nb);
    MQC_CATCH
}

int mqc_last_timing(mqc_solver_t* s, double* ms4, int64_t* counters3) {
    MQC_TRY
    if (!s) throw MqcError("null solver handle");
    if (ms4) for (int i = 0; i < 4; ++i) ms4[i] = s->ms[i];
    if (counters3) for (int i = 0; i < 3; ++i) counters3[i] = s->counters[i];
    MQC_CATCH
}

int mqc_schedule_info(mqc_solver_t* s, int32_t which, int32_t* n_passes, int32_t* tile_bits, int32_t* n_tile_factors) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz) throw MqcError("no schedule");
    if (n_passes) *n_passes = (int32_t)s->sched[which].passes.size();
    if (tile_bits) *tile_bits = s->sched[which].k;
    if (n_tile_factors) *n_tile_factors = (int32_t)s->sched[which].tf.size();
    MQC_CATCH
}

int mqc_schedule_pass(mqc_solver_t* s, int32_t which, int32_t pass, int32_t* first_factor, int32_t* n_factors,
                      uint64_t* act_mask, int32_t* pos16, int32_t* perm16, int32_t* has_perm) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz) throw MqcError("no schedule");
    const MqcSchedule& S = s->sched[which];
    if (pass < 0 || pass >= (int)S.passes.size()) throw MqcError("pass out of range");
    const MqcPass& P = S.passes[pass];
    if (first_factor) *first_factor = P.f0;
    if (n_factors) *n_factors = P.f1 - P.f0;
    if (act_mask) *act_mask = P.act_mask;
    if (pos16) for (int j = 0; j < 16; ++j) pos16[j] = j < P.k ? P.pos[j] : -1;
    if (perm16) for (int j = 0; j < 16; ++j) perm16[j] = j < P.k ? P.perm[j] : -1;
    if (has_perm) *has_perm = P.has_perm;
    MQC_CATCH
}

int mqc_schedule_factor(mqc_solver_t* s, int32_t which, int32_t tile_factor, uint32_t* m, uint32_t* v, uint32_t* z,
                        uint64_t* z_out, int32_t* sign0, int32_t* slot) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz) throw MqcError("no schedule");
    const MqcSchedule& S = s->sched[which];
    if (tile_factor < 0 || tile_factor >= (int)S.tf.size()) throw MqcError("factor out of range");
    const MqcTileFactor& F = S.tf[tile_factor];
    if (m) *m = F.m;
    if (v) *v = F.v;
    if (z) *z = F.z;
    if (z_out) *z_out = F.z_out;
    if (sign0) *sign0 = F.sign0;
    if (slot) *slot = F.slot;
    MQC_CATCH
}

int mqc_schedule_final_layout(mqc_solver_t* s, int32_t which, int32_t* phys_of_logical) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz || !phys_of_logical) throw MqcError("no schedule");
    for (int l = 0; l < s->n; ++l) phys_of_logical[l] = s->sched[which].phys_final[l];
    MQC_CATCH
}

int mqc_schedule_initial_layout(mqc_solver_t* s, int32_t which, int32_t* phys_of_logical) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz || !phys_of_logical) throw MqcError("no schedule");
    for (int l = 0; l < s->n; ++l) phys_of_logical[l] = s->sched[which].phys_init[l];
    MQC_CATCH
}

}  // extern "C"
