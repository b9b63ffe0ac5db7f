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
#include "mqc_sigma.h"
#include "mqc_sigma_ab.h"
#include "mqc_cplan.h"
#include "mqc_csweep.cuh"
#include "mqc_rdm.cuh"

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

struct SigmaTables {                // link-table form of a sector Hamiltonian (mqc_sigma.h), layout independent
    DevBuf<unsigned> lnk_a1, lnk_a2, ell_b1, ell_b2, m2_mask, sa_b2, sb_a2, sa_b4, sb_a4, aorb, borb;
    DevBuf<int> cls_off, cls_of_m2;
    DevBuf<double> w22, w4a, w4b, ta1, tb1, tab, tbb, hdiag, diag_c;
    DevBuf<u64> aq, bq, diag_z;
    MqcSigmaDev dev{};
    // alpha-beta doubles in the sign-adjusted basis (mqc_sigma_ab.h); their Pauli terms are NOT in the tables above
    DevBuf<double> ab_w;
    DevBuf<unsigned> ab_lnk, ab_ell, ab_g;
    MqcSigmaABDev ab{};
    bool ab_ready = false;
    size_t n_ab_terms = 0;
    bool ready = false;
    unsigned row0 = 0, row1 = 0;
    size_t n_fallback = 0, n_fast = 0;
};

struct CPlanDev {                   // device copy of a compact-sector plan (mqc_cplan.h)
    DevBuf<MqcCPassDev> pass;
    DevBuf<MqcCClass> cls;
    DevBuf<MqcCFactor> fac;
    DevBuf<uint32_t> maps, tiles;
    DevBuf<MqcCLive> live_tab;
    DevBuf<MqcCProgHdr> prog_hdr;
    DevBuf<uint32_t> prog;
    void reset() { pass.release(); cls.release(); fac.release(); maps.release(); tiles.release(); live_tab.release(); prog_hdr.release(); prog.release(); }
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

// mail buffer layout (doubles): [0..3] scalars, [8..15] barrier flags (as u64), [16..) gradient.
// One allocation so that a peer maps it with one CUDA IPC handle.
#define MQC_MAIL_FLAGS 8
#define MQC_MAIL_GRAD 16
#define MQC_BLOB_BYTES 512

struct mqc_solver {
    int n = 0, device = 0;
    u64 dim = 0;                    // amplitudes held by THIS rank (2^n_local)
    int n_global = 0, rank = 0, n_ranks = 1;   // sharded state: top n_global physical bits = rank id
    int n_local = 0;
    // 0: one rank.  1: all ranks are handles of this process, driven in lockstep by the leader
    // (members[0]), ordered by CUDA events.  2: one process per GPU, peers mapped with CUDA IPC,
    // ordered by k_rank_barrier.
    int comm_mode = 0;
    std::vector<mqc_solver*> members;          // ranks this handle drives (itself first)
    bool connected = false;
    double2* peer_psi[MQC_MAX_RANKS] = {};
    double2* peer_lam[MQC_MAX_RANKS] = {};
    double* peer_mail[MQC_MAX_RANKS] = {};
    std::vector<std::pair<std::string, void*>> ipc_open;   // handle bytes -> mapped base
    u64 epoch = 0;
    cudaEvent_t sync_ev = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t user_ev[2] = {nullptr, nullptr};
    bool gpu_ready = false;
    int n_sm = 148, sp2_carveout_set = -1;

    // options
    int tile_bits = 14, low_bits = -1, grad_tile_bits = 14, mode = 1, use_sector = 1, tile_threads = 256, sparse_tiles = 1, sparse_threads = 128;

    DevBuf<double2> psi, lam, cs, scratch_state, csec, lsec, cown, cblk;
    double2* peer_cown[MQC_MAX_RANKS] = {};
    int sigma_blocked = -1;          // -1 auto, 0 replicate C on every rank, 1 distribute its rows
    bool blocked = false, connected_cown = false;
    DevBuf<double> theta, mail, red, rdm_acc, rdm_gpart, rdm_g1part, rdm_out;
    int rdm_gram = 1;                // K4 on the compact matrix as a Gram update (mqc_rdm.cuh); 0: literal loops with atomics
    DevBuf<int> theta_index, err;
    DevBuf<double> theta_scale, gtheta;      // per-factor angle scale; gradient with respect to the parameters
    DevBuf<MqcTileFactor> tf[2];
    DevBuf<uint16_t> dlut[2];       // pair-address tables of the dense tile kernel
    MqcSparsePlan plan[2];          // compressed-sector tile plan per schedule (host copy of the pass records)
    DevBuf<u64> pl_dep[2], pl_depp[2];
    DevBuf<uint32_t> pl_lists[2];
    DevBuf<uint16_t> pl_pat[2];
    DevBuf<uint4> pl_hdr[2];
    MqcTileLists tlist[2];          // host copy of the offsets; tiles live in pl_tiles
    DevBuf<u64> pl_tiles[2];
    int sp2 = 1, sp2_warps = 4, sp2_list = 1, sp2_smem_kb = 220, sp2_team = 0;
    // compact sector storage (mqc_cplan.h): the state is the matrix C[alpha string][beta string];
    // cpsi / clam ping-pong between the layouts of consecutive passes
    int compact = 1, csweep_threads = 128, csweep_ctas = 0, csweep_snake = 1;
    // Reproducible sums (default): the adjoint sweep leaves per-(CTA, factor) gradient partials that are added in a fixed
    // order, K2 adds its per-CTA energy partials in a fixed order - two evaluations with the same input give the same
    // bits, so an optimiser or a chemical-potential search on top takes the same path every time.  0: atomics.
    int deterministic = 1;
    DevBuf<double> cgpart, e_part;
    DevBuf<uint4> cgwhere;
    DevBuf<int> th_ptr, th_fac;      // factors of every parameter (CSR, factor order)
    DevBuf<unsigned> e_ticket;
    std::vector<size_t> gp_off;
    bool gpart_ready = false;
    int csweep_pdl = 0;              // programmatic dependent launch of the sweep passes: measured slower (0.71 / 1.30 -> 0.86 / 1.49 ms:
                                     // the early CTAs land on whichever SMs drain first and the next pass is unbalanced), kept as an option
    int csweep_adj_threads = 0;      // threads per tile of the adjoint sweep; 0: 64 with real amplitudes (measured), else csweep_threads
    bool use_compact = false;
    // Real amplitudes: every ansatz factor is a real rotation, so with a real Hamiltonian (even number of Y per
    // term, real coefficients, all of it on the link-table kernel) psi and lambda never get an imaginary part
    // and the compact buffers hold 8-byte amplitudes (half the shared-memory traffic and arithmetic of every
    // kernel).  The handle falls back to complex storage by itself when a Hamiltonian or a loaded state is not real.
    int real_amplitudes = 1;         // option: 0 = always complex
    bool creal = false;              // the compact buffers hold doubles
    DevBuf<double2> cscratch;        // complex copy of the real state for the consumers that want one
    MqcCPlan cplan[2];
    CPlanDev cdev[2];
    DevBuf<double2> cpsi[2], clam[2];
    int ccur = 0, clcur = 0;         // buffers that hold psi / lambda
    bool cvalid = false;             // cpsi[ccur] holds the prepared state (canonical order) of schedule cwhich
    int cwhich = 0;
    double* scalars_p() const { return mail.p; }
    double* grad_p() const { return mail.p + MQC_MAIL_GRAD; }
    unsigned long long* flags_p() const { return reinterpret_cast<unsigned long long*>(mail.p + MQC_MAIL_FLAGS); }

    HamHost ham, ham_fb;            // all terms / the terms the link-table kernel leaves to k_happly_sector
    SigmaTables sig;
    int sigma_stage_one = 0;
    int sigma_stage_bytes = 98304;                        // shared memory the staged source rows of K2 may take
    int sigma_split = -1;                                 // K2 row split: -1 auto, 0 off, k: part 0 takes the first k alpha single links
    int use_sigma = 1, sigma_stage = 1, sigma_ab = 0;     // sigma_ab: opposite-spin doubles through k_sigma_ab (measured: no faster, both are shared-memory-gather bound)
    // The launch sequence of an evaluation depends on theta only through device buffers, so a
    // single-rank handle replays it as a CUDA graph from the third call on (first call: eager,
    // allocates; second: captured).  Any change of tables or options drops the graphs.
    int use_graph = 1;
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    int eager_done[2] = {0, 0};
    int64_t gcounters[2][3] = {{0, 0, 0}, {0, 0, 0}};
    bool capturing = false;
    bool have_ham = false, have_ansatz = false;

    std::vector<MqcFactor> fac;     // logical coordinates
    std::vector<int> fac_theta;     // parameter index of factor k (host copy)
    std::vector<double> fac_scale;  // angle of factor k = fac_scale[k] * theta[fac_theta[k]]
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

static void invalidate_graphs(mqc_solver* s);
extern "C" {                      // defined inside the extern "C" block below
static const double2* compact_complex_view(mqc_solver* s);
static void drop_real(mqc_solver* s);
static bool ham_is_real(const HamHost& H);
}

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
    } else if (kind == 3) {
        // number-dressed single n_c (a_a^+ a_i - h.c.): the rotation of kind 1 on the determinants with c occupied
        const int i = ix[0], a = ix[1], c = ix[2];
        chk(i); chk(a); chk(c);
        if (i == a || c == i || c == a) throw MqcError("dressed single needs three distinct orbitals");
        mq = (1ull << i) | (1ull << a);
        vq = 1ull << i;
        zq = between_mask_qubits(i, a);
        u64 D = 1ull << i;
        int p1 = ladder(D, i, 0), p2 = ladder(D, a, 1);
        sign0 = (p1 + p2) & 1;
        f.ctrl = qubit_to_logical(1ull << c, n);
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
        throw MqcError("unknown factor kind (1 = single, 2 = double, 3 = number-dressed single)");
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
    CK(cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, s->device));
    for (auto& e2 : s->ev) CK(cudaEventCreate(&e2));
    CK(cudaEventCreateWithFlags(&s->sync_ev, cudaEventDisableTiming));
    CK(cudaFuncSetAttribute(k_tile_sweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CK(cudaFuncSetAttribute(k_tile_sweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CK(cudaFuncSetAttribute(k_tile_sparse<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_tile_sparse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
#define MQC_SP2_ATTR(TW_) \
    CK(cudaFuncSetAttribute(k_tile_sp2<false, false, TW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000)); \
    CK(cudaFuncSetAttribute(k_tile_sp2<true, false, TW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000)); \
    CK(cudaFuncSetAttribute(k_tile_sp2<false, true, TW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000)); \
    CK(cudaFuncSetAttribute(k_tile_sp2<true, true, TW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    MQC_SP2_ATTR(1) MQC_SP2_ATTR(2) MQC_SP2_ATTR(4)
#define MQC_CS_ATTR(NT_, CH_) \
    CK(cudaFuncSetAttribute(k_csweep<false, NT_, CH_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)); \
    CK(cudaFuncSetAttribute(k_csweep<true, NT_, CH_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)); \
    CK(cudaFuncSetAttribute(k_csweep<false, NT_, CH_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)); \
    CK(cudaFuncSetAttribute(k_csweep<true, NT_, CH_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    MQC_CS_ATTR(32, 256) MQC_CS_ATTR(64, 256) MQC_CS_ATTR(128, 256) MQC_CS_ATTR(256, 256) MQC_CS_ATTR(512, 256)
    MQC_CS_ATTR(32, 1024) MQC_CS_ATTR(64, 1024) MQC_CS_ATTR(128, 1024) MQC_CS_ATTR(256, 1024) MQC_CS_ATTR(512, 1024)
    CK(cudaFuncSetAttribute(k_sigma_ab<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_sigma_ab<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_rdm_gram<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_rdm_gram<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_rdm_gram<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_rdm_gram<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_sigma<true, MQC_SG_LB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_sigma<false, MQC_SG_LB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_sigma<true, MQC_SG_LB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_sigma<true, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_sigma<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_sigma<true, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_sigma<true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
    CK(cudaFuncSetAttribute(k_sigma<false, MQC_SG_LB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000));
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
    invalidate_graphs(s);               // device tables move
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

static double binom_d(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

static void sector_dims_pre(const mqc_solver* s, unsigned& nA, unsigned& nB) {
    nA = (unsigned)(binom_d((s->n + 1) / 2, s->na) + 0.5);
    nB = (unsigned)(binom_d(s->n / 2, s->nb) + 0.5);
}

static u64 logical_to_qubit(u64 m, int n) {
    u64 out = 0;
    for (int q = 0; q < n; ++q) if ((m >> (n - 1 - q)) & 1ull) out |= 1ull << q;
    return out;
}

static void row_slice(const mqc_solver* s, unsigned n_rows, unsigned& r0, unsigned& r1) {
    r0 = (unsigned)(((u64)n_rows * s->rank) / s->n_ranks);
    r1 = (unsigned)(((u64)n_rows * (s->rank + 1)) / s->n_ranks);
}

// link tables of operator H for the (na, nb) sector; the terms they do not cover go to `fb`
static void build_sigma(mqc_solver* s, const HamHost& H, int na, int nb, SigmaTables& T, HamHost& fb, bool split_ab = false) {
    const int n = s->n, n_orb = n / 2;
    std::vector<u64> xq(H.x.size()), zq(H.x.size());
    for (size_t t = 0; t < H.x.size(); ++t) { xq[t] = logical_to_qubit(H.x[t], n); zq[t] = logical_to_qubit(H.z[t], n); }
    // opposite-spin double excitations go to k_sigma_ab; the link tables get the rest
    std::vector<size_t> keep(H.x.size());
    std::iota(keep.begin(), keep.end(), 0);
    std::vector<double2> cq = H.c;
    T.ab_ready = false; T.n_ab_terms = 0;
    if (split_ab) {
        std::vector<unsigned> astr, bstr;
        std::vector<int> ra, rb;
        mqc_sg_strings(n_orb, na, astr, ra);
        mqc_sg_strings(n_orb, nb, bstr, rb);
        MqcSigmaAB AB = mqc_sigma_ab_build(n_orb, na, nb, xq, zq, H.c, astr, bstr);
        if (AB.ok) {
            std::vector<char> used(H.x.size(), 0);
            for (size_t t : AB.used_terms) used[t] = 1;
            keep.clear();
            for (size_t t = 0; t < H.x.size(); ++t) if (!used[t]) keep.push_back(t);
            std::vector<u64> x2, z2; std::vector<double2> c2;
            for (size_t t : keep) { x2.push_back(xq[t]); z2.push_back(zq[t]); c2.push_back(H.c[t]); }
            xq.swap(x2); zq.swap(z2); cq.swap(c2);
            T.ab_w.upload(AB.W, s->stream); T.ab_lnk.upload(AB.lnk_a, s->stream); T.ab_ell.upload(AB.ell_b, s->stream);
            T.ab_g.upload(AB.gB, s->stream);
            T.ab = MqcSigmaABDev{T.ab_w.p, T.ab_lnk.p, T.ab_ell.p, T.ab_g.p, nullptr, AB.n_a, AB.n_b, AB.e_a, AB.e_b, AB.n_pair};
            T.ab_ready = true; T.n_ab_terms = AB.used_terms.size();
        }
    }
    MqcSigmaHost S = mqc_sigma_build(n_orb, na, nb, xq, zq, cq);
    fb.x.clear(); fb.z.clear(); fb.c.clear();
    for (size_t t : S.fallback_terms) { fb.x.push_back(H.x[keep[t]]); fb.z.push_back(H.z[keep[t]]); fb.c.push_back(H.c[keep[t]]); }
    T.n_fallback = S.fallback_terms.size(); T.n_fast = S.n_fast_terms;
    cudaStream_t st = s->stream;
    T.lnk_a1.upload(S.lnk_a1, st); T.lnk_a2.upload(S.lnk_a2, st); T.ell_b1.upload(S.ell_b1, st); T.ell_b2.upload(S.ell_b2, st);
    T.m2_mask.upload(S.m2_mask.empty() ? std::vector<unsigned>(1, 0) : S.m2_mask, st);
    T.sa_b2.upload(S.sa_b2, st); T.sb_a2.upload(S.sb_a2, st); T.sa_b4.upload(S.sa_b4, st); T.sb_a4.upload(S.sb_a4, st);
    T.aorb.upload(S.aorb, st); T.borb.upload(S.borb, st);
    T.cls_off.upload(S.cls_off, st); T.cls_of_m2.upload(S.cls_of_m2, st);
    T.w22.upload(S.w22, st); T.w4a.upload(S.w4a, st); T.w4b.upload(S.w4b, st);
    T.ta1.upload(S.ta1, st); T.tb1.upload(S.tb1, st); T.tab.upload(S.tab, st); T.tbb.upload(S.tbb, st);
    T.aq.upload(S.aq, st); T.bq.upload(S.bq, st);
    if (S.diag_z.empty()) { S.diag_z.push_back(0); S.diag_c.push_back(0.0); }
    T.diag_z.upload(S.diag_z, st); T.diag_c.upload(S.diag_c, st);
    row_slice(s, S.n_a, T.row0, T.row1);
    const u64 cnt = (u64)(T.row1 - T.row0) * S.n_b;
    T.hdiag.alloc(std::max<u64>(1, cnt));
    if (cnt) k_sigma_hdiag<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(T.aq.p, T.bq.p, T.diag_z.p, T.diag_c.p, (int)S.diag_z.size(),
                                                                          T.row0, T.row1, S.n_b, T.hdiag.p);
    MqcSigmaDev D{};
    D.n_a = S.n_a; D.n_b = S.n_b; D.n_m2 = S.n_m2; D.n_m4 = S.n_m4;
    D.lnk_a1 = T.lnk_a1.p; D.e1a = S.e1a; D.lnk_a2 = T.lnk_a2.p; D.e2a = S.e2a;
    D.ell_b1 = T.ell_b1.p; D.cls_off = T.cls_off.p; D.cls_of_m2 = T.cls_of_m2.p; D.ell_b2 = T.ell_b2.p; D.e2b = S.e2b;
    D.w22 = T.w22.p; D.w4a = T.w4a.p; D.w4b = T.w4b.p; D.ta1 = T.ta1.p; D.tb1 = T.tb1.p; D.tab = T.tab.p; D.tbb = T.tbb.p;
    D.m2_mask = T.m2_mask.p; D.sa_b2 = T.sa_b2.p; D.sb_a2 = T.sb_a2.p; D.sa_b4 = T.sa_b4.p; D.sb_a4 = T.sb_a4.p;
    D.aorb = T.aorb.p; D.borb = T.borb.p; D.hdiag = T.hdiag.p; D.row0 = T.row0;
    T.dev = D;
    T.ab.aorb = T.aorb.p;
    CK(cudaStreamSynchronize(st));
    T.ready = true;
}

static void ensure_layout_ready(mqc_solver* s, int which) {
    LayoutTables& L = s->lay[which];
    if (!s->have_ham) throw MqcError("mqc_set_hamiltonian has not been called");
    if (!L.ham_ready && !(s->conserving && s->use_sector)) {
        invalidate_graphs(s);
        build_ham_tables(s, s->ham, L.phys, L.ham);
        L.ham_ready = true;
    }
    if (s->conserving && s->use_sector) {
        build_sector_tables(s, L, s->na, s->nb);
        const bool sigma = s->use_sigma && s->n / 2 <= 18;
        if (sigma && !s->sig.ready) { invalidate_graphs(s); build_sigma(s, s->ham, s->na, s->nb, s->sig, s->ham_fb, s->use_compact && s->sigma_ab); }
        if (s->creal && !(sigma && s->sig.ready && !s->sig.n_fallback && !s->sig.ab_ready)) drop_real(s);   // part of H is outside the link tables
        if (!L.sham_ready) { invalidate_graphs(s); build_sec_ham(s, sigma ? s->ham_fb : s->ham, L.phys, L.sham); L.sham_ready = true; }
    }
}

// ---------------------------------------------------------------------------------------
// launch sequences.  Every evaluation is written for a GROUP of ranks: `L` is the handle the
// caller holds; it drives L->members (all ranks in same-process mode, only itself otherwise).
// Per-rank work is launched rank by rank on each rank's own stream; group_sync() orders the
// ranks wherever one reads or writes another's shard.
// ---------------------------------------------------------------------------------------
static void invalidate_graphs(mqc_solver* s) {
    s->gpart_ready = false;          // grids of the adjoint passes may change with the options / plans
    for (int w = 0; w < 2; ++w) {
        if (s->gexec[w]) { cudaGraphExecDestroy(s->gexec[w]); s->gexec[w] = nullptr; }
        s->eager_done[w] = 0;
    }
}
static void record_ev(mqc_solver* s, int i) {
    CK(cudaSetDevice(s->device));
    if (s->capturing) CK(cudaEventRecordWithFlags(s->ev[i], s->stream, cudaEventRecordExternal));
    else CK(cudaEventRecord(s->ev[i], s->stream));
}

static double binom(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

static void require_connected(mqc_solver* L) {
    if (L->n_ranks > 1 && !L->connected)
        throw MqcError("sharded handle is not connected: call mqc_shard_connect (or create it with mqc_shard_group_create)");
}

static void group_sync(mqc_solver* L) {
    if (L->n_ranks == 1) return;
    if (L->comm_mode == 1) {
        for (mqc_solver* m : L->members) { CK(cudaSetDevice(m->device)); CK(cudaEventRecord(m->sync_ev, m->stream)); }
        for (mqc_solver* m : L->members) {
            CK(cudaSetDevice(m->device));
            for (mqc_solver* q : L->members) if (q != m) CK(cudaStreamWaitEvent(m->stream, q->sync_ev, 0));
        }
    } else {
        MqcFlags F{};
        F.n_ranks = L->n_ranks; F.rank = L->rank;
        for (int r = 0; r < L->n_ranks; ++r)
            F.flags[r] = reinterpret_cast<unsigned long long*>(L->peer_mail[r] + MQC_MAIL_FLAGS);
        ++L->epoch;
        CK(cudaSetDevice(L->device));
        k_rank_barrier<<<1, 32, 0, L->stream>>>(F, (unsigned long long)L->epoch, L->err.p);
        L->counters[0]++;
    }
}

static MqcPeers peers_of(const mqc_solver* s) {
    MqcPeers R{};
    for (int r = 0; r < s->n_ranks; ++r) { R.psi[r] = s->peer_psi[r]; R.lam[r] = s->peer_lam[r]; }
    if (s->n_ranks == 1) { R.psi[0] = s->psi.p; R.lam[0] = s->lam.p; }
    R.n_local = s->n_local;
    R.tau_hi = 0; R.fixed_bits = (u64)s->rank << s->n_local;
    return R;
}

static bool pass_is_global(const mqc_solver* s, int which, int pass) {
    return s->n_global > 0 && (s->sched[which].passes[pass].act_mask >> s->n_local) != 0ull;
}

static void launch_tile(mqc_solver* s, int which, int pass, bool adj, int reverse) {
    const MqcSchedule& S = s->sched[which];
    const MqcPass& P = S.passes[pass];
    const int nf = P.f1 - P.f0;
    const MqcTileMap TM = mqc_tile_map(P, s->n, s->n_global, s->rank);
    const unsigned grid = (unsigned)TM.n_tiles;
    MqcPeers PR = peers_of(s);
    PR.tau_hi = TM.tau_hi; PR.fixed_bits = TM.fixed_bits;
    CK(cudaSetDevice(s->device));
    const bool sparse = s->conserving && s->use_sector && s->sparse_tiles;
    if (sparse && s->sp2 && s->plan[which].ok) {
        const MqcSpPass& SPP = s->plan[which].pass[pass];
        // warps per tile: large tiles are shared by a team of warps
        int tw = s->sp2_team;
        if (tw <= 0) tw = SPP.cap >= 300 ? 2 : 1;
        int warps = std::max(s->sp2_warps, tw);
        warps -= warps % tw;
        size_t smem = mqc_sp2_smem_bytes(SPP, nf, adj, warps, tw);
        while (warps > tw && smem > 200000) { warps -= tw; smem = mqc_sp2_smem_bytes(SPP, nf, adj, warps, tw); }
        if (smem <= 200000) {
            int per_sm = (int)std::min<size_t>(16, ((size_t)s->sp2_smem_kb * 1024) / (smem + 1024));
            per_sm = std::max(1, std::min(per_sm, 64 / warps));
            const bool list = s->tlist[which].ok && s->sp2_list;
            MqcSpTables TB{s->pl_dep[which].p, s->pl_depp[which].p, s->pl_lists[which].p, s->pl_pat[which].p, s->pl_hdr[which].p, nullptr, 0};
            u64 n_work = TM.n_tiles;
            if (list) {
                TB.tiles = s->pl_tiles[which].p + s->tlist[which].off[pass];
                TB.n_list = s->tlist[which].off[pass + 1] - s->tlist[which].off[pass];
                n_work = TB.n_list;
            }
            if (n_work == 0) return;
            const int teams = warps / tw;
            const u64 want = (n_work + teams - 1) / teams;
            const unsigned g2 = (unsigned)std::min<u64>(want, (u64)s->n_sm * per_sm);
            const unsigned nt = warps * 32;
#define MQC_SP2_LAUNCH(ADJ_, LIST_, TW_) k_tile_sp2<ADJ_, LIST_, TW_><<<g2, nt, smem, s->stream>>>(PR, P, SPP, TB, s->tf[which].p, s->cs.p, adj ? s->grad_p() : nullptr, reverse, TM.n_tiles)
#define MQC_SP2_PICK(TW_) do { if (adj) { if (list) MQC_SP2_LAUNCH(true, true, TW_); else MQC_SP2_LAUNCH(true, false, TW_); } \
                               else { if (list) MQC_SP2_LAUNCH(false, true, TW_); else MQC_SP2_LAUNCH(false, false, TW_); } } while (0)
            if (tw == 4) MQC_SP2_PICK(4); else if (tw == 2) MQC_SP2_PICK(2); else MQC_SP2_PICK(1);
            s->counters[0]++;
            return;
        }
    }
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
            if (adj) k_tile_sparse<true><<<grid, nthr, smem, s->stream>>>(PR, P, s->tf[which].p, s->cs.p, s->grad_p(), reverse, SP);
            else k_tile_sparse<false><<<grid, nthr, smem, s->stream>>>(PR, P, s->tf[which].p, s->cs.p, nullptr, reverse, SP);
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
        k_tile_sweep<true><<<grid, nthr, smem, s->stream>>>(PR, P, s->tf[which].p, s->dlut[which].p, s->cs.p, s->grad_p(), reverse);
    else
        k_tile_sweep<false><<<grid, nthr, smem, s->stream>>>(PR, P, s->tf[which].p, s->dlut[which].p, s->cs.p, nullptr, reverse);
    s->counters[0]++;
}

// ---- compact sector storage: one launch per pass over the matrix C[alpha][beta] ---------------
static void csweep_dims(const mqc_solver* s, int which, int pass, bool adj, int& nt, unsigned& grid, size_t& smem) {
    const MqcCPlan& PL = s->cplan[which];
    const MqcCPassDev& P = PL.pass[pass];
    nt = s->csweep_threads;
    if (adj) nt = s->csweep_adj_threads > 0 ? s->csweep_adj_threads : (s->creal ? std::min(64, s->csweep_threads) : s->csweep_threads);
    // small tiles do not feed many threads: most factor visits of a tile couple <= capA * capB / 4 pairs
    while (nt > 128 && P.capA * P.capB < 2 * nt) nt >>= 1;
    smem = mqc_c_smem_bytes(P, adj, nt / 32, PL.chunk, PL.elem);
    if (smem > 232448) throw MqcError("compact tile does not fit shared memory");
    int per_sm = (int)std::min<size_t>(32, (size_t)(227 * 1024) / (smem + 1024));
    per_sm = std::max(1, std::min(per_sm, std::min(32, 2048 / nt)));
    // a few thousand tiles: one CTA per tile, the hardware deals them out as SM slots free up (measured 1-3 % better
    // than persistent CTAs striding over the list); more: persistent CTAs (one gradient flush per CTA, not per tile)
    grid = (unsigned)std::min<u64>(P.n_tiles, std::max<u64>((u64)s->n_sm * per_sm, 2048));
    if (s->csweep_ctas > 0) grid = (unsigned)std::min<u64>(P.n_tiles, (u64)s->csweep_ctas);
}
// per-(CTA, factor) gradient partials of the adjoint passes of schedule 1 and where k_grad_combine_parts finds them
static void ensure_gpart(mqc_solver* s) {
    if (s->gpart_ready || !s->deterministic) return;
    const MqcCPlan& PL = s->cplan[1];
    const int np = (int)PL.pass.size();
    const size_t nfac = s->fac.size();
    std::vector<uint4> where(std::max<size_t>(1, nfac), make_uint4(0u, 0u, 1u, 0u));
    s->gp_off.assign(std::max(1, np), 0);
    size_t total = 0;
    for (int p = 0; p < np; ++p) {
        int nt; unsigned grid; size_t smem;
        csweep_dims(s, 1, p, true, nt, grid, smem);
        const MqcCPassDev& P = PL.pass[p];
        s->gp_off[p] = total;
        for (int f = 0; f < P.nf; ++f) {
            const size_t slot = (size_t)(PL.fac[P.fac_off + f].theta_sign & 0x7fffffff);
            const unsigned long long o = total + (size_t)f;
            if (slot < nfac) where[slot] = make_uint4((unsigned)(o & 0xffffffffull), (unsigned)(o >> 32), (unsigned)P.nf, grid);
        }
        total += (size_t)grid * P.nf;
    }
    CK(cudaSetDevice(s->device));
    if (s->cgpart.n < std::max<size_t>(1, total)) s->cgpart.alloc(std::max<size_t>(1, total));
    s->cgwhere.upload(where, s->stream);
    CK(cudaStreamSynchronize(s->stream));
    s->gpart_ready = true;
}
static void launch_csweep(mqc_solver* s, int which, int pass, bool adj) {
    const MqcCPlan& PL = s->cplan[which];
    const MqcCPassDev& P = PL.pass[pass];
    const CPlanDev& D = s->cdev[which];
    MqcCArgs A{};
    A.pass = D.pass.p + pass; A.cls = D.cls.p; A.fac = D.fac.p; A.maps = D.maps.p; A.tiles = D.tiles.p;
    A.live_tab = D.live_tab.p; A.prog_hdr = D.prog_hdr.p; A.prog = D.prog.p; A.cs = s->cs.p;
    A.psi_in = s->cpsi[s->ccur].p; A.psi_out = s->cpsi[s->ccur ^ 1].p;
    A.lam_in = adj ? (const void*)s->clam[s->clcur].p : nullptr; A.lam_out = adj ? (void*)s->clam[s->clcur ^ 1].p : nullptr;
    A.grad = adj ? s->grad_p() : nullptr;
    A.nB = PL.pitch;                                   // the kernel only uses it as the row pitch
    int nt; unsigned grid; size_t smem;
    csweep_dims(s, which, pass, adj, nt, grid, smem);
    const bool real = s->creal;
    if (adj && s->deterministic) {
        if (!s->gpart_ready) throw MqcError("internal: gradient partial buffers are not set up");
        A.gpart = s->cgpart.p + s->gp_off[pass];
    }
    CK(cudaSetDevice(s->device));
    // programmatic dependent launch: the CTAs of this pass may start (and set up their tables) while the previous
    // kernel of the stream drains; k_csweep waits (griddepcontrol.wait) before it touches the state
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute pdl_attr[1];
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(grid); cfg.dynamicSmemBytes = smem; cfg.stream = s->stream;
    cfg.attrs = pdl_attr; cfg.numAttrs = s->csweep_pdl ? 1 : 0;
#define MQC_CS_LAUNCH3(ADJ_, NT_, CH_, REAL_) do { cfg.blockDim = dim3(NT_); CK(cudaLaunchKernelEx(&cfg, k_csweep<ADJ_, NT_, CH_, REAL_>, A)); } while (0)
#define MQC_CS_LAUNCH2(NT_, CH_) do { \
        if (real) { if (adj) MQC_CS_LAUNCH3(true, NT_, CH_, true); else MQC_CS_LAUNCH3(false, NT_, CH_, true); } \
        else { if (adj) MQC_CS_LAUNCH3(true, NT_, CH_, false); else MQC_CS_LAUNCH3(false, NT_, CH_, false); } } while (0)
#define MQC_CS_LAUNCH(NT_) do { if (PL.chunk == 256) MQC_CS_LAUNCH2(NT_, 256); else MQC_CS_LAUNCH2(NT_, 1024); } while (0)
    if (nt >= 512) MQC_CS_LAUNCH(512); else if (nt >= 256) MQC_CS_LAUNCH(256); else if (nt >= 128) MQC_CS_LAUNCH(128);
    else if (nt >= 64) MQC_CS_LAUNCH(64); else MQC_CS_LAUNCH(32);
    s->counters[0]++;
    s->ccur ^= 1;
    if (adj) s->clcur ^= 1;
}

static void forward_compact(mqc_solver* s, int which, const double* theta) {
    const int nfac = (int)s->fac.size();
    const MqcCPlan& PL = s->cplan[which];
    CK(cudaSetDevice(s->device));
    if (theta) {
        s->theta_host.assign(theta, theta + s->n_theta);
        CK(cudaMemcpyAsync(s->theta.p, s->theta_host.data(), sizeof(double) * s->n_theta, cudaMemcpyHostToDevice, s->stream));
    }
    if (nfac) { k_sincos<<<(nfac + 255) / 256, 256, 0, s->stream>>>(s->theta.p, s->theta_index.p, s->theta_scale.p, s->cs.p, nfac); s->counters[0]++; }
    s->ccur = 0;
    const size_t cnt = (size_t)PL.nA * PL.pitch;
    CK(cudaMemsetAsync(s->cpsi[0].p, 0, (s->creal ? sizeof(double) : sizeof(double2)) * cnt, s->stream));
    if (s->creal) k_set_real<<<1, 32, 0, s->stream>>>(reinterpret_cast<double*>(s->cpsi[0].p), (u64)PL.ref_row * PL.pitch + PL.ref_col, 1.0);
    else k_set_amplitude<<<1, 32, 0, s->stream>>>(s->cpsi[0].p, (u64)PL.ref_row * PL.pitch + PL.ref_col, 1.0, 0.0);
    s->counters[0]++;
    const int np = (int)PL.pass.size();
    for (int p = 0; p < np; ++p) launch_csweep(s, which, p, false);
    s->counters[1] += np;
    s->cvalid = true; s->cwhich = which;
    s->state_layout = -1;
}

static void forward(mqc_solver* L, int which, const double* theta) {
    if (!L->have_ansatz) throw MqcError("mqc_set_ansatz has not been called");
    require_connected(L);
    if (L->use_compact) { forward_compact(L, which, theta); return; }
    const int nfac = (int)L->fac.size();
    for (mqc_solver* s : L->members) {
        CK(cudaSetDevice(s->device));
        if (theta) {
            s->theta_host.assign(theta, theta + s->n_theta);
            CK(cudaMemcpyAsync(s->theta.p, s->theta_host.data(), sizeof(double) * s->n_theta, cudaMemcpyHostToDevice, s->stream));
        }
        if (nfac) { k_sincos<<<(nfac + 255) / 256, 256, 0, s->stream>>>(s->theta.p, s->theta_index.p, s->theta_scale.p, s->cs.p, nfac); s->counters[0]++; }
        CK(cudaMemsetAsync(s->psi.p, 0, sizeof(double2) * s->dim, s->stream));
        const u64 ref_phys = (s->mode == 1) ? mqc_map_mask(s->ref_index, s->sched[which].phys_init) : s->ref_index;
        if ((int)(ref_phys >> s->n_local) == s->rank) {
            k_set_amplitude<<<1, 32, 0, s->stream>>>(s->psi.p, ref_phys & (s->dim - 1ull), 1.0, 0.0);
            s->counters[0]++;
        }
    }
    if (L->mode == 1) {
        const int np = (int)L->sched[which].passes.size();
        bool prev_g = false;
        for (int p = 0; p < np; ++p) {
            const bool g = pass_is_global(L, which, p);
            if (g || prev_g) group_sync(L);
            for (mqc_solver* s : L->members) launch_tile(s, which, p, false, 0);
            prev_g = g;
        }
        L->counters[1] += np;
    } else {
        mqc_solver* s = L;
        for (int k = 0; k < nfac; ++k) {
            const MqcFactor& F = s->fac[k];
            const u64 npairs = s->dim >> F.npos;
            const unsigned grid = (unsigned)std::min<u64>((npairs + 255) / 256, 148ull * 16);
            k_excite_direct<<<grid, 256, 0, s->stream>>>(s->psi.p, F, s->cs.p, 0, npairs);
            s->counters[0]++;
        }
        s->counters[1] += nfac;
    }
    group_sync(L);          // every shard complete before anyone reads a peer's
    for (mqc_solver* s : L->members) s->state_layout = which;
}

// lambda = H psi (optional) and this group's energy partials in every member's scalars
struct HamView { const HamTables* T; const SecHamTables* ST; const SigmaTables* SG; };
static void sector_dims(const mqc_solver* s, unsigned& nA, unsigned& nB) {
    nA = (unsigned)(binom_d((s->n + 1) / 2, s->na) + 0.5);
    nB = (unsigned)(binom_d(s->n / 2, s->nb) + 0.5);
}
static void rank_rows(int n_ranks, int rank, unsigned n_rows, unsigned& r0, unsigned& r1) {
    r0 = (unsigned)(((u64)n_rows * rank) / n_ranks);
    r1 = (unsigned)(((u64)n_rows * (rank + 1)) / n_ranks);
}
// row-distributed sector matrix: every rank keeps only its own rows of C (and of Lambda)
static void alloc_blocked(mqc_solver* s) {
    unsigned nA, nB, r0, r1;
    sector_dims(s, nA, nB);
    rank_rows(s->n_ranks, s->rank, nA, r0, r1);
    unsigned maxrows = 0;
    for (int r = 0; r < s->n_ranks; ++r) { unsigned a, b2; rank_rows(s->n_ranks, r, nA, a, b2); maxrows = std::max(maxrows, b2 - a); }
    CK(cudaSetDevice(s->device));
    if (!s->cown.p) s->cown.alloc(std::max<u64>(1, (u64)(r1 - r0) * nB));
    if (!s->cblk.p) s->cblk.alloc(std::max<u64>(1, (u64)maxrows * nB));
}

static void launch_sigma(mqc_solver* s, const SigmaTables* SG, const double2* Cown, const double2* Csrc, double2* Lp,
                         unsigned nB, unsigned r0, unsigned r1, unsigned src_lo, unsigned src_hi, unsigned own_lo,
                         unsigned l_lo, int accumulate, bool may_split = false, bool real = false, unsigned pitch = 0) {
    // real: the three buffers hold doubles (compact storage with real amplitudes)
    MqcSigmaDev D = SG->dev;
    D.src_lo = src_lo; D.src_hi = src_hi; D.own_lo = own_lo; D.l_lo = l_lo; D.accumulate = accumulate;
    D.pitch = pitch ? pitch : nB;
    const int elem = real ? 8 : 16;
    // source rows staged in shared memory, four links at a time; longer rows two at a time (28 qubits with real
    // amplitudes: K2 36.2 -> 32.9 ms); one at a time is slower than unstaged gathers from L1 / L2 (30 qubits: 171
    // against 151 ms) and only kept for tests (sigma_stage_one = 1)
    int lb = MQC_SG_LB;
    const size_t cap = (size_t)s->sigma_stage_bytes;
    while (lb > (s->sigma_stage_one ? 1 : 2) && (size_t)lb * nB * elem > cap) lb >>= 1;
    const bool stage = s->sigma_stage && (size_t)lb * nB * elem <= cap;
    if (!stage) lb = MQC_SG_LB;
    const size_t smem = mqc_sigma_smem_bytes(D, stage, lb, elem);
    dim3 grid(r1 - r0, (nB + MQC_SG_THREADS * MQC_SG_R - 1) / (MQC_SG_THREADS * MQC_SG_R));
    // Few CTAs per SM (924 rows on 444 slots at 24 qubits: the third wave is 8 % full): halve the work per CTA.
    // sigma_split: -1 auto, 0 off, k > 0: part 0 takes the first k alpha single links.
    D.split_l = 0;
    if (may_split && !accumulate && s->sigma_split != 0 && D.e1a >= 2 * MQC_SG_LB) {
        const u64 ctas = (u64)grid.x * grid.y;
        if (s->sigma_split > 0 || ctas < (u64)24 * s->n_sm) {
            int k = s->sigma_split > 0 ? s->sigma_split : (int)(0.6 * D.e1a);
            k = std::max(MQC_SG_LB, std::min(k / MQC_SG_LB * MQC_SG_LB, (D.e1a - 1) / MQC_SG_LB * MQC_SG_LB));
            D.split_l = k;
            grid.z = 2;
            if (Lp) CK(cudaMemsetAsync(Lp, 0, (size_t)(r1 - r0) * D.pitch * elem, s->stream));
        }
    }
    D.e_part = nullptr; D.ticket = nullptr;
    if (s->deterministic) {
        const size_t ctas = (size_t)grid.x * grid.y * grid.z;
        if (s->e_part.n < 2 * ctas) {
            if (s->capturing) throw MqcError("internal: energy partial buffer grows during graph capture");
            invalidate_graphs(s);                        // captured launches hold the old pointer
            s->e_part.alloc(2 * ctas);
        }
        if (!s->e_ticket.p) { s->e_ticket.alloc(1); CK(cudaMemsetAsync(s->e_ticket.p, 0, sizeof(unsigned), s->stream)); }
        D.e_part = s->e_part.p; D.ticket = s->e_ticket.p;
    }
#define MQC_SG_LAUNCH(LB_) do { \
        if (real) k_sigma<true, LB_, true><<<grid, MQC_SG_THREADS, smem, s->stream>>>(reinterpret_cast<const double*>(Cown), \
                      reinterpret_cast<const double*>(Csrc), reinterpret_cast<double*>(Lp), D, s->scalars_p()); \
        else k_sigma<true, LB_, false><<<grid, MQC_SG_THREADS, smem, s->stream>>>(Cown, Csrc, Lp, D, s->scalars_p()); } while (0)
    if (stage) { if (lb == 4) MQC_SG_LAUNCH(4); else if (lb == 2) MQC_SG_LAUNCH(2); else MQC_SG_LAUNCH(1); }
    else if (real) k_sigma<false, MQC_SG_LB, true><<<grid, MQC_SG_THREADS, smem, s->stream>>>(reinterpret_cast<const double*>(Cown),
                       reinterpret_cast<const double*>(Csrc), reinterpret_cast<double*>(Lp), D, s->scalars_p());
    else k_sigma<false, MQC_SG_LB, false><<<grid, MQC_SG_THREADS, smem, s->stream>>>(Cown, Csrc, Lp, D, s->scalars_p());
    s->counters[0]++;
}

// K2 directly on the compact matrix: C = cpsi[ccur] (canonical order), Lambda -> clam[0]
static void happly_compact(mqc_solver* s, int which, const HamView& view, bool want_lam) {
    LayoutTables& Lt = s->lay[which];
    const unsigned nA = s->cplan[which].nA, nB = s->cplan[which].nB;
    CK(cudaSetDevice(s->device));
    CK(cudaMemsetAsync(s->scalars_p(), 0, 4 * sizeof(double), s->stream));
    const double2* C = s->cpsi[s->ccur].p;
    s->clcur = 0;
    double2* Lm = want_lam ? s->clam[0].p : nullptr;
    const SigmaTables* SG = view.SG;
    bool have = false;
    bool real = s->creal;
    if (real && !(SG && SG->ready && !SG->n_fallback && !SG->ab_ready)) {
        // an operator with a part outside the link tables (mqc_expectation of arbitrary Pauli terms): expectation
        // value on a complex copy of the state.  The handle's own Hamiltonian never gets here (ensure_layout_ready).
        if (want_lam) throw MqcError("internal: real amplitudes with a Hamiltonian part outside the link tables");
        C = compact_complex_view(s);
        real = false;
    }
    if (SG && SG->ready) { launch_sigma(s, SG, C, C, Lm, nB, 0u, nA, 0u, nA, 0u, 0u, 0, true, real, real ? s->cplan[which].pitch : nB); have = true; }
    if (!have || SG->n_fallback) {
        unsigned bd = 256;
        while (bd > 32 && (u64)(bd / 2) * MQC_HS_R >= nB) bd >>= 1;
        dim3 grid(nA, (nB + bd * MQC_HS_R - 1) / (bd * MQC_HS_R));
        k_happly_sector<<<grid, bd, 0, s->stream>>>(C, Lm, sec_ham_dev(Lt, *view.ST), s->scalars_p(), 0u, have ? 1 : 0);
        s->counters[0]++;
    }
    if (SG && SG->ready && SG->ab_ready) {                // opposite-spin doubles: L += H_ab C (mqc_sigma_ab.h)
        const bool stage = s->sigma_stage && mqc_sigma_ab_smem_bytes(nB, true) <= 113 * 1024;
        const size_t smem = mqc_sigma_ab_smem_bytes(nB, stage);
        dim3 grid(nA, (nB + MQC_AB_THREADS * MQC_AB_R - 1) / (MQC_AB_THREADS * MQC_AB_R));
        if (stage) k_sigma_ab<true><<<grid, MQC_AB_THREADS, smem, s->stream>>>(C, Lm, SG->ab, s->scalars_p());
        else k_sigma_ab<false><<<grid, MQC_AB_THREADS, smem, s->stream>>>(C, Lm, SG->ab, s->scalars_p());
        s->counters[0]++;
    }
}

static void happly(mqc_solver* L, int which, const std::vector<HamView>& views, bool want_lam) {
    if (L->use_compact && L->cvalid && L->state_layout < 0) { happly_compact(L, which, views[0], want_lam); return; }
    const bool sec = L->conserving && L->use_sector && views[0].ST;
    const bool blocked = sec && L->blocked;
    if (blocked) {
        // Row-distributed C: gather the own rows, then G rounds - round b applies the links whose
        // source rows belong to rank (rank + b) % G, whose block is copied over NVLink first.
        for (size_t i = 0; i < L->members.size(); ++i) {
            const SigmaTables* SG = views[i].SG;
            if (!SG || !SG->ready || SG->n_fallback)
                throw MqcError("row-distributed H|psi> needs a Hamiltonian the link tables cover completely (option sigma_blocked=0 replicates C)");
        }
        for (mqc_solver* s : L->members) {
            LayoutTables& Lt = s->lay[which];
            unsigned nA = (unsigned)Lt.astr.n, nB = (unsigned)Lt.bstr.n, r0, r1;
            rank_rows(s->n_ranks, s->rank, nA, r0, r1);
            alloc_blocked(s);
            if (!s->connected_cown) throw MqcError("row-distributed H|psi>: peers' row blocks are not mapped (set sigma_blocked before mqc_shard_export)");
            CK(cudaSetDevice(s->device));
            CK(cudaMemsetAsync(s->scalars_p(), 0, 4 * sizeof(double), s->stream));
            const u64 cnt = (u64)(r1 - r0) * nB;
            if (cnt) k_sector_gather<<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(peers_of(s), s->cown.p, Lt.astr.p, Lt.bstr.p, r0, r1, nB, r0);
            if (want_lam) {
                if (s->lsec.n < std::max<u64>(1, cnt)) s->lsec.alloc(std::max<u64>(1, cnt));
                CK(cudaMemsetAsync(s->lam.p, 0, sizeof(double2) * s->dim, s->stream));
            }
            s->counters[0]++;
        }
        group_sync(L);      // every rank's rows are complete
        for (int b = 0; b < L->n_ranks; ++b) {
            for (size_t i = 0; i < L->members.size(); ++i) {
                mqc_solver* s = L->members[i];
                LayoutTables& Lt = s->lay[which];
                unsigned nA = (unsigned)Lt.astr.n, nB = (unsigned)Lt.bstr.n, r0, r1, q0, q1;
                rank_rows(s->n_ranks, s->rank, nA, r0, r1);
                const int owner = (s->rank + b) % s->n_ranks;
                rank_rows(s->n_ranks, owner, nA, q0, q1);
                if (r1 == r0 || q1 == q0) continue;
                CK(cudaSetDevice(s->device));
                const double2* src = s->cown.p;
                if (owner != s->rank) {
                    CK(cudaMemcpyAsync(s->cblk.p, s->peer_cown[owner], sizeof(double2) * (size_t)(q1 - q0) * nB, cudaMemcpyDeviceToDevice, s->stream));
                    src = s->cblk.p;
                }
                launch_sigma(s, views[i].SG, s->cown.p, src, want_lam ? s->lsec.p : nullptr, nB, r0, r1, q0, q1, r0, r0, b > 0 ? 1 : 0);
            }
        }
        if (want_lam) {
            group_sync(L);  // all shards of lambda cleared
            for (mqc_solver* s : L->members) {
                LayoutTables& Lt = s->lay[which];
                unsigned nA = (unsigned)Lt.astr.n, nB = (unsigned)Lt.bstr.n, r0, r1;
                rank_rows(s->n_ranks, s->rank, nA, r0, r1);
                if (r1 == r0) continue;
                CK(cudaSetDevice(s->device));
                const u64 cnt = (u64)(r1 - r0) * nB;
                k_sector_scatter<<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(peers_of(s), s->lsec.p, Lt.astr.p, Lt.bstr.p, r0, r1, nB, r0);
                s->counters[0]++;
            }
        }
        group_sync(L);      // peers are done reading this rank's rows
        return;
    }
    for (size_t i = 0; i < L->members.size(); ++i) {
        mqc_solver* s = L->members[i];
        LayoutTables& Lt = s->lay[which];
        CK(cudaSetDevice(s->device));
        CK(cudaMemsetAsync(s->scalars_p(), 0, 4 * sizeof(double), s->stream));
        const MqcPeers PR = peers_of(s);
        if (sec) {
            const unsigned nA = (unsigned)Lt.astr.n, nB = (unsigned)Lt.bstr.n;
            const u64 cnt = (u64)nA * nB;
            if (s->csec.n < cnt) s->csec.alloc(cnt);
            if (want_lam && s->lsec.n < cnt) s->lsec.alloc(cnt);
            const unsigned gl = (unsigned)((cnt + 255) / 256);
            k_sector_gather<<<gl, 256, 0, s->stream>>>(PR, s->csec.p, Lt.astr.p, Lt.bstr.p, 0u, nA, nB, 0u);
            unsigned r0, r1;
            row_slice(s, nA, r0, r1);
            if (r1 > r0) {
                const SigmaTables* SG = views[i].SG;
                bool have = false;
                if (SG && SG->ready) {
                    launch_sigma(s, SG, s->csec.p, s->csec.p, want_lam ? s->lsec.p : nullptr, nB, r0, r1, 0u, nA, 0u, 0u, 0);
                    have = true;
                }
                if (!have || SG->n_fallback) {
                    unsigned bd = 256;
                    while (bd > 32 && (u64)(bd / 2) * MQC_HS_R >= nB) bd >>= 1;
                    dim3 grid(r1 - r0, (nB + bd * MQC_HS_R - 1) / (bd * MQC_HS_R));
                    k_happly_sector<<<grid, bd, 0, s->stream>>>(s->csec.p, want_lam ? s->lsec.p : nullptr,
                                                                sec_ham_dev(Lt, *views[i].ST), s->scalars_p(), r0, have ? 1 : 0);
                    s->counters[0]++;
                }
            }
            s->counters[0] += 1;
            if (want_lam) CK(cudaMemsetAsync(s->lam.p, 0, sizeof(double2) * s->dim, s->stream));
        } else {
            MqcSectorDev S = sector_dev(Lt, false);
            const unsigned grid = (unsigned)((s->dim + 255) / 256);
            k_happly<<<grid, 256, 0, s->stream>>>(PR, want_lam ? s->lam.p : nullptr, views[i].T->dev(), S, s->dim,
                                                  (u64)s->rank << s->n_local, s->scalars_p());
            s->counters[0]++;
        }
    }
    if (sec && want_lam) {
        group_sync(L);      // all shards of lambda cleared before rows are scattered into them
        for (mqc_solver* s : L->members) {
            LayoutTables& Lt = s->lay[which];
            const unsigned nB = (unsigned)Lt.bstr.n;
            unsigned r0, r1;
            row_slice(s, (unsigned)Lt.astr.n, r0, r1);
            if (r1 == r0) continue;
            CK(cudaSetDevice(s->device));
            const u64 cnt = (u64)(r1 - r0) * nB;
            k_sector_scatter<<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(peers_of(s), s->lsec.p, Lt.astr.p, Lt.bstr.p, r0, r1, nB, 0u);
            s->counters[0]++;
        }
    }
    group_sync(L);
}

static void reverse_sweep(mqc_solver* L, int which) {
    for (mqc_solver* s : L->members) {
        CK(cudaSetDevice(s->device));
        CK(cudaMemsetAsync(s->grad_p(), 0, sizeof(double) * std::max<size_t>(1, s->fac.size()), s->stream));
    }
    const int nfac = (int)L->fac.size();
    if (L->use_compact) {
        const int np = (int)L->cplan[which].pass.size();
        ensure_gpart(L);
        for (int p = np - 1; p >= 0; --p) launch_csweep(L, which, p, true);
        L->counters[2] += np;
        L->cvalid = false;
        L->state_layout = -1;
        return;
    }
    if (L->mode == 1) {
        const int np = (int)L->sched[which].passes.size();
        bool prev_g = false;
        for (int p = np - 1; p >= 0; --p) {
            const bool g = pass_is_global(L, which, p);
            if (g || prev_g) group_sync(L);
            for (mqc_solver* s : L->members) launch_tile(s, which, p, true, 1);
            prev_g = g;
        }
        L->counters[2] += np;
    } else {
        mqc_solver* s = L;
        for (int k = nfac - 1; k >= 0; --k) {
            const MqcFactor& F = s->fac[k];
            const u64 npairs = s->dim >> F.npos;
            const unsigned grid = (unsigned)std::min<u64>((npairs + 255) / 256, 148ull * 16);
            k_adjoint_direct<<<grid, 256, 0, s->stream>>>(s->psi.p, s->lam.p, F, s->cs.p, s->grad_p(), npairs);
            s->counters[0]++;
        }
        s->counters[2] += nfac;
    }
    for (mqc_solver* s : L->members) s->state_layout = -1;
}

// Sum every rank's mail[0 .. count) into the leader's `red` buffer (all ranks do it for
// themselves when each is its own process: an all-reduce over peer memory).
static void reduce_mail(mqc_solver* L, int count) {
    if (L->n_ranks == 1) return;
    group_sync(L);
    MqcSumPeers SP{};
    SP.n_ranks = L->n_ranks;
    for (int r = 0; r < L->n_ranks; ++r) SP.src[r] = L->peer_mail[r];
    CK(cudaSetDevice(L->device));
    if (L->red.n < (size_t)count) L->red.alloc(count);
    k_sum_peers<<<(count + 255) / 256, 256, 0, L->stream>>>(SP, L->red.p, count);
    L->counters[0]++;
    group_sync(L);          // nobody clears its mail before every peer has read it
}
static const double* reduced_scalars(const mqc_solver* L) { return L->n_ranks == 1 ? L->scalars_p() : L->red.p; }
static const double* reduced_grad(const mqc_solver* L) { return L->n_ranks == 1 ? L->grad_p() : L->red.p + MQC_MAIL_GRAD; }

static void sync_all(mqc_solver* L) {
    for (mqc_solver* s : L->members) { CK(cudaSetDevice(s->device)); CK(cudaStreamSynchronize(s->stream)); }
    if (L->comm_mode == 2 && L->err.p) {
        int e = 0;
        CK(cudaMemcpy(&e, L->err.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) throw MqcError("rank barrier timed out: a peer rank is not making progress");
    }
    CK(cudaSetDevice(L->device));
}

static void reset_counters(mqc_solver* L) {
    for (mqc_solver* s : L->members) {
        s->counters[0] = s->counters[1] = s->counters[2] = 0;
        s->ms[0] = s->ms[1] = s->ms[2] = s->ms[3] = 0;
    }
}

static void finish_timing(mqc_solver* s, bool with_reverse) {
    float t = 0;
    CK(cudaEventElapsedTime(&t, s->ev[0], s->ev[1])); s->ms[0] = t;
    CK(cudaEventElapsedTime(&t, s->ev[1], s->ev[2])); s->ms[1] = t;
    if (with_reverse) { CK(cudaEventElapsedTime(&t, s->ev[2], s->ev[3])); s->ms[2] = t; }
    CK(cudaEventElapsedTime(&t, s->ev[0], with_reverse ? s->ev[3] : s->ev[2])); s->ms[3] = t;
    for (size_t i = 1; i < s->members.size(); ++i) s->counters[0] += s->members[i]->counters[0];
}

static std::vector<HamView> own_views(mqc_solver* L, int which) {
    std::vector<HamView> v;
    for (mqc_solver* s : L->members)
        v.push_back(HamView{&s->lay[which].ham, &s->lay[which].sham, (s->use_sigma && s->sig.ready) ? &s->sig : nullptr});
    return v;
}

static void alloc_state(mqc_solver* s, bool want_lam) {
    CK(cudaSetDevice(s->device));
    if (!s->psi.p) s->psi.alloc(s->dim);
    if (want_lam && !s->lam.p) s->lam.alloc(s->dim);
    if (s->n_ranks == 1) { s->peer_psi[0] = s->psi.p; s->peer_lam[0] = s->lam.p; }
}

// same-process group: every member sees every member's buffers (peer access across devices)
static void connect_group(mqc_solver* L, bool want_lam) {
    if (L->comm_mode != 1) return;
    bool need = !L->connected;
    for (mqc_solver* s : L->members) if (!s->psi.p || (want_lam && !s->lam.p)) need = true;
    if (!need) return;
    for (mqc_solver* s : L->members) { alloc_state(s, want_lam); if (s->blocked) alloc_blocked(s); }
    for (mqc_solver* s : L->members) {
        CK(cudaSetDevice(s->device));
        for (mqc_solver* q : L->members) {
            if (q->device != s->device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, s->device, q->device));
                if (!can) throw MqcError("GPUs of a shard group cannot access each other's memory");
                cudaError_t e = cudaDeviceEnablePeerAccess(q->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cuda_check(e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
            s->peer_psi[q->rank] = q->psi.p;
            s->peer_lam[q->rank] = q->lam.p;
            s->peer_mail[q->rank] = q->mail.p;
            s->peer_cown[q->rank] = q->cown.p;
        }
        s->connected = true;
        s->connected_cown = s->blocked;
    }
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

const char* mqc_last_error(void) { return g_err.c_str(); }
int mqc_version(void) { return 110; }

int mqc_device_count(int* count) {
    MQC_TRY
    if (!count) throw MqcError("null argument");
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) c = 0;
    *count = c;
    MQC_CATCH
}

static mqc_solver* new_rank(int n_qubits, int device, int rank, int n_ranks) {
    if (n_qubits < 2 || n_qubits > 40) throw MqcError("n_qubits must be in [2, 40]");
    if (n_ranks < 1 || n_ranks > MQC_MAX_RANKS || (n_ranks & (n_ranks - 1))) throw MqcError("n_ranks must be 1, 2, 4 or 8");
    if (rank < 0 || rank >= n_ranks) throw MqcError("rank out of range");
    int g = 0;
    while ((1 << g) < n_ranks) ++g;
    if (n_qubits - g < 5) throw MqcError("too few qubits per shard");
    mqc_solver* s = new mqc_solver();
    s->n = n_qubits;
    s->n_global = g;
    s->n_local = n_qubits - g;
    s->rank = rank;
    s->n_ranks = n_ranks;
    s->device = device;
    s->dim = 1ull << s->n_local;
    s->members.push_back(s);
    return s;
}

int mqc_create(int n_qubits, int device, mqc_solver_t** out) {
    MQC_TRY
    if (!out) throw MqcError("null argument");
    *out = new_rank(n_qubits, device, 0, 1);
    MQC_CATCH
}

int mqc_shard_create(int n_qubits, int device, int rank, int n_ranks, mqc_solver_t** out) {
    MQC_TRY
    if (!out) throw MqcError("null argument");
    mqc_solver* s = new_rank(n_qubits, device, rank, n_ranks);
    s->comm_mode = n_ranks > 1 ? 2 : 0;
    *out = s;
    MQC_CATCH
}

int mqc_shard_group_create(int n_qubits, int n_ranks, const int* devices, mqc_solver_t** out) {
    MQC_TRY
    if (!out || !devices) throw MqcError("null argument");
    mqc_solver* L = new_rank(n_qubits, devices[0], 0, n_ranks);
    L->comm_mode = n_ranks > 1 ? 1 : 0;
    for (int r = 1; r < n_ranks; ++r) {
        mqc_solver* m = new_rank(n_qubits, devices[r], r, n_ranks);
        m->comm_mode = 1;
        L->members.push_back(m);
    }
    *out = L;
    MQC_CATCH
}

static void destroy_rank(mqc_solver* s) {
    if (s->gpu_ready) {
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
        for (auto& e : s->ev) if (e) cudaEventDestroy(e);
        for (auto& e : s->user_ev) if (e) cudaEventDestroy(e);
        if (s->sync_ev) cudaEventDestroy(s->sync_ev);
        invalidate_graphs(s);
        for (auto& kv : s->ipc_open) cudaIpcCloseMemHandle(kv.second);
    }
    cudaStream_t st = s->stream;
    const bool ready = s->gpu_ready;
    const int dev = s->device;
    delete s;
    if (ready && st) { cudaSetDevice(dev); cudaStreamDestroy(st); }
}

int mqc_destroy(mqc_solver_t* s) {
    MQC_TRY
    if (!s) return 0;
    std::vector<mqc_solver*> m = s->members;
    for (size_t i = m.size(); i-- > 0;) destroy_rank(m[i]);
    MQC_CATCH
}

static void set_option_rank(mqc_solver* s, const std::string& k, int64_t value) {
    if (k == "tile_bits") s->tile_bits = (int)value;
    else if (k == "low_bits") s->low_bits = (int)value;
    else if (k == "grad_tile_bits") s->grad_tile_bits = (int)value;
    else if (k == "mode") s->mode = (int)value;
    else if (k == "sector") s->use_sector = (int)value;
    else if (k == "tile_threads") s->tile_threads = (int)value;
    else if (k == "sparse_tiles") s->sparse_tiles = (int)value;
    else if (k == "sparse_threads") s->sparse_threads = (int)value;
    else if (k == "sigma") s->use_sigma = (int)value;
    else if (k == "sigma_stage") s->sigma_stage = (int)value;
    else if (k == "sigma_split") s->sigma_split = (int)value;
    else if (k == "sigma_stage_one") s->sigma_stage_one = (int)value;
    else if (k == "sigma_stage_bytes") { if (value < 0 || value > 98304) throw MqcError("sigma_stage_bytes must be 0 .. 98304"); s->sigma_stage_bytes = (int)value; }
    else if (k == "real_amplitudes") s->real_amplitudes = (int)value;
    else if (k == "csweep_adj_threads") s->csweep_adj_threads = (int)value;
    else if (k == "csweep_pdl") s->csweep_pdl = (int)value;
    else if (k == "deterministic") s->deterministic = (int)value;
    else if (k == "sigma_ab") s->sigma_ab = (int)value;
    else if (k == "graph") s->use_graph = (int)value;
    else if (k == "sigma_blocked") s->sigma_blocked = (int)value;
    else if (k == "compact") s->compact = (int)value;
    else if (k == "rdm_gram") s->rdm_gram = (int)value;
    else if (k == "csweep_threads") { if (value != 32 && value != 64 && value != 128 && value != 256 && value != 512) throw MqcError("csweep_threads must be 32, 64, 128, 256 or 512"); s->csweep_threads = (int)value; }
    else if (k == "csweep_ctas") s->csweep_ctas = (int)value;
    else if (k == "csweep_snake") s->csweep_snake = (int)value;
    else if (k == "sp2") s->sp2 = (int)value;
    else if (k == "sp2_list") s->sp2_list = (int)value;
    else if (k == "sp2_team") { if (value != 0 && value != 1 && value != 2 && value != 4) throw MqcError("sp2_team must be 0 (auto), 1, 2 or 4"); s->sp2_team = (int)value; }
    else if (k == "sp2_smem_kb") { if (value < 16 || value > 220) throw MqcError("sp2_smem_kb must be in [16,220]"); s->sp2_smem_kb = (int)value; }
    else if (k == "sp2_warps") { if (value < 1 || value > 8) throw MqcError("sp2_warps must be in [1,8]"); s->sp2_warps = (int)value; }
    else throw MqcError("unknown option: " + k);
    if (s->tile_bits < 5 || s->tile_bits > 16 || s->grad_tile_bits < 5 || s->grad_tile_bits > 16)
        throw MqcError("tile_bits and grad_tile_bits must be in [5,16] (compact sector storage: <= 16; sector tiles of the dense register: <= 14; dense tiles: <= 13 / 12)");
    if (s->low_bits < -1 || s->low_bits > 6) throw MqcError("low_bits must be in [0,6] (-1: 3 for dense tiles, 0 for sector tiles)");
    if (s->tile_threads < 32 || s->tile_threads > 512 || (s->tile_threads & (s->tile_threads - 1)))
        throw MqcError("tile_threads must be a power of two in [32,512]");
    if (s->sparse_threads < 32 || s->sparse_threads > 256 || (s->sparse_threads & (s->sparse_threads - 1)))
        throw MqcError("sparse_threads must be a power of two in [32,256]");
}

int mqc_set_option(mqc_solver_t* s, const char* key, int64_t value) {
    MQC_TRY
    if (!s || !key) throw MqcError("null argument");
    // options that shape the schedules, layouts, plans and uploaded tables are frozen by mqc_set_ansatz
    static const char* frozen[] = {"tile_bits", "low_bits", "grad_tile_bits", "mode", "sector", "sparse_tiles", "sp2", "compact",
                                   "sigma_blocked", "csweep_snake", "real_amplitudes", nullptr};
    if (s->have_ansatz)
        for (const char** f = frozen; *f; ++f)
            if (std::string(key) == *f)
                throw MqcError(std::string("option '") + key + "' must be set before mqc_set_ansatz (it shapes the schedules and device tables)");
    for (mqc_solver* m : s->members) {
        invalidate_graphs(m);
        set_option_rank(m, key, value);
        if (std::string(key) == "sigma" || std::string(key) == "sigma_ab") {      // Hamiltonian tables are rebuilt on the next evaluation
            m->sig.ready = false;
            m->lay[0].sham_ready = m->lay[1].sham_ready = false;
        }
    }
    MQC_CATCH
}

int mqc_set_hamiltonian(mqc_solver_t* L, int64_t n_terms, const uint64_t* xmask, const uint64_t* zmask,
                        const double* coeff_re, const double* coeff_im) {
    MQC_TRY
    if (!L || n_terms < 0 || (n_terms && (!xmask || !zmask || !coeff_re))) throw MqcError("bad arguments");
    const u64 lim = (1ull << L->n) - 1ull;
    for (int64_t t = 0; t < n_terms; ++t)
        if ((xmask[t] | zmask[t]) & ~lim) throw MqcError("Pauli mask outside the register");
    for (mqc_solver* s : L->members) {
        s->ham.x.assign(xmask, xmask + n_terms);
        s->ham.z.assign(zmask, zmask + n_terms);
        s->ham.c.resize(n_terms);
        for (int64_t t = 0; t < n_terms; ++t) s->ham.c[t] = make_double2(coeff_re[t], coeff_im ? coeff_im[t] : 0.0);
        s->have_ham = true;
        s->lay[0].ham_ready = s->lay[1].ham_ready = false;
        s->lay[0].sham_ready = s->lay[1].sham_ready = false;
        s->sig.ready = false;
        invalidate_graphs(s);
        if (s->creal && !ham_is_real(s->ham)) drop_real(s);
    }
    MQC_CATCH
}

// Same Pauli strings as the last mqc_set_hamiltonian, new coefficients (a chemical-potential step
// or the next DMET iteration of the same fragment): the handle, its ansatz, sweep schedules,
// plans and state buffers stay; only the Hamiltonian tables are rebuilt.
int mqc_update_hamiltonian_coeffs(mqc_solver_t* L, int64_t n_terms, const double* coeff_re, const double* coeff_im) {
    MQC_TRY
    if (!L || !coeff_re) throw MqcError("bad arguments");
    if (!L->have_ham) throw MqcError("mqc_set_hamiltonian has not been called");
    if ((size_t)n_terms != L->ham.x.size()) throw MqcError("mqc_update_hamiltonian_coeffs: term count differs from the stored Pauli strings");
    for (mqc_solver* s : L->members) {
        for (int64_t t = 0; t < n_terms; ++t) s->ham.c[t] = make_double2(coeff_re[t], coeff_im ? coeff_im[t] : 0.0);
        s->lay[0].ham_ready = s->lay[1].ham_ready = false;
        s->lay[0].sham_ready = s->lay[1].sham_ready = false;
        s->sig.ready = false;
        invalidate_graphs(s);
        if (s->creal && !ham_is_real(s->ham)) drop_real(s);
    }
    MQC_CATCH
}


// ---- compact sector storage: plans for 8-byte (real) or 16-byte (complex) amplitudes -----------
static bool ham_is_real(const HamHost& H) {
    for (size_t t = 0; t < H.x.size(); ++t) {
        if (H.c[t].x == 0.0 && H.c[t].y == 0.0) continue;
        if (H.c[t].y != 0.0 || (__builtin_popcountll(H.x[t] & H.z[t]) & 1)) return false;    // odd number of Y: imaginary matrix elements
    }
    return true;
}
static bool want_real(const mqc_solver* s) {
    return s->real_amplitudes != 0 && s->use_sigma && !s->sigma_ab && s->n / 2 <= 18 && (!s->have_ham || ham_is_real(s->ham));
}
static void build_compact_plans(mqc_solver* s, bool real) {
    s->use_compact = false; s->cvalid = false; s->creal = false;
    for (int w = 0; w < 2; ++w) s->cplan[w] = MqcCPlan();
    if (!(s->compact && s->conserving && s->use_sector && s->mode == 1 && s->n_ranks == 1)) return;
    for (int attempt = real ? 0 : 1; attempt < 2; ++attempt) {
        const int elem = attempt == 0 ? 8 : 16;
        bool ok = true;
        for (int w = 0; w < 2 && ok; ++w) {
            s->cplan[w] = mqc_build_cplan(s->sched[w], s->fac, s->na, s->nb, s->ref_index, elem);
            ok = s->cplan[w].ok && s->cplan[w].max_smem[w] <= 232448;
            if (!ok && getenv("MQC_DEBUG")) fprintf(stderr, "[mqc] compact plan %d (%d-byte amplitudes) not used: %s (smem %zu)\n", w, elem, s->cplan[w].why, s->cplan[w].max_smem[w]);
        }
        if (ok) { s->use_compact = true; s->creal = elem == 8; return; }
    }
    for (int w = 0; w < 2; ++w) s->cplan[w] = MqcCPlan();
}
static size_t cbuf_len(const mqc_solver* s, int w) {       // double2 units of one compact buffer
    const size_t cnt = (size_t)s->cplan[w].nA * s->cplan[w].pitch;
    return s->creal ? (cnt + 1) / 2 : cnt;
}
static void upload_compact(mqc_solver* s) {
        for (int w = 0; w < 2; ++w) {
            CPlanDev& D = s->cdev[w];
            D.reset();
            if (!s->use_compact) continue;
            const MqcCPlan& PL = s->cplan[w];
            D.pass.upload(PL.pass.empty() ? std::vector<MqcCPassDev>(1) : PL.pass, s->stream);
            D.cls.upload(PL.cls, s->stream); D.fac.upload(PL.fac, s->stream); D.maps.upload(PL.maps, s->stream);
            // Tiles are listed heaviest first and CTA b starts on SM b mod n_sm: deal them to the SMs in
            // boustrophedon order (row r of n_sm tiles reversed when r is odd) so that every SM gets the same
            // mix of heavy and light tiles (measured before: 23 % of the SM-cycles idle at the end of a pass).
            std::vector<uint32_t> tiles = PL.tiles;
            if (s->csweep_snake) {
                const size_t W = (size_t)std::max(1, s->n_sm);
                for (const MqcCPassDev& P : PL.pass) {
                    if (!P.tiles_explicit) continue;
                    for (size_t r0 = W; r0 < P.n_tiles; r0 += 2 * W) {
                        const size_t r1 = std::min<size_t>(P.n_tiles, r0 + W);
                        std::reverse(tiles.begin() + P.tile_off + r0, tiles.begin() + P.tile_off + r1);
                    }
                }
            }
            D.tiles.upload(tiles, s->stream); D.live_tab.upload(PL.live_tab, s->stream);
            D.prog_hdr.upload(PL.prog_hdr, s->stream); D.prog.upload(PL.prog, s->stream);
        }
        if (s->use_compact) {
            const size_t cnt = cbuf_len(s, 0);
            for (int b = 0; b < 2; ++b) if (s->cpsi[b].n != cnt) s->cpsi[b].alloc(cnt);
            s->cscratch.release();
            for (int b = 0; b < 2; ++b) s->clam[b].release();
            s->psi.release(); s->lam.release();          // the dense register is only allocated on demand (bridge_to_dense)
        } else {
            for (int b = 0; b < 2; ++b) { s->cpsi[b].release(); s->clam[b].release(); }
            if (s->n_ranks == 1) alloc_state(s, false);
        }
}
// real storage no longer applies (complex Hamiltonian or state): same schedules, complex plans and buffers
static void drop_real(mqc_solver* s) {
    if (!(s->use_compact && s->creal)) return;
    invalidate_graphs(s);
    build_compact_plans(s, false);
    if (!s->use_compact) throw MqcError("the compact plan does not fit with complex amplitudes; set option real_amplitudes = 0 and lower tile_bits");
    if (s->gpu_ready) { upload_compact(s); CK(cudaStreamSynchronize(s->stream)); }
    s->state_layout = -1;
}

static void set_ansatz_rank(mqc_solver* s, int32_t n_factors, const int32_t* kind, const int32_t* idx4,
                            const int32_t* theta_index, const double* theta_scale, int32_t n_theta, uint64_t ref_index,
                            const mqc_solver* copy_from) {
    const int n = s->n;
    if (copy_from) {
        s->fac = copy_from->fac; s->conserving = copy_from->conserving;
        s->sched[0] = copy_from->sched[0]; s->sched[1] = copy_from->sched[1];
    } else {
        std::vector<MqcFactor> fac;
        fac.reserve(n_factors);
        bool conserving = (n % 2 == 0);
        for (int k = 0; k < n_factors; ++k) {
            if (theta_index[k] < 0 || theta_index[k] >= n_theta) throw MqcError("theta_index out of range");
            fac.push_back(make_factor(n, kind[k], idx4 + 4 * k, k, k));          // gradients accumulate per factor (k_grad_combine)
            const int32_t* ix = idx4 + 4 * k;
            if (kind[k] == 1 || kind[k] == 3) { if ((ix[0] ^ ix[1]) & 1) conserving = false; }
            else {
                const int sa = (ix[0] & 1) + (ix[1] & 1), sc = (ix[2] & 1) + (ix[3] & 1);
                if (sa != sc) conserving = false;
            }
        }
        s->fac = fac;
        s->conserving = conserving;
        // schedules: 0 = energy-only, 1 = energy + gradient
        // compressed-sector tiles keep only the sector in shared memory and take 14-bit tiles;
        // dense tiles hold 2^K amplitudes (x2 with the adjoint): 13 / 12 bits at most
        const bool sector_tiles = s->conserving && s->use_sector && s->sparse_tiles && s->sp2;
        for (int w = 0; w < 2; ++w) {
            int kb = (w == 0) ? s->tile_bits : s->grad_tile_bits;
            if (!sector_tiles) kb = std::min(kb, 12);
            // dense tiles want 128-byte runs (3 low bits always active); sector tiles touch
            // isolated amplitudes anyway and fuse better without the constraint
            const int lb = s->low_bits >= 0 ? s->low_bits : (sector_tiles ? 0 : 3);
            if (s->mode == 1) {
                s->sched[w] = mqc_build_schedule(n, kb, lb, s->fac, MQC_MAX_PASS_FACTORS, s->n_global);
            } else {
                s->sched[w] = MqcSchedule();
                s->sched[w].n_bits = n;
                s->sched[w].phys_final.resize(n);
                std::iota(s->sched[w].phys_final.begin(), s->sched[w].phys_final.end(), 0);
                s->sched[w].phys_init = s->sched[w].phys_final;
            }
        }
    }
    s->n_theta = n_theta;
    s->fac_theta.assign(theta_index, theta_index + n_factors);
    s->fac_scale.assign(n_factors, 1.0);
    if (theta_scale) s->fac_scale.assign(theta_scale, theta_scale + n_factors);
    s->ref_index = ref_index;
    s->na = s->nb = 0;
    for (int q = 0; q < n; ++q)
        if ((ref_index >> (n - 1 - q)) & 1ull) { if (q & 1) s->nb++; else s->na++; }
    {
        unsigned nA, nB;
        sector_dims_pre(s, nA, nB);
        const double full_c = 16.0 * (double)nA * (double)nB;
        s->blocked = s->n_ranks > 1 && s->conserving && s->use_sector && s->use_sigma &&
                     (s->sigma_blocked == 1 || (s->sigma_blocked < 0 && full_c > 24e9));
        s->connected_cown = false;
        s->cown.release(); s->cblk.release();
    }
    for (int w = 0; w < 2; ++w) {
        if (copy_from) s->plan[w] = copy_from->plan[w];
        else if (s->conserving && s->mode == 1) s->plan[w] = mqc_build_sparse_plan(s->sched[w], s->na, s->nb);
        else s->plan[w] = MqcSparsePlan();
    }
    build_compact_plans(s, want_real(s));
    for (const MqcFactor& f : s->fac)
        if (f.ctrl && !s->use_compact)
            throw MqcError("number-dressed factors (kind 3) need the compact sector storage: an unsharded handle, an ansatz that "
                           "conserves (N_alpha, N_beta), options compact = 1, mode = 1, sector = 1");
    for (int w = 0; w < 2; ++w) {
        s->lay[w].reset();
        s->lay[w].phys = s->sched[w].phys_final;
    }
    s->have_ansatz = true;
    s->state_layout = -1;
    s->sig.ready = false;
    invalidate_graphs(s);
    s->connected = (s->n_ranks == 1);
    // device side is set up lazily so that schedules can be inspected without a GPU
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) == cudaSuccess && cnt > 0) {
        require_gpu(s);
        std::vector<int> ti(n_factors);
        for (int k = 0; k < n_factors; ++k) ti[k] = theta_index[k];
        s->theta_index.upload(ti.empty() ? std::vector<int>(1, 0) : ti, s->stream);
        {
            std::vector<int> ptr(std::max(1, n_theta) + 1, 0), lst(std::max(1, n_factors), 0);
            for (int k = 0; k < n_factors; ++k) ptr[ti[k] + 1]++;
            for (int t = 0; t < n_theta; ++t) ptr[t + 1] += ptr[t];
            std::vector<int> pos(ptr.begin(), ptr.end() - 1);
            for (int k = 0; k < n_factors; ++k) lst[pos[ti[k]]++] = k;
            s->th_ptr.upload(ptr, s->stream);
            s->th_fac.upload(lst, s->stream);
        }
        std::vector<double> sc(std::max(1, n_factors), 1.0);
        if (theta_scale) for (int k = 0; k < n_factors; ++k) sc[k] = theta_scale[k];
        s->theta_scale.upload(sc, s->stream);
        s->theta.alloc(std::max(1, n_theta));
        s->gtheta.alloc(std::max(1, n_theta));
        s->mail.alloc(MQC_MAIL_GRAD + std::max(1, n_factors));
        CK(cudaMemsetAsync(s->mail.p, 0, sizeof(double) * s->mail.n, s->stream));
        s->err.alloc(1);
        CK(cudaMemsetAsync(s->err.p, 0, sizeof(int), s->stream));
        s->epoch = 0;
        s->cs.alloc(std::max(1, n_factors));
        for (int w = 0; w < 2; ++w) {
            if (s->plan[w].ok) {
                s->pl_dep[w].upload(s->plan[w].dep, s->stream);
                s->pl_depp[w].upload(s->plan[w].depp, s->stream);
                s->pl_lists[w].upload(s->plan[w].lists, s->stream);
                s->pl_pat[w].upload(s->plan[w].pat, s->stream);
                std::vector<uint4> hd(s->plan[w].hdr.size() / 4);
                for (size_t i = 0; i < hd.size(); ++i)
                    hd[i] = make_uint4(s->plan[w].hdr[4 * i], s->plan[w].hdr[4 * i + 1], s->plan[w].hdr[4 * i + 2], s->plan[w].hdr[4 * i + 3]);
                s->pl_hdr[w].upload(hd, s->stream);
                // size-sorted tile lists balance the SMs while the sector is cache resident; once it
                // lives in HBM the natural tile order has the better DRAM locality (measured at 30 qubits)
                const double sector_bytes = 16.0 * binom_d((s->n + 1) / 2, s->na) * binom_d(s->n / 2, s->nb);
                s->tlist[w] = sector_bytes <= 64e6 ? mqc_build_tile_lists(s->sched[w], s->plan[w], s->rank) : MqcTileLists();
                if (s->tlist[w].ok) {
                    s->pl_tiles[w].upload(s->tlist[w].tiles, s->stream);
                    CK(cudaStreamSynchronize(s->stream));
                    s->tlist[w].tiles.clear(); s->tlist[w].tiles.shrink_to_fit();
                }
            }
            s->tf[w].upload(s->sched[w].tf, s->stream);
            if (s->mode == 1 && s->sched[w].k <= 13) s->dlut[w].upload(mqc_build_dense_luts(s->sched[w]), s->stream);
            std::vector<int> ph = s->lay[w].phys;
            s->lay[w].phys_dev.upload(ph, s->stream);
        }
        upload_compact(s);
        CK(cudaStreamSynchronize(s->stream));
    } else {
        cudaGetLastError();
    }
}

int mqc_set_ansatz_scaled(mqc_solver_t* L, int32_t n_factors, const int32_t* kind, const int32_t* idx4,
                          const int32_t* theta_index, const double* theta_scale, int32_t n_theta, uint64_t ref_index);

int mqc_set_ansatz(mqc_solver_t* L, int32_t n_factors, const int32_t* kind, const int32_t* idx4,
                   const int32_t* theta_index, int32_t n_theta, uint64_t ref_index) {
    return mqc_set_ansatz_scaled(L, n_factors, kind, idx4, theta_index, nullptr, n_theta, ref_index);
}

int mqc_set_ansatz_scaled(mqc_solver_t* L, int32_t n_factors, const int32_t* kind, const int32_t* idx4,
                          const int32_t* theta_index, const double* theta_scale, int32_t n_theta, uint64_t ref_index) {
    MQC_TRY
    if (!L || n_factors < 0 || n_theta < 0 || (n_factors && (!kind || !idx4 || !theta_index)))
        throw MqcError("bad arguments");
    if (ref_index >> L->n) throw MqcError("reference index outside the register");
    if (L->n_ranks > 1 && L->mode != 1) throw MqcError("a sharded state needs mode = 1 (tile sweep)");
    if (L->comm_mode == 2 && L->connected)
        throw MqcError("this rank's buffers are mapped by its peers (mqc_shard_connect): create a new handle for a new ansatz");
    for (size_t i = 0; i < L->members.size(); ++i)
        set_ansatz_rank(L->members[i], n_factors, kind, idx4, theta_index, theta_scale, n_theta, ref_index, i ? L : nullptr);
    MQC_CATCH
}

// ---- one process per GPU: CUDA IPC exchange of the shard buffers --------------------------
struct BlobEntry { cudaIpcMemHandle_t h; unsigned long long offset; unsigned long long valid; };

typedef int (*cuMemGetAddressRange_t)(unsigned long long*, size_t*, unsigned long long);
static void base_of(void* p, void** base, size_t* off) {
    static cuMemGetAddressRange_t fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &qr));
        if (!f) throw MqcError("cuMemGetAddressRange is not available");
        fn = (cuMemGetAddressRange_t)f;
    }
    unsigned long long b = 0; size_t sz = 0;
    if (fn(&b, &sz, (unsigned long long)(uintptr_t)p) != 0) throw MqcError("cuMemGetAddressRange failed");
    *base = (void*)(uintptr_t)b;
    *off = (size_t)((uintptr_t)p - (uintptr_t)b);
}

int mqc_shard_export(mqc_solver_t* s, int32_t want_lambda, void* blob, int64_t blob_bytes) {
    MQC_TRY
    require_gpu(s);
    if (!blob || blob_bytes < MQC_BLOB_BYTES) throw MqcError("blob must hold MQC_SHARD_BLOB_BYTES bytes");
    if (s->comm_mode != 2) throw MqcError("not a per-process shard handle (mqc_shard_create with n_ranks > 1)");
    if (!s->have_ansatz) throw MqcError("call mqc_set_ansatz before mqc_shard_export");
    alloc_state(s, want_lambda != 0);
    if (s->blocked) alloc_blocked(s);
    CK(cudaStreamSynchronize(s->stream));
    BlobEntry e[4];
    std::memset(e, 0, sizeof(e));
    void* ptrs[4] = {s->psi.p, s->lam.p, s->mail.p, s->blocked ? s->cown.p : nullptr};
    for (int i = 0; i < 4; ++i) {
        if (!ptrs[i]) continue;
        void* base; size_t off;
        base_of(ptrs[i], &base, &off);
        CK(cudaIpcGetMemHandle(&e[i].h, base));
        e[i].offset = off; e[i].valid = 1;
    }
    std::memset(blob, 0, MQC_BLOB_BYTES);
    std::memcpy(blob, e, sizeof(e));
    MQC_CATCH
}

int mqc_shard_connect(mqc_solver_t* s, const void* blobs, int64_t blobs_bytes) {
    MQC_TRY
    require_gpu(s);
    if (s->comm_mode != 2) throw MqcError("not a per-process shard handle");
    if (!blobs || blobs_bytes < (int64_t)MQC_BLOB_BYTES * s->n_ranks) throw MqcError("need one blob per rank");
    for (int r = 0; r < s->n_ranks; ++r) {
        if (r == s->rank) { s->peer_psi[r] = s->psi.p; s->peer_lam[r] = s->lam.p; s->peer_mail[r] = s->mail.p; s->peer_cown[r] = s->cown.p; continue; }
        BlobEntry e[4];
        std::memcpy(e, (const char*)blobs + (size_t)r * MQC_BLOB_BYTES, sizeof(e));
        void* out[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int i = 0; i < 4; ++i) {
            if (!e[i].valid) continue;
            const std::string key((const char*)&e[i].h, sizeof(cudaIpcMemHandle_t));
            void* base = nullptr;
            for (auto& kv : s->ipc_open) if (kv.first == key) base = kv.second;
            if (!base) {
                CK(cudaIpcOpenMemHandle(&base, e[i].h, cudaIpcMemLazyEnablePeerAccess));
                s->ipc_open.emplace_back(key, base);
            }
            out[i] = (char*)base + e[i].offset;
        }
        s->peer_psi[r] = (double2*)out[0]; s->peer_lam[r] = (double2*)out[1]; s->peer_mail[r] = (double*)out[2];
        s->peer_cown[r] = (double2*)out[3];
        if (!out[0] || !out[2] || (s->blocked && !out[3])) throw MqcError("peer blob is incomplete");
    }
    s->connected = true;
    s->connected_cown = s->blocked;
    MQC_CATCH
}

int mqc_shard_info(mqc_solver_t* s, int32_t* rank, int32_t* n_ranks, int32_t* n_local_bits, int32_t* members) {
    MQC_TRY
    if (!s) throw MqcError("null solver handle");
    if (rank) *rank = s->rank;
    if (n_ranks) *n_ranks = s->n_ranks;
    if (n_local_bits) *n_local_bits = s->n_local;
    if (members) *members = (int32_t)s->members.size();
    MQC_CATCH
}

int mqc_prepare(mqc_solver_t* s, const double* theta) {
    MQC_TRY
    require_gpu(s);
    if (!theta) throw MqcError("null theta");
    reset_counters(s);
    connect_group(s, false);
    forward(s, 0, theta);
    sync_all(s);
    MQC_CATCH
}

// The launch sequence of one evaluation (which = 0: energy, 1: energy + adjoint gradient) on
// the handle's stream(s), with the phase events.  theta == NULL: parameters already on the device.
static void evaluation_body(mqc_solver* s, int which, const double* theta) {
    record_ev(s, 0);
    forward(s, which, theta);
    record_ev(s, 1);
    happly(s, which, own_views(s, which), which == 1);
    if (which == 0) reduce_mail(s, 4);
    record_ev(s, 2);
    if (which == 1) {
        reverse_sweep(s, which);
        reduce_mail(s, MQC_MAIL_GRAD + (int)s->fac.size());
        {   // per-factor gradients -> parameters (tied parameters, angle scales)
            CK(cudaSetDevice(s->device));
            CK(cudaMemsetAsync(s->gtheta.p, 0, sizeof(double) * std::max(1, s->n_theta), s->stream));
            const int nfac = (int)s->fac.size();
            if (nfac && s->use_compact && s->deterministic) {
                k_grad_combine_parts<<<(unsigned)(((size_t)std::max(1, s->n_theta) * 32 + 255) / 256), 256, 0, s->stream>>>(
                    s->cgpart.p, s->cgwhere.p, s->th_ptr.p, s->th_fac.p, s->theta_scale.p, s->gtheta.p, s->n_theta);
                s->counters[0]++;
            } else if (nfac) { k_grad_combine<<<(nfac + 255) / 256, 256, 0, s->stream>>>(reduced_grad(s), s->theta_index.p, s->theta_scale.p, s->gtheta.p, nfac); s->counters[0]++; }
        }
        record_ev(s, 3);
    }
}

static void run_evaluation(mqc_solver* s, int which, const double* theta) {
    const bool graphable = s->use_graph && s->n_ranks == 1 && s->mode == 1;
    if (!graphable) { evaluation_body(s, which, theta); return; }
    CK(cudaSetDevice(s->device));
    if (theta) {
        s->theta_host.assign(theta, theta + s->n_theta);
        CK(cudaMemcpyAsync(s->theta.p, s->theta_host.data(), sizeof(double) * s->n_theta, cudaMemcpyHostToDevice, s->stream));
    }
    if (s->gexec[which]) {
        CK(cudaGraphLaunch(s->gexec[which], s->stream));
        for (int i = 0; i < 3; ++i) s->counters[i] = s->gcounters[which][i];
        for (mqc_solver* m : s->members) m->state_layout = which == 1 ? -1 : which;
        if (s->use_compact) {
            s->state_layout = -1;
            s->cvalid = (which == 0); s->cwhich = which;
            s->ccur = (int)(s->cplan[which].pass.size() & 1);
        }
        return;
    }
    if (!s->eager_done[which]) {                  // first call: allocates its work buffers
        evaluation_body(s, which, nullptr);
        s->eager_done[which] = 1;
        return;
    }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    s->capturing = true;
    try {
        evaluation_body(s, which, nullptr);
    } catch (...) {
        s->capturing = false;
        cudaStreamEndCapture(s->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        throw;
    }
    s->capturing = false;
    CK(cudaStreamEndCapture(s->stream, &graph));
    for (int i = 0; i < 3; ++i) s->gcounters[which][i] = s->counters[i];
    const cudaError_t e = cudaGraphInstantiate(&s->gexec[which], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { s->gexec[which] = nullptr; cuda_check(e, "cudaGraphInstantiate"); }
    CK(cudaGraphLaunch(s->gexec[which], s->stream));
}

int mqc_energy(mqc_solver_t* s, const double* theta, double* energy) {
    MQC_TRY
    require_gpu(s);
    if (!theta || !energy) throw MqcError("null argument");
    reset_counters(s);
    connect_group(s, false);
    for (mqc_solver* m : s->members) ensure_layout_ready(m, 0);
    double val[2] = {0, 0};
    run_evaluation(s, 0, theta);
    CK(cudaSetDevice(s->device));
    CK(cudaMemcpyAsync(val, reduced_scalars(s), 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    sync_all(s);
    finish_timing(s, false);
    *energy = val[0];
    MQC_CATCH
}

static void run_energy_grad(mqc_solver* s, const double* theta, double* energy, double* grad) {
    reset_counters(s);
    connect_group(s, true);
    for (mqc_solver* m : s->members) {
        ensure_layout_ready(m, 1);
        if (m->use_compact) {
            const size_t cnt = cbuf_len(m, 1);
            for (int b = 0; b < 2; ++b) if (m->clam[b].n != cnt) { invalidate_graphs(m); m->clam[b].alloc(cnt); }
        } else if (!m->lam.p) {
            if (m->n_ranks > 1) throw MqcError("sharded handle was exported without lambda");
            invalidate_graphs(m); alloc_state(m, true);
        }
    }
    double val[2] = {0, 0};
    run_evaluation(s, 1, theta);
    CK(cudaSetDevice(s->device));
    CK(cudaMemcpyAsync(val, reduced_scalars(s), 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (grad && s->n_theta)
        CK(cudaMemcpyAsync(grad, s->gtheta.p, sizeof(double) * s->n_theta, cudaMemcpyDeviceToHost, s->stream));
    sync_all(s);
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

// the prepared state exists: dense psi in layout state_layout, or (compact storage) cpsi[ccur]
static void ensure_prepared(mqc_solver* s, const double* theta) {
    if (theta) { connect_group(s, false); forward(s, 0, theta); return; }
    if (s->state_layout >= 0 || (s->use_compact && s->cvalid)) return;
    if ((int)s->theta_host.size() != s->n_theta || !s->have_ansatz) throw MqcError("no prepared state: pass theta");
    std::vector<double> th = s->theta_host;
    connect_group(s, false);
    forward(s, 0, th.data());
}
// the compact state as complex amplitudes (canonical order): the buffer itself, or a converted copy of the real one
static const double2* compact_complex_view(mqc_solver* s) {
    if (!s->creal) return s->cpsi[s->ccur].p;
    const size_t cnt = (size_t)s->cplan[s->cwhich].nA * s->cplan[s->cwhich].nB;
    if (s->cscratch.n != cnt) s->cscratch.alloc(cnt);
    k_real_to_complex<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 16), 256, 0, s->stream>>>(
        reinterpret_cast<const double*>(s->cpsi[s->ccur].p), s->cscratch.p, cnt, s->cplan[s->cwhich].nB, s->cplan[s->cwhich].pitch);
    return s->cscratch.p;
}
// compact storage -> dense register in the physical layout of schedule cwhich (for the consumers
// that still address 2^N amplitudes: mqc_state and RDM sectors other than the ansatz's)
static void bridge_to_dense(mqc_solver* s) {
    if (!(s->use_compact && s->cvalid) || s->state_layout >= 0) return;
    const int w = s->cwhich;
    CK(cudaSetDevice(s->device));
    if (!s->psi.p) { invalidate_graphs(s); s->psi.alloc(s->dim); s->peer_psi[0] = s->psi.p; }
    LayoutTables& Lt = s->lay[w];
    build_sector_tables(s, Lt, s->na, s->nb);
    CK(cudaMemsetAsync(s->psi.p, 0, sizeof(double2) * s->dim, s->stream));
    const unsigned nA = s->cplan[w].nA, nB = s->cplan[w].nB;
    const u64 cnt = (u64)nA * nB;
    MqcPeers PR = peers_of(s);
    PR.lam[0] = s->psi.p;               // k_sector_scatter writes through the lambda table: point it at psi
    k_sector_scatter<<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(PR, compact_complex_view(s), Lt.astr.p, Lt.bstr.p, 0u, nA, nB, 0u);
    s->state_layout = w;
}
static void ensure_state(mqc_solver* s, const double* theta) {
    ensure_prepared(s, theta);
    bridge_to_dense(s);
}

int mqc_state(mqc_solver_t* s, const double* theta, double* re_im, int64_t n_amplitudes) {
    MQC_TRY
    require_gpu(s);
    if (s->n_ranks > 1) throw MqcError("mqc_state needs an unsharded handle: read shards with mqc_shard_read");
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

int mqc_shard_read(mqc_solver_t* L, const double* theta, int32_t member, double* re_im, int64_t n_amplitudes,
                   int32_t* phys_of_logical) {
    MQC_TRY
    require_gpu(L);
    if (member < 0 || member >= (int)L->members.size()) throw MqcError("member out of range");
    mqc_solver* s = L->members[member];
    if (!re_im || (u64)n_amplitudes != s->dim) throw MqcError("shard buffer must hold 2^n_local amplitudes");
    ensure_state(L, theta);
    sync_all(L);
    CK(cudaSetDevice(s->device));
    CK(cudaMemcpy(re_im, s->psi.p, sizeof(double2) * s->dim, cudaMemcpyDeviceToHost));
    if (phys_of_logical) for (int l = 0; l < s->n; ++l) phys_of_logical[l] = s->lay[s->state_layout].phys[l];
    CK(cudaSetDevice(L->device));
    MQC_CATCH
}

int mqc_expectation(mqc_solver_t* L, int64_t n_terms, const uint64_t* xmask, const uint64_t* zmask,
                    const double* coeff_re, const double* coeff_im, double* value_re, double* value_im) {
    MQC_TRY
    require_gpu(L);
    if (n_terms < 0 || (n_terms && (!xmask || !zmask || !coeff_re)) || !value_re) throw MqcError("bad arguments");
    ensure_prepared(L, nullptr);
    const int w = (L->use_compact && L->cvalid && L->state_layout < 0) ? L->cwhich : L->state_layout;
    HamHost H;
    H.x.assign(xmask, xmask + n_terms);
    H.z.assign(zmask, zmask + n_terms);
    H.c.resize(n_terms);
    for (int64_t t = 0; t < n_terms; ++t) H.c[t] = make_double2(coeff_re[t], coeff_im ? coeff_im[t] : 0.0);
    const bool sec = L->conserving && L->use_sector;
    const size_t nm = L->members.size();
    std::vector<HamTables> T(nm);
    std::vector<SecHamTables> ST(nm);
    std::vector<SigmaTables> SG(nm);
    std::vector<HamView> views;
    for (size_t i = 0; i < nm; ++i) {
        mqc_solver* s = L->members[i];
        CK(cudaSetDevice(s->device));
        const bool sigma = sec && s->use_sigma && s->n / 2 <= 18;
        if (sec) {
            build_sector_tables(s, s->lay[w], s->na, s->nb);
            HamHost fb;
            if (sigma) build_sigma(s, H, s->na, s->nb, SG[i], fb, s->use_compact && s->sigma_ab);
            build_sec_ham(s, sigma ? fb : H, s->lay[w].phys, ST[i]);
        } else build_ham_tables(s, H, s->lay[w].phys, T[i]);
        views.push_back(HamView{&T[i], sec ? &ST[i] : nullptr, sigma ? &SG[i] : nullptr});
    }
    double val[2] = {0, 0};
    happly(L, w, views, false);
    reduce_mail(L, 4);
    CK(cudaSetDevice(L->device));
    CK(cudaMemcpyAsync(val, reduced_scalars(L), 2 * sizeof(double), cudaMemcpyDeviceToHost, L->stream));
    sync_all(L);
    *value_re = val[0];
    if (value_im) *value_im = val[1];
    MQC_CATCH
}

// RDMs of this handle's ranks (every rank of the group when they share the process; this
// rank's partial sums when each rank is its own process: the host adds them up)
int mqc_rdm12(mqc_solver_t* L, int32_t n_orb, int32_t n_alpha, int32_t n_beta, double* rdm1, double* rdm2) {
    MQC_TRY
    require_gpu(L);
    if (!rdm1) throw MqcError("null rdm1");
    if (2 * n_orb != L->n) throw MqcError("n_orb must be n_qubits / 2");
    ensure_prepared(L, nullptr);
    if (L->use_compact && L->cvalid && L->state_layout < 0 && L->rdm_gram && n_alpha == L->na && n_beta == L->nb && n_orb <= 31) {
        // compact storage: Gram-matrix form, no atomics (mqc_rdm.cuh)
        mqc_solver* s = L;
        CK(cudaSetDevice(s->device));
        LayoutTables& Lt = s->lay[s->cwhich];
        build_sector_tables(s, Lt, n_alpha, n_beta);
        const int n2i = n_orb * n_orb;
        const int T = n2i <= 144 ? 9 : 8, BS = 16 * T;
        const int nblk = (n2i + BS - 1) / BS;
        MqcRdmC R{};
        R.C = s->cpsi[s->ccur].p; R.aorb = Lt.aorb.p; R.borb = Lt.borb.p; R.rank_a = Lt.rank_a.p; R.rank_b = Lt.rank_b.p;
        R.nA = s->cplan[s->cwhich].nA; R.nB = s->cplan[s->cwhich].nB; R.pitch = s->cplan[s->cwhich].pitch; R.n = n_orb; R.n2 = n2i; R.nblk = nblk;
        const u64 n_blocks = (u64)R.nA * ((R.nB + MQC_RDM_KB - 1) / MQC_RDM_KB);
        const int groups = (int)std::min<u64>(n_blocks, (u64)std::max(1, s->n_sm / (nblk * nblk)));
        const size_t gsz = (size_t)groups * nblk * nblk * BS * BS, g1sz = (size_t)groups * n2i;
        if (s->rdm_gpart.n < gsz) s->rdm_gpart.alloc(gsz);
        if (s->rdm_g1part.n < g1sz) s->rdm_g1part.alloc(g1sz);
        const size_t n2s = (size_t)n2i, n4s = n2s * n2s;
        if (s->rdm_out.n < n2s + n4s) s->rdm_out.alloc(n2s + n4s);
        R.gpart = s->rdm_gpart.p; R.g1part = s->rdm_g1part.p;
        CK(cudaMemsetAsync(s->rdm_g1part.p, 0, sizeof(double) * g1sz, s->stream));
        const size_t smem = (size_t)(nblk > 1 ? 2 : 1) * MQC_RDM_KB * BS * (s->creal ? sizeof(double) : sizeof(double2));
        dim3 grid(groups, nblk * nblk);
        if (s->creal) {
            if (T == 9) k_rdm_gram<9, true><<<grid, MQC_RDM_THREADS, smem, s->stream>>>(R);
            else k_rdm_gram<8, true><<<grid, MQC_RDM_THREADS, smem, s->stream>>>(R);
        } else if (T == 9) k_rdm_gram<9, false><<<grid, MQC_RDM_THREADS, smem, s->stream>>>(R);
        else k_rdm_gram<8, false><<<grid, MQC_RDM_THREADS, smem, s->stream>>>(R);
        const size_t work = rdm2 ? n4s : n2s;
        k_rdm_finish<<<(unsigned)((work + 255) / 256), 256, 0, s->stream>>>(s->rdm_gpart.p, s->rdm_g1part.p, groups, n_orb, nblk, BS,
                                                                          s->rdm_out.p, rdm2 ? s->rdm_out.p + n2s : nullptr);
        CK(cudaMemcpyAsync(rdm1, s->rdm_out.p, sizeof(double) * n2s, cudaMemcpyDeviceToHost, s->stream));
        if (rdm2) CK(cudaMemcpyAsync(rdm2, s->rdm_out.p + n2s, sizeof(double) * n4s, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        s->counters[0] += 2;
        return 0;
    }
    bridge_to_dense(L);
    sync_all(L);
    const int w = L->state_layout;
    const size_t n2 = (size_t)n_orb * n_orb, n4 = n2 * n2;
    const int copies = 8;
    std::vector<double> acc1(n2, 0.0), acc2(rdm2 ? n4 : 0, 0.0), tmp(rdm2 ? n4 : n2);
    for (mqc_solver* s : L->members) {
        CK(cudaSetDevice(s->device));
        LayoutTables& Lt = s->lay[w];
        build_sector_tables(s, Lt, n_alpha, n_beta);
        MqcSectorDev S = sector_dev(Lt, true);
        const u64 count = (u64)Lt.astr.n * (u64)Lt.bstr.n;
        const u64 d0 = count * s->rank / s->n_ranks, d1 = count * (s->rank + 1) / s->n_ranks;
        MqcRdmDev R{n_orb, Lt.qbit.p, Lt.below.p, copies};
        const size_t need = (size_t)copies * (rdm2 ? n4 : n2);
        if (s->rdm_acc.n < need) s->rdm_acc.alloc(need);
        const MqcPeers PR = peers_of(s);
        CK(cudaMemsetAsync(s->rdm_acc.p, 0, sizeof(double) * (size_t)copies * n2, s->stream));
        if (d1 > d0) k_rdm1<<<(unsigned)((d1 - d0 + 127) / 128), 128, 0, s->stream>>>(PR, S, R, d0, d1, s->rdm_acc.p);
        k_reduce_copies<<<(unsigned)((n2 + 255) / 256), 256, 0, s->stream>>>(s->rdm_acc.p, n2, copies);
        CK(cudaMemcpyAsync(tmp.data(), s->rdm_acc.p, sizeof(double) * n2, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        for (size_t i = 0; i < n2; ++i) acc1[i] += tmp[i];
        if (rdm2) {
            CK(cudaMemsetAsync(s->rdm_acc.p, 0, sizeof(double) * (size_t)copies * n4, s->stream));
            const u64 work = (d1 - d0) * (u64)(2 * n_orb);
            if (work) k_rdm2<<<(unsigned)((work + 127) / 128), 128, 0, s->stream>>>(PR, S, R, d0, d1, s->rdm_acc.p);
            k_reduce_copies<<<(unsigned)((n4 + 255) / 256), 256, 0, s->stream>>>(s->rdm_acc.p, n4, copies);
            CK(cudaMemcpyAsync(tmp.data(), s->rdm_acc.p, sizeof(double) * n4, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaStreamSynchronize(s->stream));
            for (size_t i = 0; i < n4; ++i) acc2[i] += tmp[i];
        }
        // the sector tables of this layout may now be for a different (na, nb) than the ansatz's
        if (s->conserving && s->use_sector && (n_alpha != s->na || n_beta != s->nb)) build_sector_tables(s, Lt, s->na, s->nb);
    }
    std::memcpy(rdm1, acc1.data(), sizeof(double) * n2);
    if (rdm2) std::memcpy(rdm2, acc2.data(), sizeof(double) * n4);
    group_sync(L);          // no rank re-prepares its shard while a peer still reads it
    sync_all(L);
    MQC_CATCH
}

int mqc_last_timing(mqc_solver_t* s, double* ms4, int64_t* counters3) {
    MQC_TRY
    if (!s) throw MqcError("null solver handle");
    if (ms4) for (int i = 0; i < 4; ++i) ms4[i] = s->ms[i];
    if (counters3) for (int i = 0; i < 3; ++i) counters3[i] = s->counters[i];
    MQC_CATCH
}

int mqc_schedule_info(mqc_solver_t* s, int32_t which /*0 energy, 1 gradient*/, int32_t* n_passes, int32_t* tile_bits, int32_t* n_tile_factors) {
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

int mqc_schedule_tile_map(mqc_solver_t* s, int32_t which, int32_t pass, int32_t rank, uint64_t* tau_hi,
                          uint64_t* fixed_bits, uint64_t* n_tiles) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz) throw MqcError("no schedule");
    const MqcSchedule& S = s->sched[which];
    if (pass < 0 || pass >= (int)S.passes.size()) throw MqcError("pass out of range");
    if (rank < 0 || rank >= s->n_ranks) throw MqcError("rank out of range");
    const MqcTileMap M = mqc_tile_map(S.passes[pass], s->n, s->n_global, rank);
    if (tau_hi) *tau_hi = M.tau_hi;
    if (fixed_bits) *fixed_bits = M.fixed_bits;
    if (n_tiles) *n_tiles = M.n_tiles;
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

// Host-only: run the compressed-sector tile plan of schedule `which` on the CPU (one thread)
// and return the state in OpenFermion order.  Tests of the plan tables; small registers only.
int mqc_sparse_plan_host_state(mqc_solver_t* s, int32_t which, const double* theta, double* re_im, int64_t n_amplitudes) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz || !theta || !re_im) throw MqcError("bad arguments");
    if (s->n > 22 || s->n_ranks > 1 || s->mode != 1) throw MqcError("host plan execution: unsharded tile schedules up to 22 qubits");
    if (!s->conserving) throw MqcError("the ansatz does not conserve (N_alpha, N_beta)");
    if ((u64)n_amplitudes != (1ull << s->n)) throw MqcError("state buffer must hold 2^N amplitudes");
    const MqcSchedule& S = s->sched[which];
    MqcSparsePlan PL = mqc_build_sparse_plan(S, s->na, s->nb);
    if (!PL.ok) throw MqcError("no sparse plan for this schedule");
    const size_t dim = (size_t)1 << s->n;
    std::vector<double> re(dim, 0.0), im(dim, 0.0), cv(s->fac.size()), sv(s->fac.size());
    for (size_t k = 0; k < s->fac.size(); ++k) { const double ang = s->fac_scale[k] * theta[s->fac_theta[k]]; cv[k] = std::cos(ang); sv[k] = std::sin(ang); }
    re[mqc_map_mask(s->ref_index, S.phys_init)] = 1.0;
    mqc_sparse_plan_forward_host(S, PL, re, im, cv, sv);
    for (size_t i = 0; i < dim; ++i) {
        const u64 p = mqc_map_mask((u64)i, S.phys_final);
        re_im[2 * i] = re[p]; re_im[2 * i + 1] = im[p];
    }
    MQC_CATCH
}

// Host-only: build the link tables of an operator for the (n_alpha, n_beta) sector and apply
// them to a sector matrix C[n_a][n_b] on ONE CPU thread (tests of the table construction; the
// kernels k_sigma<> execute exactly these tables).  Terms the tables do not cover are counted
// in *n_fallback and left out of L.
int mqc_sigma_host_apply(int32_t n_qubits, int32_t n_alpha, int32_t n_beta, int64_t n_terms, const uint64_t* xmask,
                         const uint64_t* zmask, const double* coeff_re, const double* coeff_im, const double* c_re_im,
                         double* l_re_im, int64_t* n_rows, int64_t* n_cols, int64_t* n_fallback) {
    MQC_TRY
    if (n_qubits < 2 || n_qubits > 36 || (n_qubits & 1)) throw MqcError("n_qubits must be even, <= 36");
    if (n_terms < 0 || (n_terms && (!xmask || !zmask || !coeff_re))) throw MqcError("bad arguments");
    std::vector<u64> xq(n_terms), zq(n_terms);
    std::vector<double2> cf(n_terms);
    for (int64_t t = 0; t < n_terms; ++t) {
        xq[t] = logical_to_qubit(xmask[t], n_qubits); zq[t] = logical_to_qubit(zmask[t], n_qubits);
        cf[t] = make_double2(coeff_re[t], coeff_im ? coeff_im[t] : 0.0);
    }
    // same split as on the device: opposite-spin doubles -> mqc_sigma_ab.h, the rest -> link tables
    std::vector<unsigned> astr, bstr;
    { std::vector<int> ra, rb; mqc_sg_strings(n_qubits / 2, n_alpha, astr, ra); mqc_sg_strings(n_qubits / 2, n_beta, bstr, rb); }
    MqcSigmaAB AB = mqc_sigma_ab_build(n_qubits / 2, n_alpha, n_beta, xq, zq, cf, astr, bstr);
    if (AB.ok) {
        std::vector<char> used(xq.size(), 0);
        for (size_t t : AB.used_terms) used[t] = 1;
        std::vector<u64> x2, z2; std::vector<double2> c2;
        for (size_t t = 0; t < xq.size(); ++t) if (!used[t]) { x2.push_back(xq[t]); z2.push_back(zq[t]); c2.push_back(cf[t]); }
        xq.swap(x2); zq.swap(z2); cf.swap(c2);
    }
    MqcSigmaHost S = mqc_sigma_build(n_qubits / 2, n_alpha, n_beta, xq, zq, cf);
    if (n_rows) *n_rows = S.n_a;
    if (n_cols) *n_cols = S.n_b;
    if (n_fallback) *n_fallback = (int64_t)S.fallback_terms.size();
    if (c_re_im && l_re_im) {
        mqc_sigma_apply_host(S, reinterpret_cast<const double2*>(c_re_im), reinterpret_cast<double2*>(l_re_im));
        if (AB.ok) mqc_sigma_ab_apply_host(AB, astr, reinterpret_cast<const double2*>(c_re_im), reinterpret_cast<double2*>(l_re_im));
    }
    MQC_CATCH
}

// Host-only: run the compact-sector plan of schedule `which` on the CPU (one thread) and return the
// state in OpenFermion order (tests of the plan tables that k_csweep walks; small registers only).
int mqc_cplan_host_state(mqc_solver_t* s, int32_t which, const double* theta, double* re_im, int64_t n_amplitudes) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz || !theta || !re_im) throw MqcError("bad arguments");
    if (s->n > 24) throw MqcError("host plan execution: up to 24 qubits");
    if ((u64)n_amplitudes != (1ull << s->n)) throw MqcError("state buffer must hold 2^N amplitudes");
    if (!s->use_compact || !s->cplan[which].ok) throw MqcError("no compact-sector plan (the ansatz must conserve N_alpha and N_beta; option compact=1)");
    const MqcCPlan& PL = s->cplan[which];
    std::vector<double> cv(s->fac.size()), sv(s->fac.size());
    for (size_t k = 0; k < s->fac.size(); ++k) { const double ang = s->fac_scale[k] * theta[s->fac_theta[k]]; cv[k] = std::cos(ang); sv[k] = std::sin(ang); }
    const std::vector<double> C = mqc_cplan_forward_host(PL, cv, sv);
    std::memset(re_im, 0, sizeof(double) * 2 * (size_t)n_amplitudes);
    const int n = s->n;
    for (uint32_t ia = 0; ia < PL.nA; ++ia) for (uint32_t ib = 0; ib < PL.nB; ++ib) {
        u64 x = 0;
        for (int p = 0; p < PL.n_orb; ++p) {
            if ((PL.astr[ia] >> p) & 1u) x |= 1ull << (n - 1 - 2 * p);
            if ((PL.bstr[ib] >> p) & 1u) x |= 1ull << (n - 2 - 2 * p);
        }
        re_im[2 * x] = C[2 * ((size_t)ia * PL.nB + ib)]; re_im[2 * x + 1] = C[2 * ((size_t)ia * PL.nB + ib) + 1];
    }
    MQC_CATCH
}

// 8: the compact sector matrix holds real amplitudes, 16: complex, 0: no compact storage (dense register)
int mqc_amplitude_bytes(mqc_solver_t* s, int32_t* bytes) {
    MQC_TRY
    if (!s || !bytes) throw MqcError("bad arguments");
    *bytes = s->use_compact ? (s->creal ? 8 : 16) : 0;
    MQC_CATCH
}

int mqc_cplan_info(mqc_solver_t* s, int32_t which, int32_t* n_passes, int64_t* n_rows, int64_t* n_cols, int64_t* smem_forward,
                   int64_t* smem_adjoint) {
    MQC_TRY
    if (!s || which < 0 || which > 1 || !s->have_ansatz) throw MqcError("no schedule");
    const MqcCPlan& PL = s->cplan[which];
    if (n_passes) *n_passes = (s->use_compact && PL.ok) ? (int32_t)PL.pass.size() : -1;
    if (n_rows) *n_rows = PL.nA;
    if (n_cols) *n_cols = PL.nB;
    if (smem_forward) *smem_forward = (int64_t)PL.max_smem[0];
    if (smem_adjoint) *smem_adjoint = (int64_t)PL.max_smem[1];
    MQC_CATCH
}

// Replaces: nothing in the reference computes on a caller-supplied wave function except the RDM
// routines themselves (mqc/tools/rdm.py:131-172 take `fcivec`): this loads a state vector in
// OpenFermion order so that mqc_rdm12 / mqc_expectation run on it (golden-vector tests).
int mqc_load_state(mqc_solver_t* s, const double* re_im, int64_t n_amplitudes) {
    MQC_TRY
    require_gpu(s);
    if (!re_im || (u64)n_amplitudes != (1ull << s->n)) throw MqcError("state buffer must hold 2^N amplitudes");
    if (s->n_ranks > 1) throw MqcError("mqc_load_state needs an unsharded handle");
    if (!s->have_ansatz) throw MqcError("call mqc_set_ansatz first (an empty factor list fixes the sector through ref_index)");
    if (s->n > 30) throw MqcError("mqc_load_state: up to 30 qubits");
    CK(cudaSetDevice(s->device));
    const int n = s->n;
    if (s->use_compact) {
        const MqcCPlan& PL = s->cplan[0];
        std::vector<double2> C((size_t)PL.nA * PL.nB);
        double norm_in = 0.0, norm_sec = 0.0;
        for (u64 i = 0; i < (u64)n_amplitudes; ++i) norm_in += re_im[2 * i] * re_im[2 * i] + re_im[2 * i + 1] * re_im[2 * i + 1];
        for (uint32_t ia = 0; ia < PL.nA; ++ia) for (uint32_t ib = 0; ib < PL.nB; ++ib) {
            u64 x = 0;
            for (int p = 0; p < PL.n_orb; ++p) {
                if ((PL.astr[ia] >> p) & 1u) x |= 1ull << (n - 1 - 2 * p);
                if ((PL.bstr[ib] >> p) & 1u) x |= 1ull << (n - 2 - 2 * p);
            }
            const double2 v = make_double2(re_im[2 * x], re_im[2 * x + 1]);
            C[(size_t)ia * PL.nB + ib] = v;
            norm_sec += v.x * v.x + v.y * v.y;
        }
        if (norm_in - norm_sec > 1e-12 * std::max(1.0, norm_in))
            throw MqcError("mqc_load_state: the vector has weight outside the (N_alpha, N_beta) sector of the handle's reference determinant");
        s->ccur = 0;
        if (s->creal) {
            bool real = true;
            for (const double2& v : C) if (v.y != 0.0) { real = false; break; }
            if (!real) drop_real(s);
        }
        if (s->creal) {
            const MqcCPlan& PR = s->cplan[0];                     // (drop_real may have rebuilt the plans: re-read)
            std::vector<double> R((size_t)PR.nA * PR.pitch, 0.0);
            for (uint32_t ia = 0; ia < PR.nA; ++ia) for (uint32_t ib = 0; ib < PR.nB; ++ib) R[(size_t)ia * PR.pitch + ib] = C[(size_t)ia * PR.nB + ib].x;
            CK(cudaMemcpyAsync(s->cpsi[0].p, R.data(), sizeof(double) * R.size(), cudaMemcpyHostToDevice, s->stream));
            CK(cudaStreamSynchronize(s->stream));
        } else {
            CK(cudaMemcpyAsync(s->cpsi[0].p, C.data(), sizeof(double2) * C.size(), cudaMemcpyHostToDevice, s->stream));
            CK(cudaStreamSynchronize(s->stream));
        }
        s->cvalid = true; s->cwhich = 0; s->state_layout = -1;
    } else {
        const std::vector<int>& ph = s->lay[0].phys;
        std::vector<double2> P((size_t)n_amplitudes);
        for (u64 i = 0; i < (u64)n_amplitudes; ++i) P[mqc_map_mask(i, ph)] = make_double2(re_im[2 * i], re_im[2 * i + 1]);
        if (!s->psi.p) alloc_state(s, false);
        CK(cudaMemcpyAsync(s->psi.p, P.data(), sizeof(double2) * P.size(), cudaMemcpyHostToDevice, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        s->state_layout = 0;
    }
    s->theta_host.clear();              // the loaded state is not a function of the parameters
    MQC_CATCH
}

// Host-only: how many Pauli terms the opposite-spin double-excitation tables (mqc_sigma_ab.h) take over
// from the link tables, and how many non-zero weights W[ij][kl] they hold (tests).
int mqc_sigma_ab_info(int32_t n_qubits, int32_t n_alpha, int32_t n_beta, int64_t n_terms, const uint64_t* xmask,
                      const uint64_t* zmask, const double* coeff_re, const double* coeff_im, int64_t* n_ab_terms,
                      int64_t* n_weights) {
    MQC_TRY
    if (n_qubits < 2 || n_qubits > 36 || (n_qubits & 1)) throw MqcError("n_qubits must be even, <= 36");
    if (n_terms < 0 || (n_terms && (!xmask || !zmask || !coeff_re))) throw MqcError("bad arguments");
    std::vector<u64> xq(n_terms), zq(n_terms);
    std::vector<double2> cf(n_terms);
    for (int64_t t = 0; t < n_terms; ++t) {
        xq[t] = logical_to_qubit(xmask[t], n_qubits); zq[t] = logical_to_qubit(zmask[t], n_qubits);
        cf[t] = make_double2(coeff_re[t], coeff_im ? coeff_im[t] : 0.0);
    }
    std::vector<unsigned> astr, bstr;
    { std::vector<int> ra, rb; mqc_sg_strings(n_qubits / 2, n_alpha, astr, ra); mqc_sg_strings(n_qubits / 2, n_beta, bstr, rb); }
    MqcSigmaAB AB = mqc_sigma_ab_build(n_qubits / 2, n_alpha, n_beta, xq, zq, cf, astr, bstr);
    if (n_ab_terms) *n_ab_terms = AB.ok ? (int64_t)AB.used_terms.size() : 0;
    int64_t nw = 0;
    if (AB.ok) for (double w : AB.W) if (w != 0.0) ++nw;
    if (n_weights) *n_weights = nw;
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
