/* mqc_b200.h — C ABI of the B200-native UCCSD state-vector fragment solver.
 *
 * This is the drop-in boundary for the hot path of mahuan-git/MultiscaleQuantumComputing's
 * VQE fragment solver.  The reference is pure Python and reaches the arithmetic through
 * VQEChem / OpenFermion / SciPy objects; each entry point below names the reference
 * interface it stands in for (paths relative to the reference checkout).  A Python host
 * binds these with ctypes (multiscalequantumcomputing_b200/_lib.py).
 *
 * Conventions
 *   - state vector: complex128[2^N], amplitude index = OpenFermion's
 *     (qubit q <-> index bit N-1-q; test/testlog_vqe:56-94).
 *   - Pauli terms: masks over INDEX bits, operator X^xmask Z^zmask (X left of Z per qubit):
 *       <i ^ xmask| P |i> = (-1)^popcount(i & zmask).
 *     OpenFermion's label has Y where xmask&zmask and label coefficient
 *     coeff * (-i)^popcount(xmask & zmask).
 *   - ansatz factor k: exp(theta[theta_index[k]] * (T_k - T_k^+)), factor 0 acts first;
 *     kind 1: T = a_a^+ a_i        idx = {i, a, -1, -1}
 *     kind 2: T = a_a^+ a_b^+ a_j a_i   idx = {i, j, a, b}      (spin orbitals = qubits)
 *   - every function returns 0 on success, non-zero on failure; mqc_last_error() then holds
 *     a message (the reference has no exception contract: print+exit(),
 *     mqc/solver/dmet_solver_vqechem.py:59-61).  There is NO CPU fallback: without a CUDA
 *     device every compute entry point fails.
 *   - a handle is bound to one GPU and one stream; handles are not thread safe; calls
 *     return after results are in host memory.
 */
#ifndef MQC_B200_H
#define MQC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mqc_solver mqc_solver_t;

const char* mqc_last_error(void);
int mqc_version(void);
int mqc_device_count(int* count);

/* Replaces: construction of the VQEChem problem object for one fragment,
 * `DMET(...)` mqc/solver/dmet_solver_vqechem.py:43-51 (state of 2^n_qubits amplitudes). */
int mqc_create(int n_qubits, int device, mqc_solver_t** out);
int mqc_destroy(mqc_solver_t* s);

/* Sharded state (BASELINE.json configs[4]; SURVEY.md §8e "sharded state"): the 2^n_qubits
 * amplitudes are split over n_ranks = 2, 4 or 8 GPUs; the top log2(n_ranks) PHYSICAL index
 * bits are the rank.  Every other entry point of this header then works on the sharded
 * state and is COLLECTIVE: all ranks call it with the same arguments.  There is no
 * counterpart in the reference (its scipy.sparse vectors live in one host address space,
 * mqc/solver/dmet_interface.py:114-123); this replaces "does not fit" beyond one GPU.
 *
 *   mqc_shard_group_create  all ranks are driven by THIS process through one handle
 *                           (devices[r] may repeat: several ranks on one GPU, for tests).
 *   mqc_shard_create        one process per GPU (torchrun): after mqc_set_ansatz every rank
 *                           calls mqc_shard_export, the MQC_SHARD_BLOB_BYTES blobs are
 *                           all-gathered by the host (torch.distributed), and
 *                           mqc_shard_connect maps the peers' shards with CUDA IPC.
 * Tile passes whose active bits include rank bits read and write the peers' HBM directly
 * over NVLink (the global-qubit exchange is fused into the compute pass); ranks are ordered
 * by CUDA events (same process) or a system-scope flag barrier kernel (IPC).
 * mqc_rdm12 on a per-process shard handle returns that rank's partial sums (add them up
 * with an all-reduce); everything else returns the reduced result on every rank. */
#define MQC_SHARD_BLOB_BYTES 512
int mqc_shard_group_create(int n_qubits, int n_ranks, const int* devices, mqc_solver_t** out);
int mqc_shard_create(int n_qubits, int device, int rank, int n_ranks, mqc_solver_t** out);
int mqc_shard_export(mqc_solver_t* s, int32_t want_lambda, void* blob, int64_t blob_bytes);
int mqc_shard_connect(mqc_solver_t* s, const void* blobs, int64_t blobs_bytes);
int mqc_shard_info(mqc_solver_t* s, int32_t* rank, int32_t* n_ranks, int32_t* n_local_bits, int32_t* members);
/* Raw PHYSICAL-order amplitudes of one rank's shard (2^n_local complex128) and the layout
 * phys_of_logical[l] (n_qubits entries) they are stored in; `member` selects the rank of a
 * same-process group (0 otherwise).  Test / checkpoint access; replaces nothing. */
int mqc_shard_read(mqc_solver_t* s, const double* theta, int32_t member, double* re_im,
                   int64_t n_amplitudes, int32_t* phys_of_logical);

/* Options (before mqc_set_ansatz): "tile_bits" (default 14; dense tiles are clamped to 13),
 * "low_bits" (-1 = automatic: 3 for dense tiles, 0 for sector tiles), "grad_tile_bits" (14; dense: 12), "mode" (1 = fused tile sweep, 0 = one
 * factor per pass), "sp2" (1 = warp-per-tile compressed-sector sweep k_tile_sp2, 0 = first
 * generation), "sp2_warps" (4), "sp2_list" (1 = host-built lists of the non-empty tiles),
 * "sp2_smem_kb" (220), "sp2_team" (0 = automatic warps per tile), "sigma_stage" (1 = stage source
 * rows of H|psi> in shared memory), "sigma_blocked" (-1 = automatic, 1 = row-distributed sector
 * matrix on sharded states), "graph" (1 = replay evaluations as CUDA graphs on single-rank handles),
 * "sector" (1 = exploit particle-number/S_z conservation of the ansatz in H|psi>),
 * "sparse_tiles" (1 = compressed-sector tiles when the ansatz conserves N_alpha, N_beta),
 * "tile_threads" (256), "sparse_threads" (128), "sigma" (1 = link-table kernel for the sector
 * form of H|psi>, 0 = first-generation X-group kernel).
 * Compact sector storage (round 2): "compact" (1 = the state is the matrix C[alpha string][beta string] whenever the
 * ansatz conserves N_alpha, N_beta and the handle is unsharded), "real_amplitudes" (1 = 8-byte amplitudes when the
 * Hamiltonian is real, see mqc_amplitude_bytes; 0 = always complex128), "csweep_threads" (128: threads per tile of
 * the forward sweep), "csweep_adj_threads" (0 = automatic: 64 with real amplitudes), "csweep_ctas" (0 = one CTA per
 * tile up to a few thousand, else persistent CTAs), "csweep_snake" (1 = boustrophedon tile order), "csweep_pdl"
 * (0; 1 = programmatic dependent launch of the passes), "rdm_gram" (1 = Gram-matrix RDM kernel), "sigma_split"
 * (-1 = automatic: a row of H|psi> is split over two CTAs when there are few rows), "sigma_stage_bytes" (98304),
 * "sigma_ab" (0), "deterministic" (1 = gradient and energy partial sums are added in a fixed order: the same input
 * gives the same bits; 0 = floating-point atomics).  Options that shape schedules, plans or uploaded tables
 * ("tile_bits", "grad_tile_bits", "low_bits", "mode", "sector", "sparse_tiles", "sp2", "compact", "real_amplitudes",
 * "sigma_blocked", "csweep_snake") are refused after mqc_set_ansatz. */
int mqc_set_option(mqc_solver_t* s, const char* key, int64_t value);

/* Replaces: `mapping(fermi_ham,'JW',...)` + `get_sparse_operator(qubit_ham)`,
 * mqc/solver/dmet_interface.py:116-123 (the 2^N x 2^N sparse matrix is never formed). */
int mqc_set_hamiltonian(mqc_solver_t* s, int64_t n_terms, const uint64_t* xmask,
                        const uint64_t* zmask, const double* coeff_re, const double* coeff_im);

/* Replaces: the re-construction of `DMET(...)` for the SAME fragment at the next chemical
 * potential / DMET iteration (mqc/third_party/qcdmet.py:244-260 calls the solver again and the
 * reference rebuilds everything, mqc/solver/dmet_interface.py:45-128).  Same Pauli strings, in
 * the same order, as the last mqc_set_hamiltonian; only the coefficients change.  The handle
 * keeps its ansatz, sweep schedules, plans and device buffers. */
int mqc_update_hamiltonian_coeffs(mqc_solver_t* s, int64_t n_terms, const double* coeff_re,
                                  const double* coeff_im);

/* Replaces: `reference(imp, options)` + `ADAPT(options['ansatz'], imp, pool, ref)`,
 * mqc/solver/dmet_solver_vqechem.py:63-65, for a fixed factor list. */
int mqc_set_ansatz(mqc_solver_t* s, int32_t n_factors, const int32_t* kind, const int32_t* idx4,
                   const int32_t* theta_index, int32_t n_theta, uint64_t ref_index);

/* Same, with an angle scale per factor: factor k rotates by theta_scale[k] * theta[theta_index[k]] and
 * dE/dtheta[t] = sum_{k: theta_index[k] = t} theta_scale[k] * dE/d(angle_k).  This is how a generator that is a
 * linear combination of excitations, A = sum_m c_m (T_m - T_m^+) (the reference's spin-adapted fermionic pool,
 * mqc/solver/dmet_solver_vqechem.py:35-41 'spin_sym': 'sa'), enters the factorised ansatz: its terms are
 * consecutive factors with one shared parameter and theta_scale = c_m (first-order product form).  NULL = all 1.
 * kind 3 (compact sector storage only): number-dressed single n_c (a_a^+ a_i - h.c.), idx = {i, a, c, -1}: the
 * rotation of kind 1 restricted to determinants with spin orbital c occupied (what normal ordering leaves of a
 * two-body generator whose creation and annihilation index sets overlap). */
int mqc_set_ansatz_scaled(mqc_solver_t* s, int32_t n_factors, const int32_t* kind, const int32_t* idx4,
                          const int32_t* theta_index, const double* theta_scale, int32_t n_theta,
                          uint64_t ref_index);

/* Replaces: `ansatz.energy(params)`, called by scipy.optimize.minimize at
 * mqc/solver/dmet_interface.py:295. */
int mqc_energy(mqc_solver_t* s, const double* theta, double* energy);

/* Replaces: `ansatz.energy` + `ansatz.gradient` (jac=), mqc/solver/dmet_interface.py:295.
 * Adjoint method: forward sweep, lambda = H psi, reverse sweep over (psi, lambda). */
int mqc_energy_grad(mqc_solver_t* s, const double* theta, double* energy, double* grad);

/* Same evaluation with the parameters of the previous mqc_energy_grad call, which are still
 * resident in HBM; the gradient stays on the device (benchmarks: no per-step host traffic). */
int mqc_energy_grad_resident(mqc_solver_t* s, double* energy);

/* CUDA events on the solver's stream, for device-side timing of a run of calls:
 * mqc_event_record(s, 0) ... calls ... mqc_event_record(s, 1); mqc_event_elapsed_ms(s, &ms). */
int mqc_event_record(mqc_solver_t* s, int32_t which);
int mqc_event_elapsed_ms(mqc_solver_t* s, double* ms);

/* Replaces: `ansatz.state(params)`, mqc/solver/dmet_solver_vqechem.py:71.  Writes
 * 2^N interleaved (re, im) pairs in OpenFermion index order.  theta == NULL: state of the
 * last mqc_energy / mqc_prepare call. */
int mqc_prepare(mqc_solver_t* s, const double* theta);
int mqc_state(mqc_solver_t* s, const double* theta, double* re_im, int64_t n_amplitudes);

/* <psi| sum_t c_t P_t |psi> for another operator on the prepared state (e.g. S^2,
 * mqc/solver/dmet_interface.py:248, or the impurity-energy operator of
 * mqc/solver/dmet_solver_vqechem.py:75-81). */
int mqc_expectation(mqc_solver_t* s, int64_t n_terms, const uint64_t* xmask, const uint64_t* zmask,
                    const double* coeff_re, const double* coeff_im, double* value_re, double* value_im);

/* Replaces: `get_rdm12(wfn, n_elec, n_orb, Sz)` = make_strings_qubit + qubit_inverse gather +
 * make_rdm1 / make_rdm12 + reorder_rdm, mqc/solver/dmet_solver_vqechem.py:109-116 and
 * mqc/tools/rdm.py:18-30,131-172, in float64 (the reference's float32 is a defect).
 *   rdm1[p*n+q]              = sum_s <q_s^+ p_s>
 *   rdm2[((p*n+q)*n+r)*n+s]  = sum_{s,t} <p_s^+ r_t^+ s_t q_s>      (pyscf order)
 * on the prepared state, restricted to the (n_alpha, n_beta) sector.  rdm2 may be NULL. */
int mqc_rdm12(mqc_solver_t* s, int32_t n_orb, int32_t n_alpha, int32_t n_beta,
              double* rdm1, double* rdm2);

/* Host-only check of the excitation link tables behind the sector form of H|psi>
 * (csrc/mqc_sigma.h): builds them for the (n_alpha, n_beta) sector of an operator and applies
 * them on one CPU thread to a sector matrix C[n_rows][n_cols] (complex128, alpha string major,
 * strings in ascending integer order).  Terms the tables cannot express are counted in
 * *n_fallback and left out.  Test infrastructure for the table construction; the CUDA kernel
 * executes the same tables.  c_re_im / l_re_im may be NULL to query the sizes. */
int mqc_sigma_host_apply(int32_t n_qubits, int32_t n_alpha, int32_t n_beta, int64_t n_terms,
                         const uint64_t* xmask, const uint64_t* zmask, const double* coeff_re,
                         const double* coeff_im, const double* c_re_im, double* l_re_im,
                         int64_t* n_rows, int64_t* n_cols, int64_t* n_fallback);

/* Host-only: the opposite-spin double excitations a^+_{i alpha} a_{j alpha} a^+_{k beta} a_{l beta} are applied by
 * their own kernel in a sign-adjusted string basis (csrc/mqc_sigma_ab.h); their weights W[ij][kl] are read off
 * the Pauli-term table and verified term group by term group.  Reports how many Pauli terms that kernel
 * takes over and how many non-zero weights it holds; mqc_sigma_host_apply evaluates both parts. */
int mqc_sigma_ab_info(int32_t n_qubits, int32_t n_alpha, int32_t n_beta, int64_t n_terms, const uint64_t* xmask,
                      const uint64_t* zmask, const double* coeff_re, const double* coeff_im, int64_t* n_ab_terms,
                      int64_t* n_weights);

/* Host-only check of the compressed-sector tile plan (pattern tables and per-factor pair lists
 * that k_tile_sp2 walks): executes schedule `which` on one CPU thread and returns the state in
 * OpenFermion order.  Unsharded handles up to 22 qubits. */
int mqc_sparse_plan_host_state(mqc_solver_t* s, int32_t which, const double* theta, double* re_im,
                               int64_t n_amplitudes);

/* Host-only check of the COMPACT-sector plan (mqc_cplan.h: row / column layouts per pass, tile
 * classes, pair-list blobs, row / column maps - the tables k_csweep walks): executes schedule
 * `which` on one CPU thread and returns the state in OpenFermion order (<= 24 qubits).
 * mqc_cplan_info: number of passes (-1: the handle does not use compact storage), the sector
 * matrix shape and the shared memory one CTA needs (forward / adjoint). */
int mqc_cplan_host_state(mqc_solver_t* s, int32_t which, const double* theta, double* re_im,
                         int64_t n_amplitudes);
int mqc_cplan_info(mqc_solver_t* s, int32_t which, int32_t* n_passes, int64_t* n_rows, int64_t* n_cols,
                   int64_t* smem_forward, int64_t* smem_adjoint);

/* Storage of the compact sector matrix: 8 = real amplitudes, 16 = complex, 0 = the handle works on the
 * dense register.  Every factor of the reference's ansatz (qc_uccsd/vqechem/uccsd_vqe.py:66-87, T - T^+ of
 * real excitation operators) is a real rotation of the occupation basis and a Jordan-Wigner Hamiltonian with
 * real integrals (dmet_interface.py:200-241) has real matrix elements, so psi and H psi never acquire an
 * imaginary part: the handle then keeps doubles (option "real_amplitudes", default 1) and returns to complex
 * storage by itself when a Hamiltonian term has an odd number of Y, a complex coefficient or no place in
 * the link tables, or when mqc_load_state is given a complex vector. */
int mqc_amplitude_bytes(mqc_solver_t* s, int32_t* bytes);

/* Replaces: the `fcivec` argument of mqc/tools/rdm.py:131-138 (make_rdm1) and :157-172
 * (make_rdm12): loads a caller-supplied wave function (OpenFermion order, 2^N amplitudes,
 * interleaved re/im) as the prepared state, so that mqc_rdm12 / mqc_expectation evaluate it.
 * Needs mqc_set_ansatz first (an empty factor list is fine: ref_index fixes the sector); with
 * compact storage the vector must lie in that (N_alpha, N_beta) sector. */
int mqc_load_state(mqc_solver_t* s, const double* re_im, int64_t n_amplitudes);

/* Device-side timing of the last energy / energy_grad call (CUDA events on the solver's
 * stream): ms[0] forward sweep, ms[1] H apply, ms[2] reverse sweep, ms[3] total;
 * counters[0] kernel launches, counters[1] forward passes, counters[2] reverse passes. */
int mqc_last_timing(mqc_solver_t* s, double* ms4, int64_t* counters3);

/* Host-only introspection of the sweep schedule (works without a GPU). */
int mqc_schedule_info(mqc_solver_t* s, int32_t which /*0 energy, 1 gradient*/, int32_t* n_passes,
                      int32_t* tile_bits, int32_t* n_tile_factors);
int mqc_schedule_pass(mqc_solver_t* s, int32_t which, int32_t pass, int32_t* first_factor,
                      int32_t* n_factors, uint64_t* act_mask, int32_t* pos16, int32_t* perm16,
                      int32_t* has_perm);
/* tiles rank `rank` sweeps in pass `pass`: tile counter t in [0, n_tiles) -> (t | tau_hi) is
 * deposited into the inactive LOCAL bits, OR fixed_bits (the rank's inactive global bits) */
int mqc_schedule_tile_map(mqc_solver_t* s, int32_t which, int32_t pass, int32_t rank, uint64_t* tau_hi,
                          uint64_t* fixed_bits, uint64_t* n_tiles);
int mqc_schedule_factor(mqc_solver_t* s, int32_t which, int32_t tile_factor, uint32_t* m, uint32_t* v,
                        uint32_t* z, uint64_t* z_out, int32_t* sign0, int32_t* slot);
int mqc_schedule_final_layout(mqc_solver_t* s, int32_t which, int32_t* phys_of_logical);
int mqc_schedule_initial_layout(mqc_solver_t* s, int32_t which, int32_t* phys_of_logical);

#ifdef __cplusplus
}
#endif
#endif
