"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Tolerances are BASELINE.json's: energies 1e-10 Ha, RDMs and gradients 1e-9
(fp64), state vectors 1e-12."""
import numpy as np
import pytest

from oracle_ansatz import OracleAnsatz
from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.jw import fragment_term_table, s2_term_table
from multiscalequantumcomputing_b200.solver import solve
from multiscalequantumcomputing_b200.synth.hydrogen import (hydrogen_ring, hydrogen_chain, local_integrals,
                                                             random_fragment_integrals)
from multiscalequantumcomputing_b200.synth.dmet_loop import DMETRun, atom_clusters, fragment_problems
from multiscalequantumcomputing_b200.uccsd import uccsd_factors
from oracle import rdm as ordm
from oracle.hamiltonian import apply_terms
from oracle.statevector import UCCSDOracle

pytestmark = pytest.mark.gpu

E_TOL, G_TOL, R_TOL, S_TOL = 1e-10, 1e-9, 1e-9, 1e-12


def _setup(n, ne, order="blocked", opts=None, seed=1234, s2=0.25):
    h1, eri = random_fragment_integrals(n, seed)
    tab = fragment_term_table(h1, eri, 0, 0.0, s2)
    F = uccsd_factors(n, ne, order)
    h = _lib.Handle(2 * n)
    for k, v in (opts or {}).items():
        h.set_option(k, v)
    h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    orc = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, tab)
    th = np.random.default_rng(4321).normal(0, 0.05, F.n_theta)
    return h, orc, F, tab, th


# options without "compact": 0 run the compact-sector path (k_csweep) whenever the ansatz conserves
# (N_alpha, N_beta); "compact": 0 keeps the dense-register kernels (sharded states, non-conserving
# ansaetze) under the same checks
CASES = [
    (4, 4, "blocked", {}),                                       # 8 qubits: whole register in one tile
    (4, 2, "lex", {"tile_bits": 6, "grad_tile_bits": 6}),         # compact, several passes of small tiles
    (5, 4, "blocked", {"tile_bits": 7, "grad_tile_bits": 6, "csweep_threads": 32}),
    (6, 6, "lex", {"tile_bits": 9, "grad_tile_bits": 8, "csweep_threads": 64}),
    (6, 6, "blocked", {"tile_bits": 8, "grad_tile_bits": 7, "csweep_ctas": 3}),    # persistent CTAs loop over many tiles
    (7, 6, "blocked", {"tile_bits": 10, "grad_tile_bits": 9, "csweep_threads": 256}),
    (8, 8, "lex", {"tile_bits": 12, "grad_tile_bits": 10}),
    (6, 2, "blocked", {"tile_bits": 8, "grad_tile_bits": 8}),                     # dilute sector
    (8, 8, "blocked", {"sigma": 0}),                                               # compact state + X-group H kernel
    (6, 6, "blocked", {"sigma_ab": 1, "tile_bits": 8, "grad_tile_bits": 8}),       # opposite-spin doubles in the sign-adjusted basis
    (7, 6, "lex", {"sigma_ab": 1, "sigma_stage": 0, "tile_bits": 10, "grad_tile_bits": 10}),
    (5, 4, "blocked", {"sigma_ab": 1}),
    (6, 6, "lex", {"tile_bits": 9, "grad_tile_bits": 8, "real_amplitudes": 0}),      # complex compact storage (default here: real)
    (8, 8, "blocked", {"real_amplitudes": 0, "sigma_split": 0}),
    (8, 8, "blocked", {"sigma_split": 8, "csweep_adj_threads": 128}),
    (8, 8, "blocked", {"sigma_stage_bytes": 1200}),                                # K2 stages two source rows at a time
    (8, 8, "lex", {"sigma_stage_bytes": 600, "real_amplitudes": 0}),               # complex: 1120-byte rows do not fit: unstaged
    (8, 8, "lex", {"sigma_stage_bytes": 600, "sigma_stage_one": 1}),               # ... one 560-byte row at a time
    (6, 4, "blocked", {"tile_bits": 8, "grad_tile_bits": 8, "csweep_adj_threads": 32}),   # real rows starting on odd columns, one warp per tile
    (4, 4, "blocked", {"compact": 0}),
    (4, 4, "lex", {"mode": 0}),                                  # one factor per pass
    (4, 2, "lex", {"tile_bits": 6, "grad_tile_bits": 6, "low_bits": 2, "compact": 0}),
    (5, 4, "blocked", {"tile_bits": 7, "grad_tile_bits": 7, "low_bits": 3, "compact": 0}),
    (6, 6, "blocked", {"tile_bits": 8, "grad_tile_bits": 7, "low_bits": 2, "compact": 0}),
    (6, 6, "lex", {"tile_bits": 9, "grad_tile_bits": 8, "low_bits": 3, "compact": 0}),
    (6, 6, "lex", {"mode": 0}),
    (6, 6, "blocked", {"sector": 0}),                            # dense H|psi> path
    (7, 6, "blocked", {"tile_bits": 10, "grad_tile_bits": 9, "compact": 0}),
    (8, 8, "blocked", {"compact": 0}),                            # 16 qubits, dense register, sector tiles
    (8, 8, "blocked", {"grad_tile_bits": 12, "tile_threads": 512, "compact": 0}),
    (6, 6, "blocked", {"tile_bits": 8, "grad_tile_bits": 7, "sp2": 0, "compact": 0}),           # first-generation sector tiles
    (8, 8, "blocked", {"sp2": 0, "sigma": 0, "compact": 0}),                                     # ... and X-group H kernel
    (7, 6, "lex", {"tile_bits": 10, "grad_tile_bits": 9, "sp2_warps": 2, "compact": 0}),
    (6, 2, "blocked", {"tile_bits": 8, "grad_tile_bits": 8, "compact": 0}),                     # dilute sector
    (7, 6, "blocked", {"tile_bits": 10, "grad_tile_bits": 10, "sp2_list": 0, "sigma_stage": 0, "compact": 0}),   # large-register paths
]


@pytest.mark.parametrize("n,ne,order,opts", CASES)
def test_state_energy_gradient(n, ne, order, opts):
    h, orc, F, tab, th = _setup(n, ne, order, opts)
    psi = h.state(th)
    ref = orc.state(th)
    assert abs(psi - ref).max() < S_TOL
    e_ref, g_ref = orc.energy_grad(th)
    assert abs(h.energy(th) - e_ref) < E_TOL
    e, g = h.energy_grad(th)
    assert abs(e - e_ref) < E_TOL
    assert abs(g - g_ref).max() < G_TOL
    t = h.last_timing()
    assert t["launches"] > 0 and t["forward_passes"] > 0 and t["reverse_passes"] > 0
    # the state is still retrievable after the reverse sweep consumed it
    assert abs(h.state() - ref).max() < S_TOL
    h.close()


def test_gradient_at_zero_angles():
    """theta = 0 is where every optimisation starts: sin = 0 must not lose the JW sign."""
    h, orc, F, tab, th = _setup(6, 6, "blocked", {"tile_bits": 8, "grad_tile_bits": 8})
    z = np.zeros(F.n_theta)
    e_ref, g_ref = orc.energy_grad(z)
    e, g = h.energy_grad(z)
    assert abs(e - e_ref) < E_TOL and abs(g - g_ref).max() < G_TOL
    assert np.abs(g_ref).max() > 1e-3
    h.close()


@pytest.mark.parametrize("n,ne,extra", [(4, 4, {}), (5, 4, {}), (6, 6, {}), (7, 6, {}),
                                        (6, 6, {"rdm_gram": 0}),          # compact state, literal K4 loops (atomics)
                                        (5, 4, {"compact": 0})])          # dense register
def test_rdm12(n, ne, extra):
    h, orc, F, tab, th = _setup(n, ne, "blocked", dict({"tile_bits": 9, "grad_tile_bits": 8}, **extra))
    h.prepare(th)
    r1, r2 = h.rdm12(n, ne // 2, ne - ne // 2)
    o1, o2 = ordm.get_rdm12(orc.state(th), ne, n)
    assert abs(r1 - o1).max() < R_TOL and abs(r2 - o2).max() < R_TOL
    assert abs(np.trace(r1) - ne) < 1e-10
    assert abs(np.einsum("iijj->", r2) - ne * (ne - 1)) < 1e-9
    h.close()


def test_rdm_golden_vector(golden):
    """The logged VQE wave function (test/testlog_vqe:56-94) loaded into the GPU (mqc_load_state)
    and pushed through the GPU RDM kernels reproduces the reference's logged rdm1 / rdm2
    (test/testlog_vqe:98-185) to the printed digits, and the oracle's float64 RDMs to 1e-12;
    both storage forms of the state (compact sector matrix, dense register)."""
    strs = np.array(golden["ring_strings"]["values"])
    fc = np.array(golden["ring_fcivec"]["values"])
    o1, o2 = ordm.make_rdm12(fc, 4, strs)
    psi = np.zeros(256, dtype=np.complex128)
    for st, c in zip(strs, fc):
        psi[ordm.qubit_inverse(4, int(st))] = c
    g1 = np.array(golden["ring_vqe_rdm1"]["values"])
    g2 = np.array(golden["ring_vqe_rdm2"]["values"])
    for compact in (1, 0):
        h = _lib.Handle(8)
        h.set_option("compact", compact)
        h.set_ansatz(np.zeros(0, np.int32), np.zeros((0, 4), np.int32), np.zeros(0, np.int32), 0, 240)
        h.load_state(psi)
        assert abs(h.state() - psi).max() == 0.0
        r1, r2 = h.rdm12(4, 2, 2)
        assert abs(r1 - g1).max() < 6e-4 and abs(r2 - g2).max() < 6e-4          # the reference's own dump
        assert abs(r1 - o1).max() < 1e-12 and abs(r2 - o2).max() < 1e-12        # float64 restatement of rdm.py
        # <H> of the loaded state through K2
        ints = local_integrals(hydrogen_ring(10, 1.0))
        p = fragment_problems(ints, atom_clusters(10, 2), 0.0, True)[0]
        tab = fragment_term_table(p.fock, p.tei, 0, 0.0, 0.0)
        e_k2 = h.expectation(*tab).real
        assert abs(e_k2 - np.vdot(psi, apply_terms(*tab, psi)).real) < 1e-12
        h.close()
    # a vector outside the handle's sector is refused by the compact storage
    h = _lib.Handle(8)
    h.set_ansatz(np.zeros(0, np.int32), np.zeros((0, 4), np.int32), np.zeros(0, np.int32), 0, 240)
    bad = np.zeros(256, dtype=np.complex128); bad[255] = 1.0
    with pytest.raises(_lib.MqcError):
        h.load_state(bad)
    h.close()


def test_real_and_complex_storage_selection():
    """The compact matrix holds doubles when the Hamiltonian is real (include/mqc_b200.h, mqc_amplitude_bytes) and
    goes back to complex storage by itself: option, a Hamiltonian term with an odd number of Y,
    a complex loaded state.  Results agree with the oracle in every case."""
    h, orc, F, tab, th = _setup(6, 6, "blocked", {"tile_bits": 8, "grad_tile_bits": 8})
    assert h.amplitude_bytes() == 8
    e_ref, g_ref = orc.energy_grad(th)
    e, g = h.energy_grad(th)
    assert abs(e - e_ref) < E_TOL and abs(g - g_ref).max() < G_TOL
    assert h.amplitude_bytes() == 8
    # i (a^+_0 a_2 - a^+_2 a_0) on the alpha orbitals 0 and 2 = (X0 Z Z Z Y4 - Y0 Z Z Z X4) / 2 in logical qubits
    # (bit n - 1 - q): Hermitian, imaginary matrix elements
    n = 12
    bit = lambda q: np.uint64(1) << np.uint64(n - 1 - q)
    zs = bit(1) | bit(2) | bit(3)
    x2 = np.concatenate([tab[0], [bit(0) | bit(4), bit(0) | bit(4)]]).astype(np.uint64)
    z2 = np.concatenate([tab[1], [zs | bit(4), zs | bit(0)]]).astype(np.uint64)
    c2 = np.concatenate([tab[2], [0.35, -0.35]])
    h.set_hamiltonian(x2, z2, c2)
    orc2 = UCCSDOracle(n, F.as_tuples(), F.theta_index, F.ref_index, (x2, z2, c2))
    e2_ref, g2_ref = orc2.energy_grad(th)
    e2, g2 = h.energy_grad(th)
    assert h.amplitude_bytes() == 16
    assert abs(e2 - e2_ref) < E_TOL and abs(g2 - g2_ref).max() < G_TOL
    assert abs(h.state(th) - orc2.state(th)).max() < S_TOL
    h.close()
    # real Hamiltonians with extra terms: X0 Z1 Z2 X4 (the JW string misses qubit 3: still a link-table group, stays
    # real) and a six-qubit flip X0 X2 X4 X6 X8 X10 (three-body: only the X-group kernel takes it, which works on
    # complex amplitudes: decided when the tables are built, at the first evaluation)
    for xs, zs2, want in ((bit(0) | bit(4), bit(1) | bit(2), 8),
                          (bit(0) | bit(2) | bit(4) | bit(6) | bit(8) | bit(10), np.uint64(0), 16)):
        h, orc, F, tab, th = _setup(6, 6, "blocked", {"tile_bits": 8, "grad_tile_bits": 8})
        x3 = np.concatenate([tab[0], [xs]]).astype(np.uint64)
        z3 = np.concatenate([tab[1], [zs2]]).astype(np.uint64)
        c3 = np.concatenate([tab[2], [0.2]])
        h.set_hamiltonian(x3, z3, c3)
        assert h.amplitude_bytes() == 8
        orc3 = UCCSDOracle(n, F.as_tuples(), F.theta_index, F.ref_index, (x3, z3, c3))
        e3_ref, g3_ref = orc3.energy_grad(th)
        e3, g3 = h.energy_grad(th)
        assert abs(e3 - e3_ref) < E_TOL and abs(g3 - g3_ref).max() < G_TOL
        assert abs(e3 - e_ref) > 1e-6                      # the extra term does contribute
        assert h.amplitude_bytes() == want
        h.close()
    # option; odd number of columns
    h, orc, F, tab, th = _setup(6, 6, "blocked", {"real_amplitudes": 0})
    assert h.amplitude_bytes() == 16
    h.close()
    h, orc, F, tab, th = _setup(7, 6, "blocked", {"tile_bits": 10, "grad_tile_bits": 9})     # C(7, 3) = 35 columns: padded row pitch
    assert h.amplitude_bytes() == 8
    e_ref, g_ref = orc.energy_grad(th)
    e, g = h.energy_grad(th)
    assert abs(e - e_ref) < E_TOL and abs(g - g_ref).max() < G_TOL
    h.prepare(th)
    r1, r2 = h.rdm12(7, 3, 3)
    o1, o2 = ordm.get_rdm12(orc.state(th), 6, 7)
    assert abs(r1 - o1).max() < R_TOL and abs(r2 - o2).max() < R_TOL
    assert abs(h.state() - orc.state(th)).max() < S_TOL
    h.close()
    h, orc, F, tab, th = _setup(5, 4, "blocked", {"compact": 0})
    assert h.amplitude_bytes() == 0
    h.close()
    # a complex state loaded into a real handle
    h, orc, F, tab, th = _setup(5, 4, "blocked", {})
    assert h.amplitude_bytes() == 8
    psi = orc.state(th) * np.exp(0.3j)
    h.load_state(psi)
    assert h.amplitude_bytes() == 16
    assert abs(h.state() - psi).max() == 0.0
    r1, r2 = h.rdm12(5, 2, 2)
    o1, o2 = ordm.get_rdm12(psi, 4, 5)
    assert abs(r1 - o1).max() < R_TOL and abs(r2 - o2).max() < R_TOL
    assert abs(h.expectation(*tab).real - np.vdot(psi, apply_terms(*tab, psi)).real) < 1e-11
    # ... and a real one stays real
    h2, orc, F, tab, th = _setup(5, 4, "blocked", {})
    h2.load_state(orc.state(th).real.astype(np.complex128))
    assert h2.amplitude_bytes() == 8
    r1b, r2b = h2.rdm12(5, 2, 2)
    assert abs(r1b - o1).max() < R_TOL and abs(r2b - o2).max() < R_TOL
    h.close(); h2.close()


def test_reproducible_energy_and_gradient():
    """Default option deterministic = 1: the gradient partials of the adjoint sweep and the energy partials of K2 are
    added in a fixed order, so the same input gives the same bits - on the same handle, on a fresh handle, and with
    tied parameters (several factors per parameter).  An optimiser or a chemical-potential search on top then takes the
    same path every run (scripts/dmet_adapt_spread.py: one value in every run; with deterministic = 0: 1e-9 ... 4e-8)."""
    for n, ne in ((6, 6), (8, 8)):
        h, orc, F, tab, th = _setup(n, ne, "blocked", {})
        e0, g0 = h.energy_grad(th)
        for _ in range(3):
            e1, g1 = h.energy_grad(th)
            assert e1 == e0 and np.array_equal(g1, g0)
        assert h.energy(th) == e0
        h2, *_ = _setup(n, ne, "blocked", {})
        e2, g2 = h2.energy_grad(th)
        assert e2 == e0 and np.array_equal(g2, g0)
        h.close(); h2.close()
    # tied parameters: every parameter drives three consecutive factors
    h1, eri = random_fragment_integrals(6, 1234)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    F = uccsd_factors(6, 6, "blocked")
    tied = (np.arange(len(F.theta_index)) // 3).astype(np.int32)
    n_theta = int(tied.max()) + 1
    th = np.random.default_rng(7).normal(0, 0.05, n_theta)
    res = []
    for _ in range(2):
        h = _lib.Handle(12)
        h.set_hamiltonian(*tab)
        h.set_ansatz(F.kind, F.idx, tied, n_theta, F.ref_index)
        res.append(h.energy_grad(th)); res.append(h.energy_grad(th))
        h.close()
    orc = UCCSDOracle(12, F.as_tuples(), tied, F.ref_index, tab)
    e_ref, g_ref = orc.energy_grad(th)
    assert abs(res[0][0] - e_ref) < E_TOL and abs(res[0][1] - g_ref).max() < G_TOL
    for e, g in res[1:]:
        assert e == res[0][0] and np.array_equal(g, res[0][1])


def _c_oracle_check(tab, F, th, e, g):
    from oracle.cref import CRefUCCSD, use_all_cores
    use_all_cores()
    ref = CRefUCCSD(F.n_qubits, F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index, tab)
    e_ref, g_ref, _ = ref.energy_grad_sector(th)
    assert abs(e - e_ref) < E_TOL, (e, e_ref)
    assert np.abs(g - g_ref).max() < G_TOL


def test_24_qubits_against_c_oracle():
    """BASELINE.json configs[2] at FULL size - the workload bench.py quotes (H12 chain, 24 qubits,
    1 818 parameters): energy 1e-10 and gradient 1e-9 against one complete evaluation of the CPU
    oracle (oracle/c, sector-aware variant, validated against the numpy oracle in test_oracle_c.py)."""
    import bench
    tab, F, th, _ = bench.workload()
    h = _lib.Handle(F.n_qubits)
    h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    assert h.cplan_info(1)["n_passes"] > 0                      # the compact-sector path is the one under test
    e, g = h.energy_grad(th)
    _c_oracle_check(tab, F, th, e, g)
    h.close()


def test_24_qubits_generic_instance_against_c_oracle():
    """The seeded generic C3 instance of SURVEY.md §8d (rng 1234: 29 737 Pauli terms in 5 479
    X-groups, twice the K2 work of the H12 chain), lexicographic factor order."""
    n = 12
    tab = fragment_term_table(*random_fragment_integrals(n, 1234), 0, 0.0, 0.25)
    assert len(tab[0]) > 29000
    F = uccsd_factors(n, n, "lex")
    th = np.random.default_rng(4321).normal(0, 0.05, F.n_theta)
    h = _lib.Handle(2 * n)
    h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    e, g = h.energy_grad(th)
    _c_oracle_check(tab, F, th, e, g)
    h.close()


def test_expectation_s2_and_linearity():
    h, orc, F, tab, th = _setup(6, 6, "blocked", {"tile_bits": 9, "grad_tile_bits": 8})
    h.prepare(th)
    psi = orc.state(th)
    s2 = s2_term_table(6)
    v = h.expectation(*s2)
    assert abs(v - np.vdot(psi, apply_terms(*s2, psi))) < 1e-10
    x, z, c = tab
    a = h.expectation(x, z, c)
    b = h.expectation(x, z, 2.5 * c)
    assert abs(b - 2.5 * a) < 1e-9
    assert abs(a.imag) < 1e-10
    h.close()


def test_properties_20_qubits():
    """Size-independent checks where the numpy oracle is slow: norm 1, gradient vs central
    finite differences of the GPU energy, zero angles give the reference energy."""
    n = 10
    h1, eri = random_fragment_integrals(n, 7)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    F = uccsd_factors(n, n, "blocked")
    h = _lib.Handle(2 * n)
    h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    th = np.random.default_rng(11).normal(0, 0.05, F.n_theta)
    psi = h.state(th)
    assert abs(np.vdot(psi, psi).real - 1.0) < 1e-12
    e, g = h.energy_grad(th)
    for k in (0, 17, F.n_theta - 1):
        d = np.zeros_like(th); d[k] = 1e-4
        fd = (h.energy(th + d) - h.energy(th - d)) / 2e-4
        assert abs(fd - g[k]) < 1e-6
    # fused tile sweep == one-factor-per-pass sweep
    h0 = _lib.Handle(2 * n)
    h0.set_option("mode", 0)                                # also leaves the compact storage: dense register
    h0.set_hamiltonian(*tab)
    h0.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    e0, g0 = h0.energy_grad(th)
    assert abs(e - e0) < E_TOL and abs(g - g0).max() < G_TOL
    h.close(); h0.close()


def test_properties_26_qubits_rdm_blocks():
    """13 orbitals: n^2 = 169 > 144, so the Gram-form K4 (mqc_rdm.cuh) runs its blocked variant
    (2 x 2 blocks of 128 rows).  No CPU oracle at this size: traces, symmetry and the energy
    contracted from the RDMs against <H> of K2."""
    n, ne = 13, 12
    h1, eri = random_fragment_integrals(n, 5)
    bare = fragment_term_table(h1, eri, 0, 0.0, 0.0)
    F = uccsd_factors(n, ne, "blocked")
    h = _lib.Handle(2 * n)
    h.set_hamiltonian(*bare)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    th = np.random.default_rng(11).normal(0, 0.05, F.n_theta)
    h.prepare(th)
    e_bare = h.expectation(*bare).real
    r1, r2 = h.rdm12(n, ne // 2, ne // 2)
    assert abs(np.trace(r1) - ne) < 1e-10 and abs(np.einsum("iijj->", r2) - ne * (ne - 1)) < 1e-8
    assert abs(r1 - r1.T).max() < 1e-12 and abs(r2 - r2.transpose(2, 3, 0, 1)).max() < 1e-12
    assert abs(np.einsum("pq,pq->", h1, r1) + 0.5 * np.einsum("pqrs,pqrs->", eri, r2) - e_bare) < 1e-9
    h.close()


def test_solve_matches_oracle_backed_solve():
    ints = local_integrals(hydrogen_chain(10, 0.9))
    p = fragment_problems(ints, atom_clusters(10, 2), 0.0, False)[2]
    e1, r1 = solve(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.01, warm_start=False)
    e2, r2 = solve(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.01, warm_start=False,
                   _ansatz_factory=lambda nq, tab, fac: OracleAnsatz(nq, tab, fac))
    assert abs(e1 - e2) < 1e-8 and abs(r1 - r2).max() < 1e-6


def test_dmet_energy_matches_oracle(golden):
    """DMET total with the GPU solver vs the same loop with the CPU oracle behind the same
    ansatz (1e-8 Ha, BASELINE.json), and vs the reference's FCI-solver log (UCCSD error)."""
    ints = local_integrals(hydrogen_ring(10, 1.0))
    cl = atom_clusters(10, 2)
    gpu = DMETRun(ints, cl, solver=lambda *a: solve(*a, warm_start=False), trans_inv=True)
    cpu = DMETRun(ints, cl, solver=lambda *a: solve(*a, warm_start=False,
                  _ansatz_factory=lambda nq, tab, fac: OracleAnsatz(nq, tab, fac)), trans_inv=True)
    e_gpu, e_cpu = gpu.run(), cpu.run()
    assert abs(e_gpu - e_cpu) < 1e-8
    assert abs(e_gpu - golden["ring_fci"]["dmet_energy"]) < 2e-4


def test_adapt_vqe_on_gpu_reaches_reference_fci(golden):
    """The reference's default ansatz (ADAPT-VQE, generalized singles+doubles pool) on the CUDA
    sweep: pool gradients come from one adjoint evaluation; the converged fragment energy and
    impurity energy reproduce the reference's FCI log (test/testlog_fci:8,100)."""
    ints = local_integrals(hydrogen_ring(10, 1.0))
    p = fragment_problems(ints, atom_clusters(10, 2), 0.0, True)[0]
    e_imp, rdm1, det = solve(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.0, return_details=True, warm_start=False,
                             options={"ansatz": "adapt", "adapt": {"nu": 10, "eps": 1e-7}, "opt": {"gtol": 1e-9}})
    assert abs(det.e_uccsd - golden["ring_fci"]["e_fci"][0]) < 1e-8
    assert abs(e_imp - golden["ring_fci"]["e_imp"][0]) < 1e-7


def test_dmet_with_adapt_matches_reference_fci_log(golden):
    """Whole DMET loop (chemical-potential secant iterations) with the ADAPT solver on the GPU:
    the total energy matches the reference's FCI-solver log to 1e-6 Ha or better."""
    ints = local_integrals(hydrogen_ring(10, 1.0))
    run = DMETRun(ints, atom_clusters(10, 2), trans_inv=True,
                  solver=lambda *a: solve(*a, warm_start=False, options={"ansatz": "adapt", "adapt": {"eps": 1e-6}}))
    e = run.run()
    assert abs(e - golden["ring_fci"]["dmet_energy"]) < 1e-6


def test_dmet_with_tight_adapt_matches_reference_fci_log_to_1e8(golden):
    """BASELINE.json: 'DMET energies matching the reference to 1e-8 Ha'.  With tight ADAPT / optimiser thresholds the
    whole DMET loop on the GPU reproduces the reference's FCI-solver log (test/testlog_fci:438,
    E = -5.373292466490414) at the 1e-8 Ha level: 7.5e-9 with the spin-orbital pool, 2.0e-8 with the spin-adapted
    pool in product form, the same value in every run (scripts/dmet_adapt_spread.py, profiles/dmet_adapt_spread_r02c.jsonl)
    now that the gradient and energy sums are reproducible.  With atomics (option deterministic = 0) 16 runs each gave
    1.0e-9 ... 4.2e-8 and 1.5e-9 ... 3.7e-8 (profiles/dmet_adapt_spread_r02.jsonl): the tolerance of the chemical-
    potential secant search (1e-8 on mu, qcdmet.py:232 - the reference's own criterion) times dE/dmu, steered by the last
    bits of the sums.  The bounds cover that spread with a factor 2.5 / 8."""
    ints = local_integrals(hydrogen_ring(10, 1.0))
    opts = {"ansatz": "adapt", "adapt": {"eps": 1e-8, "nu": 10, "pool": "generalized"}, "opt": {"gtol": 1e-10}}
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                    # BFGS reports "precision loss" at |g| ~ 1e-8, below its own tolerance
        run = DMETRun(ints, atom_clusters(10, 2), trans_inv=True, solver=lambda *a: solve(*a, options=opts))
        e = run.run()
        sa = {"ansatz": "adapt", "adapt": {"eps": 1e-7, "nu": 10, "pool": "sa"}, "opt": {"gtol": 1e-9}}
        e_sa = DMETRun(ints, atom_clusters(10, 2), trans_inv=True, solver=lambda *a: solve(*a, options=sa)).run()
    assert abs(e - golden["ring_fci"]["dmet_energy"]) < 1e-7
    assert abs(e_sa - golden["ring_fci"]["dmet_energy"]) < 3e-7


def test_properties_24_qubits_full_size():
    """BASELINE.json configs[2] at its full size (24 qubits, 1 818 parameters), where no CPU oracle
    finishes in seconds: size-independent properties.  Norm 1; gradient = central finite differences
    of the energy; sector-tile path = dense-tile path = 4-rank sharded state; RDM traces; and the
    energy contracted from the 1-/2-RDMs (K4) equals <H> from the Pauli-term kernel (K2)."""
    from multiscalequantumcomputing_b200.synth.hydrogen import mo_integrals
    n, ne = 12, 12
    h_mo, eri_mo, e_nuc, e_rhf = mo_integrals(hydrogen_chain(n, 1.0), ne)
    tab = fragment_term_table(h_mo, eri_mo, 0, 0.0, 0.25)
    bare = fragment_term_table(h_mo, eri_mo, 0, 0.0, 0.0)
    F = uccsd_factors(n, ne, "blocked")
    th = np.random.default_rng(4321).normal(0, 0.05, F.n_theta)
    h = _lib.Handle(2 * n)
    h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    e, g = h.energy_grad(th)
    ident = (np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.ones(1))
    assert abs(h.expectation(*ident).real - 1.0) < 1e-12
    for k in (0, 911, F.n_theta - 1):
        d = np.zeros_like(th); d[k] = 1e-4
        fd = (h.energy(th + d) - h.energy(th - d)) / 2e-4
        assert abs(fd - g[k]) < 2e-7
    # K4 against K2
    h.prepare(th)
    e_bare = h.expectation(*bare).real
    r1, r2 = h.rdm12(n, ne // 2, ne // 2)
    assert abs(np.trace(r1) - ne) < 1e-10 and abs(np.einsum("iijj->", r2) - ne * (ne - 1)) < 1e-8
    e_rdm = np.einsum("pq,pq->", h_mo, r1) + 0.5 * np.einsum("pqrs,pqrs->", eri_mo, r2)
    assert abs(e_rdm - e_bare) < 1e-9
    # other execution paths of the same evaluation
    hd = _lib.Handle(2 * n)
    hd.set_option("compact", 0)
    hd.set_option("sparse_tiles", 0)
    hd.set_hamiltonian(*tab)
    hd.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    ed, gd = hd.energy_grad(th)
    assert abs(ed - e) < E_TOL and abs(gd - g).max() < G_TOL
    hd.close()
    hs = _lib.ShardGroup(2 * n, [0, 0, 0, 0])
    hs.set_hamiltonian(*tab)
    hs.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    es, gs = hs.energy_grad(th)
    assert abs(es - e) < E_TOL and abs(gs - g).max() < G_TOL
    hs.close(); h.close()


def test_empty_ansatz_is_the_reference_determinant():
    """No factors at all: the state is the reference determinant, E = <ref|H|ref>, the gradient
    is empty (the starting point of ADAPT-VQE; `test/testlog_vqe:25` logs this energy)."""
    n, ne = 4, 4
    h1, eri = random_fragment_integrals(n, 3)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    ref_index = (1 << 8) - (1 << 4)
    h = _lib.Handle(2 * n)
    h.set_hamiltonian(*tab)
    h.set_ansatz(np.zeros(0, np.int32), np.zeros((0, 4), np.int32), np.zeros(0, np.int32), 0, ref_index)
    psi = np.zeros(256, dtype=np.complex128); psi[ref_index] = 1.0
    e_ref = np.vdot(psi, apply_terms(*tab, psi)).real
    assert abs(h.energy(np.zeros(0)) - e_ref) < E_TOL
    e, g = h.energy_grad(np.zeros(0))
    assert abs(e - e_ref) < E_TOL and len(g) == 0
    assert abs(h.state(np.zeros(0)) - psi).max() == 0.0
    h.close()


@pytest.mark.parametrize("n,occ", [(5, (0, 1, 2, 3, 4)), (6, (0, 2, 4, 6, 1))])
def test_open_shell_sector_and_generalized_factors(n, occ):
    """N_alpha != N_beta reference and generalized (not occupied->virtual) singles/doubles with a
    shared parameter: the sector machinery (tiles, link tables, RDM strings) is not tied to the
    closed-shell UCCSD list."""
    from multiscalequantumcomputing_b200.adapt import generalized_pool
    from multiscalequantumcomputing_b200.uccsd import FactorList
    nq = 2 * n
    h1, eri = random_fragment_integrals(n, 11)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    kind, idx = generalized_pool(n)
    rng = np.random.default_rng(9)
    pick = rng.choice(len(kind), size=60, replace=False)
    ref_index = sum(1 << (nq - 1 - q) for q in occ)
    theta_index = np.arange(60, dtype=np.int32)
    theta_index[-1] = 0                                   # two factors share theta_0
    F = FactorList(nq, kind[pick], idx[pick], theta_index, 59, ref_index)
    na = sum(1 for q in occ if q % 2 == 0)
    assert na != len(occ) - na
    h = _lib.Handle(nq)
    h.set_option("tile_bits", 8)
    h.set_option("grad_tile_bits", 8)
    h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    th = rng.normal(0, 0.3, F.n_theta)
    orc = UCCSDOracle(nq, F.as_tuples(), F.theta_index, F.ref_index, tab)
    assert abs(h.state(th) - orc.state(th)).max() < S_TOL
    e_ref, g_ref = orc.energy_grad(th)
    e, g = h.energy_grad(th)
    assert abs(e - e_ref) < E_TOL and abs(g - g_ref).max() < G_TOL
    r1, r2 = h.rdm12(n, na, len(occ) - na)
    assert abs(np.trace(r1) - len(occ)) < 1e-10
    assert abs(np.einsum("iijj->", r2) - len(occ) * (len(occ) - 1)) < 1e-9
    e_bare = h.expectation(*fragment_term_table(h1, eri, 0, 0.0, 0.0)).real
    assert abs(np.einsum("pq,pq->", h1, r1) + 0.5 * np.einsum("pqrs,pqrs->", eri, r2) - e_bare) < 1e-9
    h.close()


@pytest.mark.parametrize("n,ne", [(4, 3), (5, 5), (6, 3)])
def test_odd_electron_fragment(n, ne):
    """Open-shell fragment (noa = N // 2, nob = N - noa, `mqc/solver/dmet_interface.py:84-85`):
    state, energy, gradient and RDMs against the oracle; then the drop-in solve() on it."""
    h, orc, F, tab, th = _setup(n, ne, "blocked", {"tile_bits": 7, "grad_tile_bits": 6})
    assert abs(h.state(th) - orc.state(th)).max() < S_TOL
    e_ref, g_ref = orc.energy_grad(th)
    e, g = h.energy_grad(th)
    assert abs(e - e_ref) < E_TOL and abs(g - g_ref).max() < G_TOL
    h.prepare(th)
    r1, r2 = h.rdm12(n, ne // 2, ne - ne // 2)
    b1, b2 = ordm.rdm12_bruteforce(orc.state(th), n)
    assert abs(r1 - b1).max() < R_TOL and abs(r2 - b2).max() < R_TOL
    assert abs(np.trace(r1) - ne) < 1e-10
    h.close()
    h1, eri = random_fragment_integrals(n, 21)
    e1, d1 = solve(h1, h1, eri, ne, 2, 0.01)
    e2, d2 = solve(h1, h1, eri, ne, 2, 0.01, _ansatz_factory=lambda nq, t, f: OracleAnsatz(nq, t, f))
    assert abs(e1 - e2) < 1e-8 and abs(d1 - d2).max() < 1e-6
    assert abs(np.trace(d1) - ne) < 1e-8


def test_update_hamiltonian_coeffs_and_handle_cache():
    """Per-fragment persistent state: a second Hamiltonian with the same Pauli
    strings only sends coefficients (mqc_update_hamiltonian_coeffs) and gives the same numbers as
    a fresh handle; solve(fragment_id=...) creates ONE handle per fragment over a whole DMET run."""
    from multiscalequantumcomputing_b200 import solver as S
    n, ne = 6, 6
    h, orc, F, tab, th = _setup(n, ne, "blocked", {"tile_bits": 8, "grad_tile_bits": 8})
    e0, g0 = h.energy_grad(th)
    h1, eri = random_fragment_integrals(n, 77)
    tab2 = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    assert np.array_equal(tab2[0], tab[0]) and np.array_equal(tab2[1], tab[1])
    h.update_hamiltonian_coeffs(tab2[2])
    e1, g1 = h.energy_grad(th)
    orc2 = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, tab2)
    e_ref, g_ref = orc2.energy_grad(th)
    assert abs(e1 - e_ref) < E_TOL and abs(g1 - g_ref).max() < G_TOL and abs(e1 - e0) > 1e-3
    h.update_hamiltonian_coeffs(tab[2])
    e2, g2 = h.energy_grad(th)
    assert abs(e2 - e0) < 1e-13 and abs(g2 - g0).max() < 1e-13
    with pytest.raises(_lib.MqcError):
        h.update_hamiltonian_coeffs(tab[2][:-1])
    h.close()
    # whole DMET loop: handles are created once per fragment, every later mu step re-uses them
    ints = local_integrals(hydrogen_chain(10, 0.9))
    cl = atom_clusters(10, 2)
    S.clear_handle_cache(); S.clear_warm_start()
    S.STATS["handle_creates"] = S.STATS["handle_reuses"] = 0
    from multiscalequantumcomputing_b200.parallel import solve_fragments
    run = DMETRun(ints, cl, batch_solver=lambda ps: solve_fragments(ps, devices=[0], fragment_ids=True, warm_start=False))
    e_cached = run.run()
    assert S.STATS["handle_creates"] == len(cl)
    assert S.STATS["handle_reuses"] == run.n_solves - len(cl) and S.STATS["handle_reuses"] > 0
    ref = DMETRun(ints, cl, solver=lambda *a: solve(*a))
    e_ref = ref.run()
    # re-used handles run the same arithmetic as fresh ones; two FRESH runs of a five-fragment chain already differ by
    # up to 2.5e-8 (8 runs, H10 chain: the order of the gradient atomics is not fixed, BFGS stops at gtol-different
    # points and the chemical-potential secant search, tolerance 1e-8 on mu, amplifies the last digits)
    assert abs(e_cached - e_ref) < 1e-7
    # with the per-fragment warm start the optimiser stops at a (gtol-)different point of the same minimum
    warm = DMETRun(ints, cl, batch_solver=lambda ps: solve_fragments(ps, devices=[0], fragment_ids=True))
    assert abs(warm.run() - e_ref) < 1e-6
    S.clear_handle_cache(); S.clear_warm_start()
    with pytest.raises(ValueError):
        solve(ints.h_loc[:4, :4], ints.h_loc[:4, :4], ints.eri_loc[:4, :4, :4, :4], 4, 2, 0.0, warm_start=True)


def test_mbe_vqechem_and_measurement_rdms():
    """Second consumer of the kernels: molecule-level VQE with the reference's active-space
    rule (`mqc/solver/mbe_solver.py:179-217`) GPU vs oracle-backed, and the measurement-style RDM path
    (`mqc/solver/dmet_solver_vqechem.py:119-179`: one expectation value per operator, K2) against K4."""
    from multiscalequantumcomputing_b200 import mbe
    from multiscalequantumcomputing_b200.synth.hydrogen import mo_integrals
    h, g, e_nuc, e_rhf = mo_integrals(hydrogen_chain(6, 1.0), 6)
    oracle = dict(_ansatz_factory=lambda nq, tab, fac: OracleAnsatz(nq, tab, fac))
    for ncas in (None, 4):
        e_gpu = mbe.vqechem(h, g, 6, e_nuc, ncas=ncas)
        e_cpu = mbe.vqechem(h, g, 6, e_nuc, ncas=ncas, **oracle)
        assert abs(e_gpu - e_cpu) < 1e-8
    e_adapt = mbe.vqechem(h, g, 6, e_nuc, ncas=4, options={"ansatz": "adapt", "adapt": {"eps": 1e-6}})
    assert e_adapt < e_gpu + 1e-9                              # ADAPT reaches (at least) the UCCSD energy of the same space
    h4, g4, en4, _ = mo_integrals(hydrogen_chain(4, 1.0), 4)
    e, ans, th = mbe.vqechem(h4, g4, 4, en4, return_ansatz=True)
    m1, m2 = mbe.get_rdm12_openfermion(ans, 4)
    r1, r2 = ans.rdm12(4, 2, 2)
    assert abs(m2 - r2).max() < R_TOL
    assert abs(m1 - r1).max() < 1e-4 and abs(np.trace(m1) - 4.0) < 1e-6
    ans.close()


def test_dressed_singles_and_angle_scales_on_gpu():
    """The building blocks of a spin-adapted pool operator in product form (mqc_set_ansatz_scaled): number-dressed
    singles (kind 3), per-factor angle scales, tied parameters - state, energy and gradient against the oracle."""
    n = 5
    rng = np.random.default_rng(21)
    kind, idx = [], []
    for _ in range(70):
        r = rng.random()
        s1, s2 = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        if r < 0.35:
            p, q = rng.choice(n, 2, replace=False)
            c = int(rng.integers(0, 2 * n))
            while c in (2 * p + s1, 2 * q + s1):
                c = int(rng.integers(0, 2 * n))
            kind.append(3); idx.append((2 * p + s1, 2 * q + s1, c, -1))
        elif r < 0.55:
            p, q = rng.choice(n, 2, replace=False)
            kind.append(1); idx.append((2 * p + s1, 2 * q + s1, -1, -1))
        else:
            while True:
                p, q, r2, t = (int(v) for v in rng.integers(0, n, 4))
                if len({2 * p + s1, 2 * q + s2, 2 * r2 + s1, 2 * t + s2}) == 4:
                    break
            kind.append(2); idx.append((2 * p + s1, 2 * q + s2, 2 * r2 + s1, 2 * t + s2))
    kind = np.array(kind, np.int32); idx = np.array(idx, np.int32)
    ti = (np.arange(len(kind)) // 4).astype(np.int32)
    nth = int(ti.max()) + 1
    scale = rng.normal(0, 1, len(kind))
    ref_index = 0b1111000000
    h1, eri = random_fragment_integrals(n, 9)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    tup = [(int(k), tuple(int(v) for v in ix[:{1: 2, 2: 4, 3: 3}[int(k)]])) for k, ix in zip(kind, idx)]
    orc = UCCSDOracle(2 * n, tup, ti, ref_index, tab, scale)
    th = rng.normal(0, 0.3, nth)
    e_ref, g_ref = orc.energy_grad(th)
    for opts in ({"tile_bits": 6, "grad_tile_bits": 6}, {}):
        h = _lib.Handle(2 * n)
        for k, v in opts.items():
            h.set_option(k, v)
        h.set_hamiltonian(*tab)
        h.set_ansatz(kind, idx, ti, nth, ref_index, theta_scale=scale)
        assert abs(h.state(th) - orc.state(th)).max() < S_TOL
        e, g = h.energy_grad(th)
        assert abs(e - e_ref) < E_TOL and abs(g - g_ref).max() < G_TOL
        h.close()


def test_spin_adapted_pool_adapt_trajectory(golden):
    """ADAPT with the reference's pool ('spin_sym': 'sa': 66 operators, `test/testlog_vqe:23-24`), growth step Nu = 10
    and threshold 1e-3 (`dmet_solver_vqechem.py:39-40`): two growth iterations, then converged - as in the
    reference's log (`test/testlog_vqe:28-55`) - and the energy agrees with its FCI value."""
    from multiscalequantumcomputing_b200.adapt import ADAPTAnsatz
    ints = local_integrals(hydrogen_ring(10, 1.0))
    p = fragment_problems(ints, atom_clusters(10, 2), 0.0, True)[0]
    from multiscalequantumcomputing_b200.solver import imp_optimizer, prepare_problem
    prob = prepare_problem(p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.0)
    a = ADAPTAnsatz(8, prob["table"], p.nelec, nu=10, eps=1e-3, pool="sa")
    assert len(a.pool_ops) == 66
    a, params = imp_optimizer({"opt": {"gtol": 1e-8}}).run(a)
    assert a.converged and len(a.pool_grad_history) == 3          # two growth steps + the converged check
    a.prepare(params)
    assert abs(a.expectation(prob["table_bare"]).real - golden["ring_fci"]["e_fci"][0]) < 1e-7
    a.close()
