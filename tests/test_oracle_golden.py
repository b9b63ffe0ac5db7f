"""The oracle against every golden vector the reference's logs hold for the hot path."""
import numpy as np

from multiscalequantumcomputing_b200.synth.hydrogen import hydrogen_chain, hydrogen_ring, local_integrals
from multiscalequantumcomputing_b200.synth.dmet_loop import atom_clusters, fragment_problems
from oracle import rdm as ordm
from oracle.fci import solve_fci, sector_indices
from oracle.hamiltonian import qubit_hamiltonian, qubit_s2, term_table, dense_matrix


def _ring_problem():
    ints = local_integrals(hydrogen_ring(10, 1.0))
    return fragment_problems(ints, atom_clusters(10, 2), 0.0, True)[0]


def test_spectrum_ring(golden):
    """'FCI Energy: State 0..5' = lowest six of H + 0.75 S^2 over the full Fock space
    (dmet_interface.py:121,246; test/testlog_vqe:14-19)."""
    p = _ring_problem()
    H = dense_matrix(*term_table(qubit_hamiltonian(p.fock, p.tei, p.n_imp_orbs), 8), 8)
    S2 = dense_matrix(*term_table(qubit_s2(4), 8), 8)
    assert abs(H - H.conj().T).max() < 1e-13
    w, v = np.linalg.eigh(H + 0.75 * S2)
    g = golden["ring_fragment_spectrum_H_plus_0p75_S2"]
    assert np.allclose(np.sort(w[:6]), np.sort(g["values"]), atol=5e-9)
    s2 = [np.vdot(v[:, k], S2 @ v[:, k]).real for k in range(6)]
    assert np.allclose(np.sort(s2), np.sort(np.abs(g["s2"])), atol=1e-7)


def test_spectrum_chain(golden):
    ints = local_integrals(hydrogen_chain(10, 0.6))
    p = fragment_problems(ints, atom_clusters(10, 2), 0.0, False)[0]
    H = dense_matrix(*term_table(qubit_hamiltonian(p.fock, p.tei, p.n_imp_orbs), 8), 8)
    S2 = dense_matrix(*term_table(qubit_s2(4), 8), 8)
    w = np.linalg.eigvalsh(H + 0.75 * S2)
    assert np.allclose(np.sort(w[:6]), np.sort(golden["chain06_frag0_spectrum_H_plus_0p75_S2"]["values"]), atol=5e-9)


def test_reference_determinant_energy(golden):
    """'HF energy' = <240|H|240>, test/testlog_vqe:25."""
    p = _ring_problem()
    H = dense_matrix(*term_table(qubit_hamiltonian(p.fock, p.tei, p.n_imp_orbs), 8), 8)
    assert abs(H[240, 240].real - golden["ring_hf_energy"]["value"]) < 1e-9


def test_strings_and_gather(golden):
    """make_strings_qubit order + qubit_inverse gather, test/testlog_vqe:56-94."""
    strs = ordm.make_strings_qubit(4, 4)
    assert strs.tolist() == golden["ring_strings"]["values"]
    wfn = dict(zip(golden["ring_vqe_wfn"]["index"], golden["ring_vqe_wfn"]["amp"]))
    gathered = [wfn[ordm.qubit_inverse(4, int(s))] for s in strs]
    assert np.allclose(gathered, golden["ring_fcivec"]["values"], atol=0, rtol=0)
    assert ordm.qubit_inverse(4, int(strs[0])) == 240


def test_rdm_dump(golden):
    """make_rdm12 on the logged fcivec reproduces the logged rdm1 / rdm2 (4 printed digits)."""
    strs = np.array(golden["ring_strings"]["values"])
    fc = np.array(golden["ring_fcivec"]["values"])
    r1, r2 = ordm.make_rdm12(fc, 4, strs)
    g1 = np.array(golden["ring_vqe_rdm1"]["values"])
    g2 = np.array(golden["ring_vqe_rdm2"]["values"])
    assert abs(r1 - g1).max() < 6e-4 and abs(r2 - g2).max() < 6e-4
    assert abs(np.trace(r1) - 4.0) < 1e-6
    assert abs(np.einsum("iijj->", r2) - 12.0) < 1e-5          # N(N-1)
    assert abs(r2 - r2.transpose(2, 3, 0, 1)).max() < 1e-12


def test_rdm_bruteforce_agrees(golden):
    strs = np.array(golden["ring_strings"]["values"])
    fc = np.array(golden["ring_fcivec"]["values"])
    psi = np.zeros(256, dtype=np.complex128)
    for s, c in zip(strs, fc):
        psi[ordm.qubit_inverse(4, int(s))] = c
    r1, r2 = ordm.make_rdm12(fc, 4, strs)
    b1, b2 = ordm.rdm12_bruteforce(psi, 4)
    assert abs(r1 - b1).max() < 1e-13 and abs(r2 - b2).max() < 1e-13


def test_fci_fragment_golden(golden):
    """E FCI, impurity energy at mu = 0: test/testlog_fci:8,100."""
    p = _ring_problem()
    e_imp, rdm1, x = solve_fci(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.0, True)
    assert abs(x["e_fci"] - golden["ring_fci"]["e_fci"][0]) < 1e-11
    assert abs(e_imp - golden["ring_fci"]["e_imp"][0]) < 1e-11
    # support of the exact state = the 36 indices of the logged VQE wave function
    strs, idx = sector_indices(4, 4)
    assert sorted(idx.tolist()) == sorted(golden["ring_vqe_wfn"]["index"])
    # energy from the RDMs equals <H>
    e2 = np.einsum("ij,ij->", p.fock, rdm1) + 0.5 * np.einsum("ijkl,ijkl->", p.tei, x["rdm2"])
    assert abs(e2 - x["e_fci"]) < 1e-12


def test_fci_with_chemical_potential(golden):
    p = _ring_problem()
    mu = golden["ring_fci"]["mu_nelec"][1][0]
    e_imp, rdm1, x = solve_fci(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, mu, True)
    assert abs(x["e_fci"] - golden["ring_fci"]["e_fci"][1]) < 1e-11
    assert abs(e_imp - golden["ring_fci"]["e_imp"][1]) < 1e-11
