"""Host logic of the drop-in solve() with the oracle plugged in behind the ansatz surface
(CPU only; the product path itself always runs the CUDA library)."""
import numpy as np

from oracle_ansatz import OracleAnsatz
from multiscalequantumcomputing_b200.solver import solve, prepare_problem
from multiscalequantumcomputing_b200.synth.hydrogen import hydrogen_ring, hydrogen_chain, local_integrals
from multiscalequantumcomputing_b200.synth.dmet_loop import DMETRun, atom_clusters, fragment_problems
from oracle.fci import solve_fci

ORACLE = dict(_ansatz_factory=lambda nq, tab, fac: OracleAnsatz(nq, tab, fac), warm_start=False)


def test_uccsd_fragment_close_to_fci(golden):
    ints = local_integrals(hydrogen_ring(10, 1.0))
    p = fragment_problems(ints, atom_clusters(10, 2), 0.0, True)[0]
    e_imp, rdm1, det = solve(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.0, return_details=True, **ORACLE)
    assert det.n_theta == 26
    assert 0.0 <= det.e_uccsd - golden["ring_fci"]["e_fci"][0] < 1e-4      # variational, near-exact
    assert abs(e_imp - golden["ring_fci"]["e_imp"][0]) < 1e-4
    assert abs(np.trace(rdm1) - 4.0) < 1e-10
    assert abs(rdm1 - rdm1.T).max() < 1e-10


def test_chemical_potential_enters_hamiltonian():
    ints = local_integrals(hydrogen_chain(10, 0.8))
    p = fragment_problems(ints, atom_clusters(10, 2), 0.0, False)[1]
    a = prepare_problem(p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.0)
    b = prepare_problem(p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.05)
    assert len(a["table"][0]) == len(b["table"][0])
    assert abs(a["table"][2] - b["table"][2]).max() > 1e-3


def test_dmet_with_uccsd_solver_close_to_fci_logs(golden):
    ints = local_integrals(hydrogen_ring(10, 1.0))
    run = DMETRun(ints, atom_clusters(10, 2), solver=lambda *a: solve(*a, **ORACLE), trans_inv=True)
    e = run.run()
    assert abs(e - golden["ring_fci"]["dmet_energy"]) < 2e-4


def test_adapt_reaches_fci(golden):
    """ADAPT-VQE (the reference's default ansatz) driven through the same
    optimiser loop: pool gradients from one adjoint evaluation of "ansatz + pool at zero angle",
    energy converges to the logged FCI value of the fragment (test/testlog_fci:8,100)."""
    ints = local_integrals(hydrogen_ring(10, 1.0))
    p = fragment_problems(ints, atom_clusters(10, 2), 0.0, True)[0]
    e_imp, rdm1, det = solve(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.0, return_details=True,
                             options={"ansatz": "adapt", "adapt": {"nu": 10, "eps": 1e-7}, "opt": {"gtol": 1e-9}}, **ORACLE)
    assert abs(det.e_uccsd - golden["ring_fci"]["e_fci"][0]) < 1e-8
    assert abs(e_imp - golden["ring_fci"]["e_imp"][0]) < 1e-7
    assert det.n_theta <= 40                                  # parameters = spin-adapted pool operators


def test_adapt_pool_gradient_is_the_commutator():
    """Pool gradient of operator k = dE/dtheta_k at zero angle = finite difference of the energy of
    the ansatz with that one operator appended."""
    from multiscalequantumcomputing_b200.adapt import ADAPTAnsatz
    ints = local_integrals(hydrogen_chain(6, 0.9))
    p = fragment_problems(ints, atom_clusters(6, 2), 0.0, False)[1]
    prob = prepare_problem(p.fock, p.tei, p.nelec, p.n_imp_orbs, 0.0)
    a = ADAPTAnsatz(2 * prob["n_orb"], prob["table"], p.nelec, backend=lambda nq, tab, fac: OracleAnsatz(nq, tab, fac))
    e0, pg = a.pool_gradients(np.zeros(0))
    k = int(np.argmax(np.abs(pg)))
    a.chosen = [k]
    a._program(a._factors(with_pool=False))
    d = 1e-5
    fd = (a.energy(np.array([d])) - a.energy(np.array([-d]))) / (2 * d)
    assert abs(fd - pg[k]) < 1e-7 and abs(pg[k]) > 1e-3
