"""Synthetic H/STO-3G inputs and the DMET caller, pinned to the reference's logs."""
import numpy as np
import pytest

from multiscalequantumcomputing_b200.synth.hydrogen import (hydrogen_chain, hydrogen_ring, local_integrals,
                                                             random_fragment_integrals, BOHR)
from multiscalequantumcomputing_b200.synth.dmet_loop import DMETRun, atom_clusters, fragment_problems
from oracle.fci import solve_fci


def test_bohr_constant():
    # test/tmp.molden:5 : 3.0 Angstrom = 5.66917837369518 a.u.
    assert abs(3.0 / BOHR - 5.66917837369518) < 1e-12


def test_rhf_ring_golden(golden):
    ints = local_integrals(hydrogen_ring(10, 1.0))
    assert abs(ints.e_rhf - golden["ring_rhf"]["value"]) < 1e-11


def test_rhf_chain_golden(golden):
    ints = local_integrals(hydrogen_chain(10, 0.6))
    assert abs(ints.e_rhf - golden["chain06_rhf"]["value"]) < 1e-11


@pytest.mark.parametrize("k", [0, 1, 2, 4, 6, 9])
def test_rhf_chain_sweep_golden(golden, k):
    d = 0.6 + 0.1 * k
    ints = local_integrals(hydrogen_chain(10, d))
    assert abs(ints.e_rhf - golden["chain_rhf_sweep"]["values"][k]) < 1e-10


def test_fragment_problem_shape():
    ints = local_integrals(hydrogen_chain(10, 0.6))
    probs = fragment_problems(ints, atom_clusters(10, 2), 0.0)
    assert len(probs) == 5
    for p in probs:
        assert p.fock.shape == (4, 4) and p.tei.shape == (4, 4, 4, 4)
        assert p.nelec == 4 and p.n_imp_orbs == 2          # "Performing a ( 4 orb, 4 el )" log_hchain_vqe:4


def test_dmet_ring_fci_golden(golden):
    """Whole DMET loop with an exact fragment solver reproduces test/testlog_fci."""
    g = golden["ring_fci"]
    ints = local_integrals(hydrogen_ring(10, 1.0))
    run = DMETRun(ints, atom_clusters(10, 2), solver=solve_fci, trans_inv=True)
    e = run.run()
    assert abs(e - g["dmet_energy"]) < 1e-9
    for (mu, ne, etot), (gmu, gne), ge in zip(run.history, g["mu_nelec"], g["e_tot"]):
        assert abs(mu - gmu) < 1e-9 and abs(ne - gne) < 1e-9 and abs(etot - ge) < 1e-9


@pytest.mark.parametrize("k", [0, 3])
def test_dmet_chain_fci_golden(golden, k):
    """test/log_hchain_dmet_fci:54950 (5 fragments, not translation invariant)."""
    d = 0.6 + 0.1 * k
    ints = local_integrals(hydrogen_chain(10, d))
    run = DMETRun(ints, atom_clusters(10, 2), solver=solve_fci, trans_inv=False)
    e = run.run()
    assert abs(e - golden["chain_dmet_fci_sweep"]["values"][k]) < 2e-8


def test_random_fragment_symmetry():
    h1, eri = random_fragment_integrals(5, 1234)
    assert np.allclose(h1, h1.T)
    assert np.allclose(eri, eri.transpose(1, 0, 2, 3)) and np.allclose(eri, eri.transpose(2, 3, 0, 1))
    assert np.allclose(eri, eri.transpose(0, 1, 3, 2))
