"""Internal consistency of the oracle pieces that the reference's logs cannot pin
(UCCSD energy / gradient are 'parity unpinned', see oracle/statevector.py)."""
import numpy as np
import pytest
import scipy.linalg as sl

from multiscalequantumcomputing_b200.jw import fragment_term_table
from multiscalequantumcomputing_b200.synth.hydrogen import random_fragment_integrals
from multiscalequantumcomputing_b200.uccsd import uccsd_factors
from oracle.statevector import (UCCSDOracle, SparseReferenceStyle, apply_excitation, apply_generator,
                                excitation_generator_matrix)


@pytest.mark.parametrize("kind,idx", [(1, (0, 4)), (1, (5, 1)), (2, (0, 1, 4, 5)), (2, (0, 3, 2, 5)),
                                      (2, (1, 2, 3, 4)), (2, (3, 0, 5, 2))])
def test_excitation_is_exponential_of_jw_generator(kind, idx):
    n = 6
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    A = excitation_generator_matrix(kind, idx, n)
    assert abs(A + A.conj().T).max() == 0
    assert abs(sl.expm(0.37 * A) @ psi - apply_excitation(kind, idx, 0.37, psi, n)).max() < 1e-13
    assert abs(A @ psi - apply_generator(kind, idx, psi, n)).max() < 1e-14


def _problem(n, seed=1234):
    h1, eri = random_fragment_integrals(n, seed)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    F = uccsd_factors(n, n if n % 2 == 0 else n - 1)
    return tab, F


def test_gradient_matches_finite_differences():
    tab, F = _problem(3)
    orc = UCCSDOracle(6, F.as_tuples(), F.theta_index, F.ref_index, tab)
    rng = np.random.default_rng(1)
    th = rng.normal(0, 0.1, F.n_theta)
    e, g = orc.energy_grad(th)
    for k in range(F.n_theta):
        d = np.zeros_like(th); d[k] = 1e-5
        fd = (orc.energy(th + d) - orc.energy(th - d)) / 2e-5
        assert abs(fd - g[k]) < 1e-8


def test_sparse_reference_style_agrees():
    tab, F = _problem(4)
    orc = UCCSDOracle(8, F.as_tuples(), F.theta_index, F.ref_index, tab)
    ref = SparseReferenceStyle(8, F.as_tuples(), F.theta_index, F.ref_index, tab)
    th = np.random.default_rng(2).normal(0, 0.1, F.n_theta)
    e1, g1 = orc.energy_grad(th)
    e2, g2 = ref.energy_grad(th)
    assert abs(e1 - e2) < 1e-12 and abs(g1 - g2).max() < 1e-11


def test_shared_theta_gradient():
    tab, F = _problem(3)
    ti = F.theta_index // 2
    orc = UCCSDOracle(6, F.as_tuples(), ti, F.ref_index, tab)
    th = np.random.default_rng(3).normal(0, 0.1, int(ti.max()) + 1)
    e, g = orc.energy_grad(th)
    d = np.zeros_like(th); d[1] = 1e-5
    assert abs((orc.energy(th + d) - orc.energy(th - d)) / 2e-5 - g[1]) < 1e-8
