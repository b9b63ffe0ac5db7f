"""The C restatement (oracle/c/uccsd_ref.c) against the numpy oracle, and its sector-aware
variant (the CPU baseline of bench.py) against the dense one.  CPU only."""
import numpy as np
import pytest

from multiscalequantumcomputing_b200.jw import fragment_term_table
from multiscalequantumcomputing_b200.synth.hydrogen import random_fragment_integrals
from multiscalequantumcomputing_b200.uccsd import uccsd_factors
from oracle.cref import CRefUCCSD
from oracle.statevector import UCCSDOracle


def _problem(n, n_elec, seed=1234):
    h1, eri = random_fragment_integrals(n, seed)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    F = uccsd_factors(n, n_elec, "blocked")
    th = np.random.default_rng(4321).normal(0, 0.05, F.n_theta)
    return tab, F, th


@pytest.mark.parametrize("n,n_elec", [(4, 4), (5, 4), (6, 6)])
def test_c_dense_matches_numpy_oracle(n, n_elec):
    tab, F, th = _problem(n, n_elec)
    orc = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, tab)
    ref = CRefUCCSD(2 * n, F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index, tab)
    e0, g0 = orc.energy_grad(th)
    e1, g1 = ref.energy_grad(th)
    assert abs(e0 - e1) < 1e-12 and np.abs(g0 - g1).max() < 1e-12
    assert np.abs(orc.state(th) - ref.state(th)).max() < 1e-14


@pytest.mark.parametrize("n,n_elec", [(4, 4), (6, 4), (6, 6), (8, 8)])
def test_c_sector_variant_matches_dense(n, n_elec):
    tab, F, th = _problem(n, n_elec)
    ref = CRefUCCSD(2 * n, F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index, tab)
    e1, g1 = ref.energy_grad(th)
    e2, g2, phases = ref.energy_grad_sector(th)
    assert abs(e1 - e2) < 1e-12 and np.abs(g1 - g2).max() < 1e-12
    assert phases.shape == (3,) and (phases >= 0).all()
    from math import comb
    assert len(ref.sector_list()) == comb(n, n_elec // 2) * comb(n, n_elec - n_elec // 2)
