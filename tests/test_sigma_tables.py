"""Excitation link tables of the sector Hamiltonian kernel (csrc/mqc_sigma.h), evaluated on the
host through the C ABI, against the oracle's Pauli-term application on the full 2^N vector."""
from itertools import combinations

import numpy as np
import pytest

from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.jw import fragment_term_table, s2_term_table
from multiscalequantumcomputing_b200.synth.hydrogen import hydrogen_chain, mo_integrals, random_fragment_integrals
from oracle.hamiltonian import apply_terms


def _strings(n_orb, k):
    return sorted(sum(1 << p for p in c) for c in combinations(range(n_orb), k))


def _sector_index(n_orb, na, nb):
    """amplitude index (OpenFermion order) of determinant (ia, ib); alpha orbital p = qubit 2p"""
    n = 2 * n_orb
    A, B = _strings(n_orb, na), _strings(n_orb, nb)
    idx = np.zeros((len(A), len(B)), dtype=np.int64)
    for i, a in enumerate(A):
        for j, b in enumerate(B):
            v = 0
            for p in range(n_orb):
                if (a >> p) & 1:
                    v |= 1 << (n - 1 - 2 * p)
                if (b >> p) & 1:
                    v |= 1 << (n - 1 - (2 * p + 1))
            idx[i, j] = v
    return idx


def _check(n_orb, na, nb, tab, expect_fallback=0):
    n = 2 * n_orb
    idx = _sector_index(n_orb, na, nb)
    rng = np.random.default_rng(3)
    c = rng.normal(size=idx.shape) + 1j * rng.normal(size=idx.shape)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[idx] = c
    ref = apply_terms(*tab, psi)[idx]
    out, shape, nfb = _lib.sigma_host_apply(n, na, nb, tab, c)
    assert shape == idx.shape
    assert nfb == expect_fallback
    assert abs(out - ref).max() < 1e-11 * max(1.0, abs(ref).max())


@pytest.mark.parametrize("n_orb,na,nb", [(4, 2, 2), (5, 2, 2), (5, 3, 2), (6, 3, 3), (6, 2, 4), (7, 3, 3)])
def test_random_fragment_hamiltonian(n_orb, na, nb):
    h1, eri = random_fragment_integrals(n_orb, 1234)
    _check(n_orb, na, nb, fragment_term_table(h1, eri, 0, 0.0, 0.25))


def test_chemical_potential_and_no_penalty():
    h1, eri = random_fragment_integrals(5, 7)
    _check(5, 2, 2, fragment_term_table(h1, eri, 2, 0.3, 0.0))


def test_symmetric_molecule_hamiltonian():
    """H6 chain MO integrals: half of the X-mask groups vanish by inversion symmetry, so the
    alpha masks split into classes with different beta partner sets."""
    h_mo, eri_mo, e_nuc, e_rhf = mo_integrals(hydrogen_chain(6, 1.0), 6)
    _check(6, 3, 3, fragment_term_table(h_mo, eri_mo, 0, 0.0, 0.25))


def test_s2_operator():
    _check(5, 3, 2, s2_term_table(5))


def test_non_fermionic_terms_fall_back():
    """A Pauli operator that is no Jordan-Wigner image: its groups are left to the general kernel."""
    n_orb = 4
    h1, eri = random_fragment_integrals(n_orb, 5)
    x, z, c = fragment_term_table(h1, eri, 0, 0.0, 0.0)
    extra_x = np.array([0b00001111, 0b11000011], dtype=np.uint64)      # 4 flips with a wrong parity string
    extra_z = np.array([0b10100000, 0b00011000], dtype=np.uint64)
    tab = (np.concatenate([x, extra_x]), np.concatenate([z, extra_z]), np.concatenate([c, [0.3, -0.2]]))
    out, shape, nfb = _lib.sigma_host_apply(2 * n_orb, 2, 2, tab, None)
    assert nfb >= 2


@pytest.mark.parametrize("n_orb,na,nb", [(4, 2, 2), (6, 3, 3), (6, 2, 3)])
def test_opposite_spin_doubles_are_extracted(n_orb, na, nb):
    """The alpha-beta double excitations go to their own tables (csrc/mqc_sigma_ab.h): for a generic two-body
    Hamiltonian every X-mask group with two alpha and two beta flipped qubits is recognised (its weights pass the
    spectator verification), i.e. [n (n - 1) / 2]^2 groups with 4 ordered (ij, kl) weights each."""
    h1, eri = random_fragment_integrals(n_orb, 1234)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    n_terms, n_weights = _lib.sigma_ab_info(2 * n_orb, na, nb, tab)
    npairs = n_orb * (n_orb - 1) // 2
    assert n_weights == 4 * npairs * npairs
    x = np.asarray(tab[0], dtype=np.uint64)
    alpha = sum(1 << (2 * n_orb - 1 - 2 * p) for p in range(n_orb))
    beta = alpha >> 1
    pop = lambda v: np.array([bin(int(t)).count("1") for t in v])
    assert n_terms == int(((pop(x & np.uint64(alpha)) == 2) & (pop(x & np.uint64(beta)) == 2)).sum())
    # an operator that is NOT a number-conserving two-body operator keeps its terms on the generic path
    # (a Z on a spectator qubit that is not part of the Jordan-Wigner string of the two excitations)
    bad = (np.array([0b11110000], np.uint64), np.array([0b00000100], np.uint64), np.array([1.0]))
    assert _lib.sigma_ab_info(8, 2, 2, bad) == (0, 0)
    ok = (np.array([0b11110000], np.uint64), np.array([0], np.uint64), np.array([1.0]))          # XXXX on alpha0 beta0 alpha1 beta1
    assert _lib.sigma_ab_info(8, 2, 2, ok) == (1, 4)
    _check(4, 2, 2, ok)
    assert _lib.sigma_host_apply(8, 2, 2, bad)[2] == 1                  # ... and is reported as a fallback term
