"""Closed-form JW term table (product) against the literal OpenFermion-style restatement."""
import numpy as np
import pytest

from multiscalequantumcomputing_b200.jw import fragment_term_table, s2_term_table, number_of_x_groups
from multiscalequantumcomputing_b200.synth.hydrogen import random_fragment_integrals
from oracle.fermion import to_label_form, jordan_wigner, FermionOperator
from oracle.hamiltonian import qubit_hamiltonian, qubit_s2, term_table


@pytest.mark.parametrize("n,mu,s2", [(2, 0.0, 0.0), (4, 0.0, 0.0), (4, 0.013, 0.25), (5, -0.2, 0.25), (6, 0.01, 0.25)])
def test_table_matches_oracle(n, mu, s2):
    h1, eri = random_fragment_integrals(n, 99 + n)
    xm, zm, cf = term_table(qubit_hamiltonian(h1, eri, n // 2, mu, s2), 2 * n, 1e-14)
    x2, z2, c2 = fragment_term_table(h1, eri, n // 2, mu, s2)
    assert len(xm) == len(x2)
    assert (xm == x2).all() and (zm == z2).all()          # identical Pauli-string key sets
    # coefficients: a few ulp of the largest contribution (summation order differs)
    assert abs(cf - c2).max() <= 64 * np.finfo(float).eps * max(1.0, abs(cf).max())
    assert abs(c2.imag).max() < 1e-15                       # real Hamiltonian


@pytest.mark.parametrize("n,terms,groups", [(4, 361, 51), (8, 5793, 981)])
def test_generic_counts(n, terms, groups):
    """1.5n^4 - n^3 + 2.5n^2 + 1 terms in 1 + n(n-1) + 2C(n,4) + C(n,2)^2 groups."""
    h1, eri = random_fragment_integrals(n, 1234)
    x, z, c = fragment_term_table(h1, eri)
    assert len(x) == terms and number_of_x_groups(x) == groups


def test_s2_table():
    xm, zm, cf = term_table(qubit_s2(3), 6, 1e-14)
    x2, z2, c2 = s2_term_table(3)
    assert (xm == x2).all() and (zm == z2).all() and abs(cf - c2).max() < 1e-15


def test_label_form_of_number_operator():
    # a_1^+ a_1 = (1 - Z_1)/2 in OpenFermion labels
    lab = to_label_form(jordan_wigner(FermionOperator(((1, 1), (1, 0)), 1.0)))
    assert abs(lab[()] - 0.5) < 1e-15 and abs(lab[((1, "Z"),)] + 0.5) < 1e-15


def test_label_form_hopping():
    # a_0^+ a_1 + h.c. = (X0 X1 + Y0 Y1)/2
    op = FermionOperator(((0, 1), (1, 0)), 1.0) + FermionOperator(((1, 1), (0, 0)), 1.0)
    lab = to_label_form(jordan_wigner(op), 1e-15)
    assert set(lab) == {((0, "X"), (1, "X")), ((0, "Y"), (1, "Y"))}
    assert all(abs(v - 0.5) < 1e-15 for v in lab.values())
