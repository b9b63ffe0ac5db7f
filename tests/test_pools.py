"""Operator pools (multiscalequantumcomputing_b200/pools.py): the reference's spin-adapted fermionic pool
(`mqc/solver/dmet_solver_vqechem.py:38` 'spin_sym': 'sa'; "Number of Operator: 66", `test/testlog_vqe:23-24`)."""
import numpy as np
import pytest

from multiscalequantumcomputing_b200.jw import s2_term_table
from multiscalequantumcomputing_b200.pools import generalized_spin_orbital_pool, singlet_gsd_pool
from oracle.fermion import FermionOperator, jordan_wigner
from oracle.hamiltonian import dense_matrix, term_table
from oracle.statevector import excitation_ops


def _matrix(op, n):
    kind, idx, c = op
    A = np.zeros((1 << (2 * n), 1 << (2 * n)), dtype=np.complex128)
    for k, ix, cm in zip(kind, idx, c):
        T = excitation_ops(int(k), tuple(int(v) for v in ix[:{1: 2, 2: 4, 3: 3}[int(k)]]))
        Td = tuple((p, 1 - a) for (p, a) in reversed(T))
        A += cm * dense_matrix(*term_table(jordan_wigner(FermionOperator(T, 1.0) + FermionOperator(Td, -1.0)), 2 * n), 2 * n)
    return A


def test_pool_sizes_match_the_reference_log():
    assert len(singlet_gsd_pool(4)) == 66                      # test/testlog_vqe:24
    assert len(singlet_gsd_pool(2)) == 4 and len(singlet_gsd_pool(3)) == 21
    assert len(generalized_spin_orbital_pool(4)) == 90


@pytest.mark.parametrize("n", [2, 3])
def test_pool_operators_are_spin_adapted_and_normalised(n):
    """Every operator is anti-Hermitian, commutes with S^2, S_z and N (it cannot break the spin symmetry of the
    reference determinant) and is normalised like the pool of the ADAPT-VQE paper (sum of squared normal-ordered
    coefficients = 1); the operators are linearly independent."""
    S2 = dense_matrix(*s2_term_table(n), 2 * n)
    nq = 2 * n
    occ = np.array([[(i >> (nq - 1 - q)) & 1 for q in range(nq)] for i in range(1 << nq)], dtype=float)
    Nop = np.diag(occ.sum(1)); Sz = np.diag(0.5 * (occ[:, 0::2].sum(1) - occ[:, 1::2].sum(1)))
    mats = []
    for op in singlet_gsd_pool(n):
        A = _matrix(op, n)
        assert abs(A + A.conj().T).max() < 1e-12
        for Q in (S2, Nop, Sz):
            assert abs(A @ Q - Q @ A).max() < 1e-12
        # a primitive tau = T - T^+ contributes two normal-ordered terms
        assert abs(2.0 * np.sum(op[2] ** 2) - 1.0) < 1e-12
        mats.append(A.real.ravel())
    assert np.linalg.matrix_rank(np.array(mats), tol=1e-9) == len(mats)
