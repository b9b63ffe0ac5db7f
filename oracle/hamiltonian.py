"""ORACLE (test infrastructure, not product code) — fragment Hamiltonian, literal restatement.

Follows `/root/reference/mqc/solver/dmet_interface.py`:

* ``get_1e_2e_integral``  :10-39   spatial -> interleaved spin-orbital integrals
* ``DMET.hamiltonian``    :200-241 second-quantised H incl. the chemical-potential rule
  (``dict.update`` REPLACES the diagonal one-body terms of the impurity orbitals, :226-233)
* ``DMET.__init__``       :112-123 S^2 penalty: ``_ham = H_bare + shift/2 * S^2`` (shift 0.5)
* ``DMET.fci``            :243-251 spectrum of ``_ham + 0.5 S^2`` (printed in every VQE log)

The symbolic algebra is oracle/fermion.py (OpenFermion restated).  Pure-Python loops:
use for n_orb <= 6; the product builds the same table in closed form
(multiscalequantumcomputing_b200/jw.py) and tests compare the two.
"""
from __future__ import annotations

import numpy as np

from .fermion import FermionOperator, normal_ordered, jordan_wigner, s2_operator


def get_1e_2e_integral(one_body_mo, two_body_mo):
    n = one_body_mo.shape[0]
    h1 = np.zeros((2 * n, 2 * n))
    h2 = np.zeros((2 * n,) * 4)
    for p in range(n):
        for q in range(n):
            h1[2 * p, 2 * q] = one_body_mo[p, q]
            h1[2 * p + 1, 2 * q + 1] = one_body_mo[p, q]
            for r in range(n):
                for s in range(n):
                    v = two_body_mo[p, s, q, r]
                    h2[2 * p, 2 * q, 2 * r, 2 * s] = v
                    h2[2 * p + 1, 2 * q + 1, 2 * r + 1, 2 * s + 1] = v
                    h2[2 * p + 1, 2 * q, 2 * r, 2 * s + 1] = v
                    h2[2 * p, 2 * q + 1, 2 * r + 1, 2 * s] = v
    return h1, h2


def fermion_hamiltonian(one_body_mo, two_body_mo, n_imp_orbs, chempot_imp=0.0) -> FermionOperator:
    h1, h2 = get_1e_2e_integral(one_body_mo, two_body_mo)
    m = h1.shape[0]
    op1 = FermionOperator()
    op2 = FermionOperator()
    for p in range(m):
        for q in range(m):
            op1 += FermionOperator(((p, 1), (q, 0)), h1[p, q])
    for p in range(m):
        for q in range(m):
            for r in range(m):
                for s in range(m):
                    op2 += FermionOperator(((p, 1), (q, 1), (r, 0), (s, 0)), h2[p, q, r, s] * 0.5)
    new = FermionOperator()
    if chempot_imp != 0.0:
        for p in range(n_imp_orbs * 2):
            new += FermionOperator(((p, 1), (p, 0)), h1[p, p] - chempot_imp)
    op1.terms.update(new.terms)
    op1 = normal_ordered(op1)
    op2 = normal_ordered(op2)
    return normal_ordered(op1 + op2)


def qubit_hamiltonian(one_body_mo, two_body_mo, n_imp_orbs, chempot_imp=0.0, s2_shift=0.0):
    """{(xmask, zmask): coeff} over qubit-number bits for  H_bare + s2_shift * S^2.
    The reference optimises with s2_shift = 0.5/2 (`dmet_interface.py:80,121`)."""
    ham = fermion_hamiltonian(one_body_mo, two_body_mo, n_imp_orbs, chempot_imp)
    if s2_shift != 0.0:
        ham = ham + s2_shift * s2_operator(one_body_mo.shape[0])
    return jordan_wigner(ham)


def qubit_s2(n_orb):
    return jordan_wigner(s2_operator(n_orb))


# ----------------------------------------------------------------------------------------
# index conventions
# ----------------------------------------------------------------------------------------
def qubit_to_index_mask(mask: int, n_qubits: int) -> int:
    """OpenFermion's get_sparse_operator puts qubit 0 in the most significant index bit:
    qubit q <-> index bit (n_qubits-1-q)  (pinned by `test/testlog_vqe:92-94`)."""
    out = 0
    for q in range(n_qubits):
        if (mask >> q) & 1:
            out |= 1 << (n_qubits - 1 - q)
    return out


def term_table(xz_terms, n_qubits, tol=0.0):
    """Sorted arrays (xmask, zmask, coeff) over INDEX bits, X^x Z^z convention:
       <i ^ x| P |i> = (-1)^{popc(i & z)}.   Sorted by (xmask, zmask)."""
    rows = []
    for (x, z), c in xz_terms.items():
        if abs(c) <= tol:
            continue
        rows.append((qubit_to_index_mask(x, n_qubits), qubit_to_index_mask(z, n_qubits), complex(c)))
    rows.sort(key=lambda r: (r[0], r[1]))
    xm = np.array([r[0] for r in rows], dtype=np.uint64)
    zm = np.array([r[1] for r in rows], dtype=np.uint64)
    cf = np.array([r[2] for r in rows], dtype=np.complex128)
    return xm, zm, cf


def dense_matrix(xm, zm, cf, n_qubits):
    """Dense 2^N x 2^N matrix of sum_t c_t X^x Z^z (index-bit masks)."""
    dim = 1 << n_qubits
    idx = np.arange(dim, dtype=np.uint64)
    H = np.zeros((dim, dim), dtype=np.complex128)
    for x, z, c in zip(xm, zm, cf):
        par = _parity(idx & z)
        H[(idx ^ x).astype(np.int64), idx.astype(np.int64)] += c * (1.0 - 2.0 * par)
    return H


def _parity(v):
    v = v.copy()
    for s in (32, 16, 8, 4, 2, 1):
        v ^= v >> np.uint64(s)
    return (v & np.uint64(1)).astype(np.float64)


def apply_terms(xm, zm, cf, psi):
    """H|psi> matrix-free."""
    dim = psi.shape[0]
    idx = np.arange(dim, dtype=np.uint64)
    out = np.zeros_like(psi, dtype=np.complex128)
    for x, z, c in zip(xm, zm, cf):
        par = _parity(idx & z)
        # out[i ^ x] += c * (-1)^{popc(i&z)} psi[i]
        out[(idx ^ x).astype(np.int64)] += c * (1.0 - 2.0 * par) * psi
    return out
