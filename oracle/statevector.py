"""ORACLE (test infrastructure, not product code) — UCCSD state vector on the CPU, numpy.

PARITY UNPINNED by the reference for the UCCSD numbers themselves: the reference runs
ADAPT-VQE inside VQEChem (not in the tree; `mqc/solver/dmet_solver_vqechem.py:35-41,65-71`),
never a fixed UCCSD, so no log holds a UCCSD energy or gradient.  What *is* pinned:
the state-vector layout (qubit 0 = most significant index bit, `test/testlog_vqe:56-94`),
the Hamiltonian (oracle/hamiltonian.py vs the logged spectra) and the generator algebra
(``excitation_generator_matrix`` below is built from the same restated OpenFermion JW map
and ``apply_excitation`` is tested against its matrix exponential).

What this restates (`mqc/solver/dmet_interface.py:279-302`, inner boundary):
    ansatz.state(theta)    |psi> = prod_k exp(theta_k A_k) |ref>        (factor 0 acts first)
    ansatz.energy(theta)   <psi|H|psi>
    ansatz.gradient(theta) dE/dtheta_k  (adjoint / reverse sweep, exact)

A factor is ``(kind, (i, a))`` with A = a_a^+ a_i - h.c. (kind 1) or
``(kind, (i, j, a, b))`` with A = a_a^+ a_b^+ a_j a_i - h.c. (kind 2); indices are spin
orbitals = qubits; several factors may share one theta through ``theta_index``.
"""
from __future__ import annotations

import numpy as np

from .fermion import FermionOperator, jordan_wigner
from .hamiltonian import term_table, dense_matrix, apply_terms


def bit_reverse_table(n_qubits):
    """index(D) for a determinant bit string D (bit P <-> qubit P): reverse N bits
    (`dmet_solver_vqechem.py:90-98` qubit_inverse)."""
    d = np.arange(1 << n_qubits, dtype=np.int64)
    out = np.zeros_like(d)
    for q in range(n_qubits):
        out |= ((d >> q) & 1) << (n_qubits - 1 - q)
    return out


def reference_index(n_qubits, n_elec):
    """`reference(imp, options)`: qubits 0..n_elec-1 set (index 240 at N=8, N_e=4;
    `test/testlog_vqe:25`)."""
    return (1 << n_qubits) - (1 << (n_qubits - n_elec))


def _popc_arr(v):
    v = v.astype(np.uint64)
    c = np.zeros(v.shape, dtype=np.int64)
    while True:
        nz = v != 0
        if not nz.any():
            break
        c += (v & np.uint64(1)).astype(np.int64)
        v = v >> np.uint64(1)
    return c


def _apply_ladder_sequence(D, ops):
    """Apply a product of ladder operators (rightmost first) to every determinant in the
    int64 array D (qubit-number bits).  Returns (D_out, sign) with sign 0 where annihilated."""
    cur = D.copy()
    sign = np.ones(D.shape, dtype=np.int64)
    for (p, action) in reversed(ops):
        bit = np.int64(1) << np.int64(p)
        occ = (cur & bit) != 0
        ok = ~occ if action else occ
        below = _popc_arr(cur & (bit - 1))
        sign = np.where(ok, sign * (1 - 2 * (below & 1)), 0)
        cur = cur ^ bit
    return cur, sign


def excitation_ops(kind, idx):
    """T of the factor exp(theta (T - T^+)): kind 1 a_a^+ a_i, kind 2 a_a^+ a_b^+ a_j a_i, kind 3 the number-dressed
    single n_c a_a^+ a_i = a_c^+ a_c a_a^+ a_i (idx = i, a, c)."""
    if kind == 1:
        i, a = idx[0], idx[1]
        return ((a, 1), (i, 0))
    if kind == 3:
        i, a, c = idx[0], idx[1], idx[2]
        return ((c, 1), (c, 0), (a, 1), (i, 0))
    i, j, a, b = idx
    return ((a, 1), (b, 1), (j, 0), (i, 0))


def apply_generator(kind, idx, psi, n_qubits, rev=None):
    """A|psi> with A = T - T^+ (literal ladder-operator action)."""
    if rev is None:
        rev = bit_reverse_table(n_qubits)
    D = np.arange(1 << n_qubits, dtype=np.int64)
    T = excitation_ops(kind, idx)
    Tdag = tuple((p, 1 - act) for (p, act) in reversed(T))
    out = np.zeros_like(psi)
    for ops, sgn in ((T, 1.0), (Tdag, -1.0)):
        D2, s = _apply_ladder_sequence(D, ops)
        m = s != 0
        np.add.at(out, rev[D2[m]], sgn * s[m] * psi[rev[D[m]]])
    return out


def apply_excitation(kind, idx, theta, psi, n_qubits, rev=None):
    """exp(theta (T - T^+)) |psi>:  on every pair (D, D' = T D) a Givens rotation
        psi'[D]  = cos(theta) psi[D]  - s sin(theta) psi[D']
        psi'[D'] = cos(theta) psi[D'] + s sin(theta) psi[D]          s = sign of T on D."""
    if rev is None:
        rev = bit_reverse_table(n_qubits)
    D = np.arange(1 << n_qubits, dtype=np.int64)
    D2, s = _apply_ladder_sequence(D, excitation_ops(kind, idx))
    m = s != 0
    lo = rev[D[m]]
    hi = rev[D2[m]]
    sg = s[m].astype(np.float64)
    c, sn = np.cos(theta), np.sin(theta)
    out = psi.copy()
    a, b = psi[lo], psi[hi]
    out[lo] = c * a - sg * sn * b
    out[hi] = c * b + sg * sn * a
    return out


def excitation_generator_matrix(kind, idx, n_qubits):
    """Dense matrix of A = T - T^+ through the restated JW map (independent of the
    ladder-action code above; used to cross-check it)."""
    T = excitation_ops(kind, idx)
    Tdag = tuple((p, 1 - act) for (p, act) in reversed(T))
    op = FermionOperator(T, 1.0) + FermionOperator(Tdag, -1.0)
    xm, zm, cf = term_table(jordan_wigner(op), n_qubits)
    return dense_matrix(xm, zm, cf, n_qubits)


class UCCSDOracle:
    """Dense-vector CPU restatement of the VQEChem ansatz object for a fixed factor list."""

    def __init__(self, n_qubits, factors, theta_index, ref_index, ham_terms, theta_scale=None):
        """theta_scale[k]: factor k rotates by theta_scale[k] * theta[theta_index[k]] (None: 1)."""
        self.n = n_qubits
        self.factors = list(factors)
        self.theta_index = np.asarray(theta_index, dtype=np.int64)
        self.n_theta = int(self.theta_index.max()) + 1 if len(self.theta_index) else 0
        self.ref_index = int(ref_index)
        self.xm, self.zm, self.cf = ham_terms
        self.rev = bit_reverse_table(n_qubits)
        self.scale = np.ones(len(self.factors)) if theta_scale is None else np.asarray(theta_scale, dtype=np.float64)

    def state(self, theta):
        psi = np.zeros(1 << self.n, dtype=np.complex128)
        psi[self.ref_index] = 1.0
        for k, ((kind, idx), ti) in enumerate(zip(self.factors, self.theta_index)):
            psi = apply_excitation(kind, idx, self.scale[k] * theta[ti], psi, self.n, self.rev)
        return psi

    def h_apply(self, psi):
        return apply_terms(self.xm, self.zm, self.cf, psi)

    def energy(self, theta):
        psi = self.state(theta)
        return float(np.vdot(psi, self.h_apply(psi)).real)

    def energy_grad(self, theta):
        psi = self.state(theta)
        lam = self.h_apply(psi)
        e = float(np.vdot(psi, lam).real)
        g = np.zeros(self.n_theta)
        for k in range(len(self.factors) - 1, -1, -1):
            kind, idx = self.factors[k]
            ti = self.theta_index[k]
            g[ti] += self.scale[k] * 2.0 * np.vdot(lam, apply_generator(kind, idx, psi, self.n, self.rev)).real
            psi = apply_excitation(kind, idx, -self.scale[k] * theta[ti], psi, self.n, self.rev)
            lam = apply_excitation(kind, idx, -self.scale[k] * theta[ti], lam, self.n, self.rev)
        return e, g

    def gradient(self, theta):
        return self.energy_grad(theta)[1]


# ----------------------------------------------------------------------------------------
# "reference-style" sparse path: what OpenFermion + SciPy do inside the reference
# (`dmet_interface.py:114,119,123` get_sparse_operator; VQEChem expm_multiply chain)
# ----------------------------------------------------------------------------------------
def sparse_operator(xm, zm, cf, n_qubits):
    import scipy.sparse as sp
    dim = 1 << n_qubits
    idx = np.arange(dim, dtype=np.uint64)
    rows, cols, vals = [], [], []
    from .hamiltonian import _parity
    for x, z, c in zip(xm, zm, cf):
        par = _parity(idx & z)
        rows.append((idx ^ x).astype(np.int64))
        cols.append(idx.astype(np.int64))
        vals.append(c * (1.0 - 2.0 * par))
    M = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(dim, dim)).tocsc()
    M.sum_duplicates()
    return M


class SparseReferenceStyle:
    """energy+gradient the way the reference's dependency stack does it: csc Hamiltonian,
    sparse anti-Hermitian generators, scipy ``expm_multiply`` chain, reverse sweep."""

    def __init__(self, n_qubits, factors, theta_index, ref_index, ham_terms):
        from scipy.sparse.linalg import expm_multiply  # noqa: F401
        self.n = n_qubits
        self.factors = list(factors)
        self.theta_index = np.asarray(theta_index, dtype=np.int64)
        self.n_theta = int(self.theta_index.max()) + 1
        self.ref_index = int(ref_index)
        self.H = sparse_operator(*ham_terms, n_qubits)
        self.A = []
        for kind, idx in self.factors:
            T = excitation_ops(kind, idx)
            Tdag = tuple((p, 1 - act) for (p, act) in reversed(T))
            op = FermionOperator(T, 1.0) + FermionOperator(Tdag, -1.0)
            self.A.append(sparse_operator(*term_table(jordan_wigner(op), n_qubits), n_qubits))

    def energy_grad(self, theta):
        from scipy.sparse.linalg import expm_multiply
        psi = np.zeros(1 << self.n, dtype=np.complex128)
        psi[self.ref_index] = 1.0
        for A, ti in zip(self.A, self.theta_index):
            psi = expm_multiply(theta[ti] * A, psi)
        lam = self.H @ psi
        e = float(np.vdot(psi, lam).real)
        g = np.zeros(self.n_theta)
        for k in range(len(self.A) - 1, -1, -1):
            ti = self.theta_index[k]
            g[ti] += 2.0 * np.vdot(lam, self.A[k] @ psi).real
            psi = expm_multiply(-theta[ti] * self.A[k], psi)
            lam = expm_multiply(-theta[ti] * self.A[k], lam)
        return e, g
