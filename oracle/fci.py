"""ORACLE (test infrastructure, not product code) — exact fragment solver in the (N_e, S_z=0) sector.

Stand-in for `/root/reference/mqc/solver/dmet_solver_fci.py:8-120` (pyscf FCI), which has
the same boundary contract as the VQE solver (`dmet_solver_vqechem.py:17-88`): chemical
potential on the impurity diagonal (:22-25), ground state, 1-/2-RDM, impurity energy
(:79-85).  Built on the restated Hamiltonian table so that it also pins oracle/hamiltonian.py
and oracle/rdm.py against `test/testlog_fci` (E FCI, impurity energy, N_elec, totals).
"""
from __future__ import annotations

import numpy as np

from .hamiltonian import qubit_hamiltonian, term_table, _parity
from .rdm import make_strings_qubit, qubit_inverse, make_rdm12, impurity_energy


def sector_indices(norb, nelec, Sz=0):
    strs = make_strings_qubit(norb, nelec, Sz)
    return strs, np.array([qubit_inverse(norb, int(s)) for s in strs], dtype=np.int64)


def sector_hamiltonian(xm, zm, cf, n_qubits, idx):
    pos = -np.ones(1 << n_qubits, dtype=np.int64)
    pos[idx] = np.arange(len(idx))
    H = np.zeros((len(idx), len(idx)), dtype=np.complex128)
    u = idx.astype(np.uint64)
    for x, z, c in zip(xm, zm, cf):
        tgt = pos[(u ^ x).astype(np.int64)]
        m = tgt >= 0
        sg = 1.0 - 2.0 * _parity(u & z)
        np.add.at(H, (tgt[m], np.arange(len(idx))[m]), c * sg[m])
    return H


def solve_fci(oei, fock, tei, nelec, n_imp_orbs, chempot_imp=0.0, return_all=False):
    norb = fock.shape[0]
    nq = 2 * norb
    xz = qubit_hamiltonian(fock, tei, n_imp_orbs, chempot_imp, 0.0)
    xm, zm, cf = term_table(xz, nq, 1e-15)
    strs, idx = sector_indices(norb, nelec)
    H = sector_hamiltonian(xm, zm, cf, nq, idx)
    assert abs(H - H.conj().T).max() < 1e-12 and abs(H.imag).max() < 1e-12
    w, v = np.linalg.eigh(H.real)
    c0 = v[:, 0]
    rdm1, rdm2 = make_rdm12(c0, norb, strs)
    e_imp = impurity_energy(rdm1, rdm2, oei, fock, tei, n_imp_orbs)
    if return_all:
        psi = np.zeros(1 << nq, dtype=np.complex128)
        psi[idx] = c0
        return e_imp, rdm1, dict(e_fci=float(w[0]), rdm2=rdm2, psi=psi, spectrum=w)
    return e_imp, rdm1
