"""ORACLE (test infrastructure, not product code) — determinant-bitstring RDMs.

Restates `/root/reference/mqc/tools/rdm.py` in float64 (the reference accumulates in
complex64 and returns float32 — a defect, `rdm.py:133,137,165,171`):

* ``make_strings_qubit``          rdm.py:18-30   (pyscf ``cistring.make_strings`` = all
                                   k-subsets as bit masks in ascending integer order)
* ``cre_des_sign``                pyscf.fci.cistring (0 if p occupied or q empty, else
                                   (-1)^{#occupied strictly between p and q})
* ``gen_linkstr_index_qubit``     rdm.py:91-129
* ``make_rdm1``                   rdm.py:131-138
* ``make_rdm12`` + ``reorder_rdm``  rdm.py:157-172; pyscf ``fci_slow.reorder_rdm``:
                                   ``rdm2[:,k,k,:] -= rdm1.T``
* ``get_rdm12`` gather            `mqc/solver/dmet_solver_vqechem.py:90-116`
* impurity energy                 `dmet_solver_vqechem.py:75-81`

Pinned by tests/test_oracle_golden.py against `test/testlog_vqe:92-185`
(strings, gathered fcivec, rdm1, rdm2 dumps).
"""
from __future__ import annotations

from itertools import combinations

import numpy as np


def make_strings(orb_list, nelec):
    orb_list = list(orb_list)
    out = []
    for comb in combinations(orb_list, int(nelec)):
        v = 0
        for o in comb:
            v |= 1 << o
        out.append(v)
    out.sort()
    return out


def make_strings_qubit(norb, nelec, Sz=0):
    nqubit = 2 * norb
    orbs = range(nqubit)
    na = (nelec + Sz) // 2
    nb = nelec - na
    stra = make_strings(orbs[::2], na)
    strb = make_strings(orbs[1::2], nb)
    return np.asarray([i + j for i in stra for j in strb], dtype=np.int64)


def cre_des_sign(p, q, string0):
    if p == q:
        return 1
    if (string0 & (1 << p)) or (not (string0 & (1 << q))):
        return 0
    if p > q:
        mask = (1 << p) - (1 << (q + 1))
    else:
        mask = (1 << q) - (1 << (p + 1))
    return (-1) ** bin(string0 & mask).count("1")


def gen_linkstr_index_qubit(norb, strs):
    nqubit = 2 * norb
    strdic = {int(s): k for k, s in enumerate(strs)}
    table = []
    for str0 in (int(s) for s in strs):
        occ = [i for i in range(nqubit) if str0 & (1 << i)]
        vir = [i for i in range(nqubit) if not str0 & (1 << i)]
        links = [(i, i, strdic[str0], 1) for i in occ]
        for i in occ:
            for a in vir:
                if i % 2 == a % 2:
                    str1 = str0 ^ (1 << i) | (1 << a)
                    if str1 in strdic:
                        links.append((a, i, strdic[str1], cre_des_sign(a, i, str0)))
        table.append(links)
    return table


def make_rdm1(fcivec, norb, strs, link_index=None):
    if link_index is None:
        link_index = gen_linkstr_index_qubit(norb, strs)
    rdm1 = np.zeros((norb, norb), dtype=np.complex128)
    for str0, tab in enumerate(link_index):
        for a, i, str1, sign in tab:
            rdm1[a // 2, i // 2] += sign * np.conj(fcivec[str0]) * fcivec[str1]
    return rdm1.real.copy()


def make_rdm12(fcivec, norb, strs):
    link_index = gen_linkstr_index_qubit(norb, strs)
    rdm1 = make_rdm1(fcivec, norb, strs, link_index)
    rdm2 = np.zeros((norb,) * 4, dtype=np.complex128)
    for str0, tab in enumerate(link_index):
        for r, s, str1, sign1 in tab:
            for p, q, str2, sign2 in link_index[str1]:
                rdm2[p // 2, q // 2, r // 2, s // 2] += sign1 * sign2 * np.conj(fcivec[str0]) * fcivec[str2]
    rdm2 = rdm2.real.copy()
    for k in range(norb):
        rdm2[:, k, k, :] -= rdm1.T
    return rdm1, rdm2


def qubit_inverse(norb, idx):
    n = 2 * norb
    out = 0
    for q in range(n):
        if (idx >> q) & 1:
            out |= 1 << (n - 1 - q)
    return out


def get_rdm12(wfn, nelec, norb, Sz=0):
    strs = make_strings_qubit(norb, nelec, Sz)
    fcivec = np.array([wfn[qubit_inverse(norb, int(s))] for s in strs])
    return make_rdm12(fcivec, norb, strs)


def impurity_energy(rdm1, rdm2, oei, fock, tei, n_imp_orbs):
    e = 0.5 * np.einsum("ij,ij->", rdm1[:n_imp_orbs, :], oei[:n_imp_orbs, :] + fock[:n_imp_orbs, :])
    e += 0.5 * np.einsum("ijkl,ijkl->", rdm2[:n_imp_orbs], tei[:n_imp_orbs])
    return float(e)


# ----------------------------------------------------------------------------------------
# independent cross-check: RDMs from literal ladder-operator action on the full vector
# ----------------------------------------------------------------------------------------
def rdm12_bruteforce(psi, norb):
    """dm1[p,q] = sum_s <q_s^+ p_s>... in pyscf convention:
       dm1[p,q] = sum_sigma <p_sigma^+ q_sigma>   (symmetric for real states)
       dm2[p,q,r,s] = sum_{sigma,tau} <p_sigma^+ r_tau^+ s_tau q_sigma>."""
    from .statevector import _apply_ladder_sequence, bit_reverse_table
    n = 2 * norb
    rev = bit_reverse_table(n)
    D = np.arange(1 << n, dtype=np.int64)
    vec = psi[rev]                       # indexed by determinant string D

    def expect(ops):
        D2, s = _apply_ladder_sequence(D, ops)
        m = s != 0
        return np.sum(np.conj(vec[D2[m]]) * s[m] * vec[D[m]])

    dm1 = np.zeros((norb, norb))
    dm2 = np.zeros((norb,) * 4)
    for p in range(norb):
        for q in range(norb):
            for sg in (0, 1):
                dm1[p, q] += expect(((2 * p + sg, 1), (2 * q + sg, 0))).real
            for r in range(norb):
                for s in range(norb):
                    for sg in (0, 1):
                        for tu in (0, 1):
                            dm2[p, q, r, s] += expect(((2 * p + sg, 1), (2 * r + tu, 1),
                                                       (2 * s + tu, 0), (2 * q + sg, 0))).real
    return dm1, dm2
