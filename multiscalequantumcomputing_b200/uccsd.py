"""UCCSD factor lists (host side).

The reference builds its ansatz inside VQEChem (`mqc/solver/dmet_solver_vqechem.py:54-65`:
``FermionOps`` pool + ``ADAPT``); BASELINE.json's north star fixes the ansatz to
Jordan-Wigner UCCSD in disentangled (factorised) form

    |psi(theta)> = prod_k exp(theta_k (T_k - T_k^+)) |ref>,      factor 0 acts first

with T_k = a_a^+ a_i (single) or a_a^+ a_b^+ a_j a_i (double) over spin orbitals
(alpha = even, beta = odd, `mqc/solver/dmet_interface.py:27-37`).  Factor order is part of
the ansatz (SURVEY.md §7 "hard parts"); two orders are provided:

``lex``      singles then doubles, lexicographic in (i,a) / (i,j,a,b)
``blocked``  the same factor SET, ordered block by block over (occupied-orbital subset,
             virtual-orbital subset) pairs so that consecutive factors move few distinct
             qubits - this is what lets the GPU sweep fuse ~50-100 factors per HBM pass.

A factor list is plain data: ``kind`` int32[n], ``idx`` int32[n,4] (single: i,a,-1,-1;
double: i,j,a,b), ``theta_index`` int32[n].
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import combinations

import numpy as np


@dataclass
class FactorList:
    n_qubits: int
    kind: np.ndarray         # int32[n]
    idx: np.ndarray          # int32[n,4]
    theta_index: np.ndarray  # int32[n]
    n_theta: int
    ref_index: int           # amplitude index of the reference determinant
    theta_scale: np.ndarray = None   # float64[n] angle of factor k = theta_scale[k] * theta[theta_index[k]] (None: 1)

    def __len__(self):
        return len(self.kind)

    def as_tuples(self):
        """(kind, indices): kind 1 (i, a), kind 2 (i, j, a, b), kind 3 (i, a, c) = n_c-dressed single."""
        out = []
        for k, ix in zip(self.kind, self.idx):
            nix = {1: 2, 2: 4, 3: 3}[int(k)]
            out.append((int(k), tuple(int(v) for v in ix[:nix])))
        return out


def electron_split(n_elec: int):
    """noa = N // 2, nob = N - noa (`mqc/solver/dmet_interface.py:84-85`): an odd electron count
    puts the extra electron into the beta string."""
    noa = n_elec // 2
    return noa, n_elec - noa


def occupied_qubits(n_elec: int):
    noa, nob = electron_split(n_elec)
    return sorted([2 * p for p in range(noa)] + [2 * p + 1 for p in range(nob)])


def reference_index(n_qubits: int, n_elec: int) -> int:
    """Lowest noa alpha and nob beta spin orbitals occupied; qubit 0 is the most significant index
    bit.  Even N: qubits 0..N-1 (`test/testlog_vqe:25`: index 240 for N=8, N_e=4)."""
    return sum(1 << (n_qubits - 1 - q) for q in occupied_qubits(n_elec))


def _all_excitations(n_orb: int, n_elec: int):
    nq = 2 * n_orb
    if n_elec < 0 or n_elec > nq or electron_split(n_elec)[1] > n_orb:
        raise ValueError("cannot place %d electrons in %d spatial orbitals" % (n_elec, n_orb))
    occ = occupied_qubits(n_elec)
    vir = [q for q in range(nq) if q not in set(occ)]
    singles = [(i, a) for i in occ for a in vir if i % 2 == a % 2]
    doubles = []
    for i, j in combinations(occ, 2):
        for a, b in combinations(vir, 2):
            if sorted((i % 2, j % 2)) == sorted((a % 2, b % 2)):
                doubles.append((i, j, a, b))
    return singles, doubles


def _pair_cover(items, size):
    """Greedy family of ``size``-subsets of ``items`` covering every pair (and every item)."""
    items = list(items)
    if len(items) <= size:
        return [tuple(items)]
    need = set(combinations(items, 2))
    fam = []
    while need:
        best, gain = None, -1
        for c in combinations(items, size):
            g = sum(1 for p in combinations(c, 2) if p in need)
            if g > gain:
                best, gain = c, g
        fam.append(best)
        need -= set(combinations(best, 2))
    return fam


def uccsd_factors(n_orb: int, n_elec: int, order: str = "blocked", block_orbs: int = 7) -> FactorList:
    """All spin-conserving singles and doubles from the closed-shell reference.

    ``block_orbs``: spatial orbitals per block (occupied + virtual); 7 spatial = 14 qubits
    matches the 14-bit compressed-sector tile of the sweep kernel (6 for 12-bit dense tiles)."""
    nq = 2 * n_orb
    singles, doubles = _all_excitations(n_orb, n_elec)
    if order == "lex":
        seq = [(1, s + (-1, -1)) for s in singles] + [(2, d) for d in doubles]
    elif order == "blocked":
        # spatial orbitals that hold an occupied / an empty spin orbital (they overlap in the singly
        # occupied orbital of an odd-electron reference)
        noa, nob = electron_split(n_elec)
        nocc, nvir = nob, n_orb - noa
        so = min(nocc, block_orbs // 2)
        sv = min(nvir, block_orbs - so)
        so = min(nocc, block_orbs - sv)
        occ_blocks = _pair_cover(range(nocc), max(so, 1)) if nocc else [()]
        vir_blocks = _pair_cover(range(noa, n_orb), max(sv, 1)) if nvir else [()]
        seq = []
        done = set()
        flip = False
        for ob in occ_blocks:
            vbs = vir_blocks[::-1] if flip else vir_blocks
            flip = not flip
            for vb in vbs:
                orbs = set(ob) | set(vb)
                for s in singles:
                    if s not in done and all((q // 2) in orbs for q in s):
                        done.add(s)
                        seq.append((1, s + (-1, -1)))
                for d in doubles:
                    if d not in done and all((q // 2) in orbs for q in d):
                        done.add(d)
                        seq.append((2, d))
        assert len(done) == len(singles) + len(doubles)
    else:
        raise ValueError("unknown order %r" % order)
    kind = np.array([k for k, _ in seq], dtype=np.int32)
    idx = np.array([ix for _, ix in seq], dtype=np.int32).reshape(-1, 4)
    theta_index = np.arange(len(seq), dtype=np.int32)
    return FactorList(nq, kind, idx, theta_index, len(seq), reference_index(nq, n_elec))
