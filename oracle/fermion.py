"""ORACLE (test infrastructure, not product code) — symbolic fermion algebra.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  Parity status: the reference delegates this algebra to
OpenFermion (unpinned, not vendored, not importable here: `README.md:13-18`,
`mqc/solver/dmet_interface.py:2-3`), so this file restates OpenFermion's *published*
algorithms and is pinned indirectly, through the spectra / energies printed in the
reference's logs (tests/test_oracle_golden.py).

Restated pieces and their call sites in the reference:

* ``FermionOperator`` dict-of-terms with ``+=`` accumulation and ``terms.update``
  — used at `mqc/solver/dmet_interface.py:207-233`.
* ``normal_ordered``  — `dmet_interface.py:234-235,241`.  Algorithm: bubble each term into
  "creators left, within a type descending index"; a swap of different types flips
  the sign and, for equal indices, spawns the contracted term; equal operators
  adjacent => zero.
* ``jordan_wigner`` (via ``vqechem.tools.mapping(..., 'JW', ...)``, `dmet_interface.py:113,
  118,122`): a_p = 1/2 (X_p + iY_p) Z_0..Z_{p-1},  a_p^+ = 1/2 (X_p - iY_p) Z_0..Z_{p-1}.
* ``s2_operator(ncas)`` (vqechem.tools, `dmet_interface.py:112`): S^2 = S_-S_+ + S_z(S_z+1)
  over interleaved spin orbitals (alpha = 2p, beta = 2p+1).

Pauli strings are kept as (xmask, zmask) over *qubit numbers* (bit q <-> qubit q) with
the operator defined as  prod_q X_q^{x_q} Z_q^{z_q}  (X to the left of Z on each qubit), so
Y_q = i X_q Z_q.  ``to_label_form`` converts to OpenFermion's string keys
(((q,'X'|'Y'|'Z'),...)) and coefficients: coeff_label = coeff_xz * (-i)^{popc(x&z)} ... see
the function.
"""
from __future__ import annotations

from typing import Dict, Tuple

Term = Tuple[Tuple[int, int], ...]


class FermionOperator:
    """Minimal dict-of-terms operator: {((index, action), ...): coefficient}; action 1 = creation."""

    def __init__(self, term: Term = None, coefficient=1.0):
        self.terms: Dict[Term, complex] = {}
        if term is not None:
            self.terms[tuple(term)] = coefficient

    def __iadd__(self, other: "FermionOperator"):
        for t, c in other.terms.items():
            if t in self.terms:
                s = self.terms[t] + c
                if s == 0:                      # OpenFermion drops exact zeros on +=
                    del self.terms[t]
                else:
                    self.terms[t] = s
            else:
                self.terms[t] = c
        return self

    def __add__(self, other):
        out = FermionOperator()
        out.terms = dict(self.terms)
        out += other
        return out

    def __mul__(self, scalar):
        out = FermionOperator()
        out.terms = {t: c * scalar for t, c in self.terms.items()}
        return out

    __rmul__ = __mul__

    def product(self, other: "FermionOperator"):
        out = FermionOperator()
        for t1, c1 in self.terms.items():
            for t2, c2 in other.terms.items():
                out += FermionOperator(t1 + t2, c1 * c2)
        return out


def _normal_ordered_term(term, coefficient, out: FermionOperator):
    term = list(term)
    for i in range(1, len(term)):
        for j in range(i, 0, -1):
            right = term[j]
            left = term[j - 1]
            if right[1] and not left[1]:
                term[j - 1], term[j] = right, left
                coefficient = -coefficient
                if right[0] == left[0]:
                    _normal_ordered_term(tuple(term[:j - 1] + term[j + 1:]), -coefficient, out)
            elif right[1] == left[1]:
                if right[0] == left[0]:
                    return
                if right[0] > left[0]:
                    term[j - 1], term[j] = right, left
                    coefficient = -coefficient
    out += FermionOperator(tuple(term), coefficient)


def normal_ordered(op: FermionOperator) -> FermionOperator:
    out = FermionOperator()
    for t, c in op.terms.items():
        _normal_ordered_term(t, c, out)
    return out


# ----------------------------------------------------------------------------------------
# Pauli algebra in (xmask, zmask) form
# ----------------------------------------------------------------------------------------
def _popc(v: int) -> int:
    return bin(v).count("1")


def pauli_mul(a, b):
    """(X^x1 Z^z1)(X^x2 Z^z2) = (-1)^{popc(z1&x2)} X^{x1^x2} Z^{z1^z2}."""
    (x1, z1), (x2, z2) = a, b
    sign = -1.0 if (_popc(z1 & x2) & 1) else 1.0
    return (x1 ^ x2, z1 ^ z2), sign


def _ladder_xz(p: int, action: int):
    """a_p / a_p^+ as two (xmask,zmask) terms.
    a_p   = 1/2 (X + iY) Z-chain = 1/2 X(1 - Z) chain   -> +1/2 X chain , -1/2 XZ chain
    a_p^+ = 1/2 (X - iY) Z-chain = 1/2 X(1 + Z) chain   -> +1/2 X chain , +1/2 XZ chain
    (Y = i X Z.)"""
    chain = (1 << p) - 1
    x = 1 << p
    s = 0.5 if action else -0.5
    return [((x, chain), 0.5), ((x, chain | x), s)]


def jordan_wigner(op: FermionOperator) -> Dict[Tuple[int, int], complex]:
    """FermionOperator -> {(xmask, zmask): coeff} in X^x Z^z convention (qubit-number bits)."""
    out: Dict[Tuple[int, int], complex] = {}
    for term, coeff in op.terms.items():
        partial = [((0, 0), complex(coeff))]
        for (p, action) in term:
            lad = _ladder_xz(p, action)
            nxt = []
            for key, c in partial:
                for lkey, lc in lad:
                    k, sgn = pauli_mul(key, lkey)
                    nxt.append((k, c * lc * sgn))
            partial = nxt
        for k, c in partial:
            out[k] = out.get(k, 0.0) + c
    return out


def to_label_form(xz_terms, tol=0.0):
    """{(x,z): c} -> OpenFermion QubitOperator-style {((q,'X'),...): coeff}.
    X^1 Z^1 on a qubit is -iY, so the label coefficient is c * (-i)^{popc(x&z)}."""
    out = {}
    for (x, z), c in xz_terms.items():
        ny = _popc(x & z)
        cl = c * ((-1j) ** ny)
        if abs(cl) <= tol:
            continue
        key = []
        q = 0
        m = x | z
        while m >> q:
            if (m >> q) & 1:
                xb, zb = (x >> q) & 1, (z >> q) & 1
                key.append((q, "Y" if (xb and zb) else ("X" if xb else "Z")))
            q += 1
        out[tuple(key)] = cl
    return out


def s2_operator(n_orb: int) -> FermionOperator:
    """S^2 = S_- S_+ + S_z (S_z + 1), alpha = even, beta = odd spin orbitals."""
    s_plus = FermionOperator()
    s_minus = FermionOperator()
    s_z = FermionOperator()
    for p in range(n_orb):
        a, b = 2 * p, 2 * p + 1
        s_plus += FermionOperator(((a, 1), (b, 0)), 1.0)
        s_minus += FermionOperator(((b, 1), (a, 0)), 1.0)
        s_z += FermionOperator(((a, 1), (a, 0)), 0.5)
        s_z += FermionOperator(((b, 1), (b, 0)), -0.5)
    s2 = s_minus.product(s_plus)
    s2 += s_z.product(s_z)
    s2 += s_z
    return normal_ordered(s2)
