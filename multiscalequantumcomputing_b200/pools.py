"""Operator pools of the ADAPT-VQE solver (host side).

The reference's default pool is VQEChem's spin-adapted fermionic pool
(`/root/reference/mqc/solver/dmet_solver_vqechem.py:35-41`: ``'ops': {'class': 'fermionic', 'spin_sym': 'sa'}``,
`:53-54` ``FermionOps(imp, options['ops'])``; `test/testlog_vqe:23-24` logs "Number of Operator: 66" for the
four-orbital fragment).  VQEChem is not in the tree; 66 = 6 + 60 is the size of the spin-adapted generalized
singles-and-doubles pool of the ADAPT-VQE paper (Grimsley, Economou, Barnes, Mayhall, Nat. Commun. 10, 3007,
"singlet GSD"), restated here from its published definition:

    singles   p < q:                 E_pq - E_qp,              E_pq = a^+_{p a} a_{q a} + a^+_{p b} a_{q b}
    doubles   pairs (p <= q) <= (r <= s) in pair order, two spin couplings A ("triplet-triplet") and B
              ("singlet-singlet") of   a^+_r a_p a^+_s a_q   over the spin labels, each minus its adjoint,

normal ordered, normalised, zero operators dropped.  A pool operator is a LINEAR COMBINATION
A = sum_m c_m tau_m of primitive anti-Hermitian excitations tau = T - T^+ with

    kind 1  T = a^+_a a_i                  kind 2  T = a^+_a a^+_b a_j a_i
    kind 3  T = n_c a^+_a a_i  (left over by normal ordering when an index is created and annihilated)

which is what the sweep kernels exponentiate; inside the factorised ansatz the operator enters in product
form prod_m exp(theta c_m tau_m) with ONE parameter (`mqc_set_ansatz_scaled`), and its pool gradient
dE/dtheta at theta = 0 is the exact sum_m c_m <[H, tau_m]> whatever the order of the factors.
"""
from __future__ import annotations

from itertools import combinations

import numpy as np


# ---- a minimal fermion algebra: {((q, action), ...): coeff}, action 1 = creation --------------------------------
def _normal_order_term(ops, coeff, out):
    """Add coeff * (product of ladder operators `ops`) to `out` in canonical order: creation operators first,
    both groups with descending spin-orbital index."""
    ops = list(ops)
    for k in range(len(ops) - 1):
        (p, ap), (q, aq) = ops[k], ops[k + 1]
        if ap == 0 and aq == 1:                      # a_p a_q^+ = delta_pq - a_q^+ a_p
            rest = ops[:k] + ops[k + 2:]
            if p == q:
                _normal_order_term(rest, coeff, out)
            _normal_order_term(ops[:k] + [ops[k + 1], ops[k]] + ops[k + 2:], -coeff, out)
            return
        if ap == aq:
            if p == q:
                return                               # a a = a^+ a^+ = 0
            if p < q:                                # same type: anticommute into descending order
                _normal_order_term(ops[:k] + [ops[k + 1], ops[k]] + ops[k + 2:], -coeff, out)
                return
    key = tuple(ops)
    out[key] = out.get(key, 0.0) + coeff


def _anti_hermitian_normal_ordered(terms):
    """X - X^+ in canonical order for X = sum coeff * product (real coefficients)."""
    out = {}
    for ops, c in terms:
        _normal_order_term(ops, c, out)
        dag = tuple((q, 1 - a) for (q, a) in reversed(ops))
        _normal_order_term(dag, -c, out)
    return {k: v for k, v in out.items() if abs(v) > 1e-12}


def _primitives(A):
    """Canonical anti-Hermitian operator -> {(kind, idx4): coeff} over the factor primitives."""
    prim = {}

    def add(kind, idx, c):
        key = (kind, idx)
        prim[key] = prim.get(key, 0.0) + c

    for ops, c in A.items():
        if len(ops) == 2:
            (w, _), (z, _) = ops                     # a_w^+ a_z
            if w == z:
                raise ValueError("number operator in an anti-Hermitian generator")
            if w > z:                                # representative: excitation from the lower to the higher index
                add(1, (z, w, -1, -1), c)
        elif len(ops) == 4:
            (w, _), (x, _), (y, _), (z, _) = ops     # a_w^+ a_x^+ a_y a_z, w > x, y > z
            common = {w, x} & {y, z}
            if len(common) == 0:
                if (w, x) > (y, z):
                    add(2, (z, y, w, x), c)          # T = a_a^+ a_b^+ a_j a_i with (i, j, a, b) = (z, y, w, x)
            elif len(common) == 1:
                cc = common.pop()
                if w == cc and y == cc:
                    a, i, s = x, z, -1.0             # a_c^+ a_x^+ a_c a_z = -n_c a_x^+ a_z
                elif w == cc and z == cc:
                    a, i, s = x, y, 1.0              # a_c^+ a_x^+ a_y a_c = +n_c a_x^+ a_y
                elif x == cc and y == cc:
                    a, i, s = w, z, 1.0              # a_w^+ a_c^+ a_c a_z = +n_c a_w^+ a_z
                else:
                    a, i, s = w, y, -1.0             # a_w^+ a_c^+ a_y a_c = -n_c a_w^+ a_y
                if a > i:
                    add(3, (i, a, cc, -1), s * c)
            # two common indices: n n, Hermitian - it cancels in X - X^+
        else:
            raise ValueError("unexpected operator length")
    return {k: v for k, v in prim.items() if abs(v) > 1e-12}


def _operator(terms):
    """(kind int32[m], idx int32[m,4], coeff float64[m]) of the normalised pool operator, or None if it vanishes."""
    A = _anti_hermitian_normal_ordered(terms)
    if not A:
        return None
    norm = np.sqrt(sum(v * v for v in A.values()))
    prim = _primitives(A)
    if not prim:
        return None
    keys = sorted(prim)
    return (np.array([k for k, _ in keys], dtype=np.int32), np.array([ix for _, ix in keys], dtype=np.int32).reshape(-1, 4),
            np.array([prim[k] / norm for k in keys], dtype=np.float64))


def singlet_gsd_pool(n_orb: int):
    """Spin-adapted generalized singles and doubles; list of (kind, idx, coeff) operators (66 for n_orb = 4)."""
    pool = []
    for p in range(n_orb):
        pa, pb = 2 * p, 2 * p + 1
        for q in range(p, n_orb):
            qa, qb = 2 * q, 2 * q + 1
            op = _operator([(((pa, 1), (qa, 0)), 1.0), (((pb, 1), (qb, 0)), 1.0)])
            if op is not None:
                pool.append(op)
    pairs = [(p, q) for p in range(n_orb) for q in range(p, n_orb)]
    s12 = 1.0 / np.sqrt(12.0)
    for pq, (p, q) in enumerate(pairs):
        pa, pb, qa, qb = 2 * p, 2 * p + 1, 2 * q, 2 * q + 1
        for rs, (r, s) in enumerate(pairs):
            if pq > rs:
                continue
            ra, rb, sa, sb = 2 * r, 2 * r + 1, 2 * s, 2 * s + 1
            term_a = [(((ra, 1), (pa, 0), (sa, 1), (qa, 0)), 2 * s12), (((rb, 1), (pb, 0), (sb, 1), (qb, 0)), 2 * s12),
                      (((ra, 1), (pa, 0), (sb, 1), (qb, 0)), s12), (((rb, 1), (pb, 0), (sa, 1), (qa, 0)), s12),
                      (((ra, 1), (pb, 0), (sb, 1), (qa, 0)), s12), (((rb, 1), (pa, 0), (sa, 1), (qb, 0)), s12)]
            term_b = [(((ra, 1), (pa, 0), (sb, 1), (qb, 0)), 0.5), (((rb, 1), (pb, 0), (sa, 1), (qa, 0)), 0.5),
                      (((ra, 1), (pb, 0), (sb, 1), (qa, 0)), -0.5), (((rb, 1), (pa, 0), (sa, 1), (qb, 0)), -0.5)]
            for terms in (term_a, term_b):
                op = _operator(terms)
                if op is not None:
                    pool.append(op)
    return pool


def generalized_spin_orbital_pool(n_orb: int):
    """Every spin-conserving generalized single and double over spin orbitals, one primitive per operator."""
    nq = 2 * n_orb
    pool = []
    for p, q in combinations(range(nq), 2):
        if p % 2 == q % 2:
            pool.append((np.array([1], np.int32), np.array([[p, q, -1, -1]], np.int32), np.ones(1)))
    for quad in combinations(range(nq), 4):
        p, q, r, s = quad
        for (i, j), (a, b) in (((p, q), (r, s)), ((p, r), (q, s)), ((p, s), (q, r))):
            if sorted((i % 2, j % 2)) == sorted((a % 2, b % 2)):
                pool.append((np.array([2], np.int32), np.array([[i, j, a, b]], np.int32), np.ones(1)))
    return pool
