"""Second consumer of the kernels: the molecule-level VQE solver of the many-body
expansion and the measurement-style RDM path.

* `vqechem(...)` mirrors `/root/reference/mqc/solver/mbe_solver.py:179-217`: choose the active space the way the
  reference does (``ncas_occ = floor(ncas / 2)``, ``ncas_vir = ceil(ncas / 2)``, ``ncore = nocc - ncas_occ``,
  ``mo_list = range(ncore, ncore + ncas)``), run the VQE on it and return the total energy
  (``ansatz._energy`` there).  The reference builds the molecule and its integrals with pyscf (absent here), so
  this entry point starts from MO integrals: ``h_mo`` (n, n), ``eri_mo`` (n, n, n, n) in chemist order and the
  constant ``e_nuc``; the frozen core is folded into an effective one-body operator and a constant.
* `get_rdm12_openfermion(...)` mirrors `/root/reference/mqc/solver/dmet_solver_vqechem.py:119-179`: every RDM
  element is the expectation value of its Jordan-Wigner mapped operator ("for quantum circuit simulations this
  can be done through measurements", :134-135) - here one `mqc_expectation` (K2) per operator on the prepared
  state instead of a 2^N x 2^N sparse matrix per operator.  Same index convention as the reference:
  ``rdm1[p][q] = 2 <a^+_{p alpha} a_{q alpha}>``, ``rdm2[p][s][q][r] = sum_{sigma tau} <a^+_{p sigma} a^+_{q tau}
  a_{r tau} a_{s sigma}>``.

All state-vector arithmetic stays behind the C ABI."""
from __future__ import annotations

import numpy as np

from . import jw
from .ansatz import UCCSDAnsatz
from .solver import S2_SHIFT, imp_optimizer
from .uccsd import uccsd_factors


def active_space(n_orb: int, n_elec: int, ncas=None, ncore=None):
    """(ncore, mo_list) of `mbe_solver.py:199-210`."""
    nocc = n_elec // 2
    if ncas is None:
        ncas = n_orb
    ncas_occ = int(np.floor(ncas / 2))
    ncas_vir = int(np.ceil(ncas / 2))
    assert ncas_occ + ncas_vir == ncas
    nc = nocc - ncas_occ
    if nc < 0 or nc + ncas > n_orb:
        raise ValueError("active space of %d orbitals does not fit %d orbitals / %d electrons" % (ncas, n_orb, n_elec))
    return nc, list(range(nc, nc + ncas))


def frozen_core(h_mo, eri_mo, ncore: int, mo_list):
    """Effective active-space integrals: h_eff = h + sum_c [2 (pq|cc) - (pc|cq)], E_core = 2 sum_c h_cc +
    sum_cd [2 (cc|dd) - (cd|dc)]."""
    h = np.asarray(h_mo, dtype=np.float64)
    g = np.asarray(eri_mo, dtype=np.float64)
    core = list(range(ncore))
    act = list(mo_list)
    e_core = 2.0 * sum(h[c, c] for c in core)
    for c in core:
        for d in core:
            e_core += 2.0 * g[c, c, d, d] - g[c, d, d, c]
    h_eff = h[np.ix_(act, act)].copy()
    for c in core:
        h_eff += 2.0 * g[np.ix_(act, act, [c], [c])][:, :, 0, 0]          # Coulomb   2 (pq|cc)
        h_eff -= g[np.ix_(act, [c], [c], act)][:, 0, 0, :]                # exchange  (pc|cq)
    return h_eff, np.ascontiguousarray(g[np.ix_(act, act, act, act)]), float(e_core)


def vqechem(h_mo, eri_mo, n_elec: int, e_nuc: float = 0.0, ncas=None, ncore=None, *, device: int = 0, options=None,
            return_ansatz: bool = False, _ansatz_factory=None):
    """Molecule-level VQE (UCCSD, or ADAPT with ``options={"ansatz": "adapt"}``) on the active space; returns the
    total energy e_nuc + E_core + <H_active> at the optimum (the reference returns ``ansatz._energy``)."""
    n_orb = np.asarray(h_mo).shape[0]
    nc, mo_list = active_space(n_orb, n_elec, ncas, ncore)
    h_act, g_act, e_core = frozen_core(h_mo, eri_mo, nc, mo_list)
    n_act, ne_act = len(mo_list), n_elec - 2 * nc
    table = jw.fragment_term_table(h_act, g_act, 0, 0.0, S2_SHIFT / 2)
    bare = jw.fragment_term_table(h_act, g_act, 0, 0.0, 0.0)
    if (options or {}).get("ansatz", "uccsd") == "adapt":
        from .adapt import ADAPTAnsatz
        aopt = (options or {}).get("adapt", {})
        ansatz = ADAPTAnsatz(2 * n_act, table, ne_act, device=device, options=(options or {}).get("gpu"),
                             nu=aopt.get("nu", 10), eps=aopt.get("eps", 1e-5), max_ops=aopt.get("max_ops", 200),
                             backend=_ansatz_factory, pool=aopt.get("pool", "sa"))
    else:
        factors = uccsd_factors(n_act, ne_act, "blocked")
        make = _ansatz_factory or (lambda nq, tab, fac: UCCSDAnsatz(nq, tab, fac, device=device, options=(options or {}).get("gpu")))
        ansatz = make(2 * n_act, table, factors)
    ansatz, params = imp_optimizer(options).run(ansatz)
    ansatz.prepare(params)
    e = float(e_nuc) + e_core + ansatz.expectation(bare).real
    if return_ansatz:
        return e, ansatz, params
    if hasattr(ansatz, "close"):
        ansatz.close()
    return e


def _operator_table(n_qubits, ops):
    """JW Pauli table (index-bit masks, the C ABI's convention) of one ladder-operator product
    ``ops = ((qubit, action), ...)``, action 1 = creation (OpenFermion's FermionOperator term)."""
    idx = np.array([[q for q, _ in ops]], dtype=np.int64)
    act = np.array([[a for _, a in ops]], dtype=np.int64)
    x, z, c = jw._expand_products(idx, act, np.ones(1, dtype=np.complex128))
    return jw._combine([x], [z], [c], n_qubits, 0.0)


def get_rdm12_openfermion(ansatz, n_orb: int):
    """Measurement-style RDMs of the PREPARED state of ``ansatz`` (`dmet_solver_vqechem.py:119-179`)."""
    nq = 2 * n_orb
    ident = (np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.ones(1))
    S = ansatz.expectation(ident).real
    rdm1 = np.zeros([n_orb] * 2)
    rdm2 = np.zeros([n_orb] * 4)
    ev = lambda ops: ansatz.expectation(_operator_table(nq, ops)).real
    for p in range(0, nq, 2):
        for q in range(0, nq, 2):
            rdm1[p // 2][q // 2] = ev(((p, 1), (q, 0))) * 2 / S
            for r in range(0, nq, 2):
                for s in range(0, nq, 2):
                    v = 0.0
                    for (dp, dq, dr, ds) in ((0, 0, 0, 0), (1, 1, 1, 1), (1, 0, 0, 1), (0, 1, 1, 0)):
                        P, Q, R, T = p + dp, q + dq, r + dr, s + ds
                        if P == Q or R == T:
                            continue                     # a^+ a^+ or a a on the same spin orbital vanishes
                        v += ev(((P, 1), (Q, 1), (R, 0), (T, 0)))
                    rdm2[p // 2][s // 2][q // 2][r // 2] += v / S
    return rdm1, rdm2
