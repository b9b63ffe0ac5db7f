"""Outer boundary: drop-in for `mqc.solver.dmet_solver_vqechem.solve`.

    solve(oei_mo, one_body_mo, two_body_mo, n_imp, n_imp_orbs, chempot_imp=0.0, skip_vqe=True)
        -> (imp_energy: float, rdm1: ndarray[n, n])

Same positional signature, argument meaning and return contract as
`/root/reference/mqc/solver/dmet_solver_vqechem.py:17-88` (called positionally from
`mqc/third_party/qcdmet.py:108-110`): fragment one-/two-electron integrals in the embedding
basis, electron count (the reference's misleadingly named ``n_imp``), impurity size and
chemical potential go in; impurity energy and spin-summed 1-RDM come out.  The 2-RDM is
consumed internally (`:75-81`).  Differences, all deliberate:

* ansatz: fixed Jordan-Wigner UCCSD (BASELINE.json north star) instead of ADAPT-VQE; it runs
  in the RHF molecular-orbital basis of the embedding problem (as
  `mqc/solver/dmet_solver_fci.py:28-42` does for FCI) and the RDMs are rotated back to the
  embedding basis before the impurity-energy contraction.
* RDMs are float64 (the reference's float32 is a defect, `mqc/tools/rdm.py:137,171`).
* all state-vector arithmetic runs in the CUDA library (no CPU fallback).

`imp_optimizer` mirrors `mqc/solver/dmet_interface.py:273-313`.
"""
from __future__ import annotations

import threading
import time
import warnings
from dataclasses import dataclass, field

import numpy as np
from scipy.optimize import minimize

from . import jw
from .ansatz import UCCSDAnsatz
from .synth.hydrogen import rhf
from .uccsd import uccsd_factors

S2_SHIFT = 0.5          # options['scf']['shift'], dmet_solver_vqechem.py:37


class imp_optimizer:
    """VQE optimiser loop for DMET impurities (`dmet_interface.py:273-313`)."""

    def __init__(self, options=None):
        opt = (options or {}).get("opt", {})
        self._maxiter = opt.get("maxiter", 300)
        self._cgtol = opt.get("gtol", 1e-7)
        self._cgiter = opt.get("cgiter", 2000)
        self._method = opt.get("method", "BFGS")
        self._print_ham = 0
        self.verbose = opt.get("verbose", False)

    def check_converged(self, ansatz, result):
        # a fixed ansatz has nothing to grow: converged once the inner minimisation returns;
        # ADAPT: converged when the last update found no pool gradient above its threshold
        return getattr(ansatz, "converged", True)

    def print_info(self, ansatz, result):
        if self.verbose:
            print(" uccsd converged with %8d parameters" % ansatz.get_nparams_opt())
            print(" Converged Energy: %16.10f" % ansatz._energy)

    def run(self, ansatz, params=None):
        if params is None:
            params = [] if ansatz.get_nparams_opt() == 0 else ansatz._params
        params = np.asarray(params, dtype=np.float64).flatten()
        options = {"gtol": self._cgtol, "disp": False, "maxiter": self._cgiter}
        converged = False
        result = None
        for niter in range(self._maxiter):
            params = ansatz.update(params, niter)
            if len(params):
                result = minimize(ansatz.energy, params, method=self._method, jac=ansatz.gradient,
                                  options=options, callback=ansatz.callback)
                params = result["x"]
                if not result.get("success", True) and float(np.abs(result["jac"]).max()) > 100 * self._cgtol:
                    warnings.warn("VQE parameter optimisation stopped without converging: %s (max |g| = %.2e)"
                                  % (result.get("message", ""), float(np.abs(result["jac"]).max())))
            converged = self.check_converged(ansatz, result)
            if converged:
                self.print_info(ansatz, result)
                break
        self.result = result
        return ansatz, params


@dataclass
class SolveDetails:
    e_imp: float = 0.0
    rdm1: np.ndarray = None
    rdm2: np.ndarray = None
    e_uccsd: float = 0.0            # <H_bare> at the optimum (with the chemical potential)
    theta: np.ndarray = None
    n_evals: int = 0
    n_theta: int = 0
    n_terms: int = 0
    mo_coeff: np.ndarray = None
    t_setup: float = 0.0
    t_opt: float = 0.0
    t_rdm: float = 0.0
    timing: dict = field(default_factory=dict)


def embedding_mos(fock_mu, tei, n_elec):
    """RHF in the embedding space (pattern of dmet_solver_fci.py:28-42): gives the
    occupied/virtual split UCCSD needs."""
    e_el, mo_e, C, dm, ok = rhf(fock_mu, tei, n_elec, conv=1e-12)
    if not ok:
        warnings.warn("embedding RHF did not converge: the UCCSD reference determinant is built on unconverged orbitals")
    # fix the gauge so that CPU-oracle and GPU runs see identical orbitals
    for k in range(C.shape[1]):
        j = np.argmax(np.abs(C[:, k]))
        if C[j, k] < 0:
            C[:, k] = -C[:, k]
    return C, e_el


def prepare_problem(one_body_mo, two_body_mo, n_elec, n_imp_orbs, chempot_imp, s2_shift=S2_SHIFT / 2,
                    order="blocked"):
    """Everything the host computes before the GPU is involved: chemical potential, embedding
    RHF, MO transformation, JW term table (H + s2_shift S^2 and bare H), UCCSD factor list."""
    fock = np.array(one_body_mo, dtype=np.float64)
    tei = np.asarray(two_body_mo, dtype=np.float64)
    n_orb = fock.shape[0]
    fock_mu = fock.copy()
    if chempot_imp:
        d = np.arange(n_imp_orbs)
        fock_mu[d, d] -= chempot_imp
    C, e_rhf = embedding_mos(fock_mu, tei, n_elec)
    h_mo = C.T @ fock_mu @ C
    tei_mo = np.einsum("pqrs,pa,qb,rc,sd->abcd", tei, C, C, C, C, optimize=True)
    table = jw.fragment_term_table(h_mo, tei_mo, 0, 0.0, s2_shift)
    table_bare = jw.fragment_term_table(h_mo, tei_mo, 0, 0.0, 0.0) if s2_shift else table
    factors = uccsd_factors(n_orb, n_elec, order)
    return dict(n_orb=n_orb, C=C, e_rhf=e_rhf, table=table, table_bare=table_bare, factors=factors,
                h_mo=h_mo, tei_mo=tei_mo)


def finish_solve(rdm1_mo, rdm2_mo, C, oei_mo, one_body_mo, two_body_mo, n_imp_orbs):
    """MO -> embedding-basis RDMs and the impurity energy, dmet_solver_vqechem.py:75-81."""
    rdm1 = C @ rdm1_mo @ C.T
    rdm2 = np.einsum("abcd,pa,qb,rc,sd->pqrs", rdm2_mo, C, C, C, C, optimize=True)
    oei = np.asarray(oei_mo, dtype=np.float64)
    fock = np.asarray(one_body_mo, dtype=np.float64)
    tei = np.asarray(two_body_mo, dtype=np.float64)
    e = 0.5 * np.einsum("ij,ij->", rdm1[:n_imp_orbs, :], oei[:n_imp_orbs, :] + fock[:n_imp_orbs, :])
    e += 0.5 * np.einsum("ijkl,ijkl->", rdm2[:n_imp_orbs], tei[:n_imp_orbs])
    return float(e), rdm1, rdm2


# Per-fragment persistent state: QC-DMET calls the solver again for the same
# fragment at every chemical-potential step and DMET iteration (`mqc/third_party/qcdmet.py:244-260`)
# and the reference rebuilds everything each time (`mqc/solver/dmet_interface.py:45-128`).  A caller
# that passes ``fragment_id`` gets (a) its GPU handle back - ansatz, sweep schedules, plans, buffers
# stay, only the Hamiltonian coefficients travel (`mqc_update_hamiltonian_coeffs`) - and (b) its own
# warm-start parameters.  Handles are not thread safe: one fragment id must not be solved from two
# threads at the same time (`parallel.solve_fragments` keeps fragment i on worker i mod W).
_LOCK = threading.Lock()
_WARM = {}
_HANDLES = {}
STATS = {"handle_creates": 0, "handle_reuses": 0}


def _cache_key(fragment_id, n_orb, n_elec, device, options):
    gpu = tuple(sorted(((options or {}).get("gpu") or {}).items()))
    devs = tuple((options or {}).get("devices") or ())
    return (fragment_id, n_orb, n_elec, device, gpu, devs)


def solve(oei_mo, one_body_mo, two_body_mo, n_imp: int, n_imp_orbs: int, chempot_imp=0.0,
          skip_vqe: bool = True, *, device: int = 0, options=None, theta0=None, warm_start=None,
          fragment_id=None, return_details=False, verbose=False, _ansatz_factory=None):
    """``fragment_id`` (any hashable) switches on the per-fragment cache: the GPU handle is re-used
    across calls and, unless ``warm_start=False``, the optimisation starts from this fragment's last
    parameters.  Without it every call is independent of every other (the reference's behaviour);
    ``warm_start=True`` without a fragment id is refused because fragments of equal shape would seed
    each other."""
    t0 = time.perf_counter()
    n_elec = int(n_imp)                      # the reference's n_imp IS the electron count (:21)
    if chempot_imp is None:
        chempot_imp = 0.0
    if warm_start is None:
        warm_start = fragment_id is not None
    if warm_start and fragment_id is None:
        raise ValueError("warm_start=True needs fragment_id: the warm-start store is per fragment")
    prob = prepare_problem(one_body_mo, two_body_mo, n_elec, n_imp_orbs, chempot_imp)
    n_orb, factors = prob["n_orb"], prob["factors"]
    n_qubits = 2 * n_orb
    if verbose:
        print("running %d qubit uccsd calculation" % n_qubits)
    kind = (options or {}).get("ansatz", "uccsd")
    key = _cache_key(fragment_id, n_orb, n_elec, device, options)
    cached = False
    if kind == "adapt":
        # the reference's default: ADAPT-VQE, Nu operators per macro iteration (dmet_solver_vqechem.py:35-41)
        from .adapt import ADAPTAnsatz
        aopt = (options or {}).get("adapt", {})
        ansatz = ADAPTAnsatz(n_qubits, prob["table"], n_elec, device=device, options=(options or {}).get("gpu"),
                             nu=aopt.get("nu", 10), eps=aopt.get("eps", 1e-5), max_ops=aopt.get("max_ops", 200),
                             backend=_ansatz_factory,
                             pool=aopt.get("pool", "generalized" if (options or {}).get("devices") else "sa"))
        warm_start = False
    else:
        ansatz = None
        if fragment_id is not None and _ansatz_factory is None:
            with _LOCK:
                ansatz = _HANDLES.get(key)
        if ansatz is not None:
            ansatz.set_hamiltonian(prob["table"])
            ansatz._params = np.zeros(factors.n_theta)
            cached = True
            with _LOCK:
                STATS["handle_reuses"] += 1
        else:
            # options["devices"] = [0, 1, ...]: shard the fragment's state vector over these GPUs
            make = _ansatz_factory or (lambda nq, tab, fac: UCCSDAnsatz(nq, tab, fac, device=device, options=(options or {}).get("gpu"),
                                                                        devices=(options or {}).get("devices")))
            ansatz = make(n_qubits, prob["table"], factors)
            with _LOCK:
                STATS["handle_creates"] += 1
                if fragment_id is not None and _ansatz_factory is None:
                    _HANDLES[key] = ansatz
                    cached = True
    if theta0 is None and warm_start:
        with _LOCK:
            w = _WARM.get(key)
        if w is not None and len(w) == factors.n_theta:
            theta0 = w
    if theta0 is not None:
        ansatz._params = np.array(theta0, dtype=np.float64)
    t1 = time.perf_counter()
    opt = imp_optimizer(options)
    ansatz, params = opt.run(ansatz)
    t2 = time.perf_counter()
    if warm_start:
        with _LOCK:
            _WARM[key] = np.array(params)
    ansatz.prepare(params)
    n_alpha = n_elec // 2                      # noa / nob split of dmet_interface.py:84-85
    rdm1_mo, rdm2_mo = ansatz.rdm12(n_orb, n_alpha, n_elec - n_alpha)
    e_imp, rdm1, rdm2 = finish_solve(rdm1_mo, rdm2_mo, prob["C"], oei_mo, one_body_mo, two_body_mo, n_imp_orbs)
    t3 = time.perf_counter()
    if verbose:
        print("Impurity energy: %20.16f" % e_imp)
        print("time for uccsd calculation = %f" % (t2 - t0))
        print("time for constructing rdm1 and rdm2 = %f " % (t3 - t2))
    if return_details:
        e_bare = ansatz.expectation(prob["table_bare"]).real
        det = SolveDetails(e_imp=e_imp, rdm1=rdm1, rdm2=rdm2, e_uccsd=e_bare, theta=np.array(params),
                           n_evals=getattr(ansatz, "n_evals", 0), n_theta=ansatz.get_nparams_opt(),
                           n_terms=len(prob["table"][0]), mo_coeff=prob["C"],
                           t_setup=t1 - t0, t_opt=t2 - t1, t_rdm=t3 - t2)
        if hasattr(ansatz, "close") and not cached:
            ansatz.close()
        return e_imp, rdm1, det
    if hasattr(ansatz, "close") and not cached:
        ansatz.close()
    return e_imp, rdm1


def clear_warm_start():
    with _LOCK:
        _WARM.clear()


def clear_handle_cache():
    """Destroy the cached per-fragment GPU handles (end of a DMET run)."""
    with _LOCK:
        hs = list(_HANDLES.values())
        _HANDLES.clear()
    for a in hs:
        try:
            a.close()
        except Exception:
            pass
