"""Host-side DMET caller used to drive the fragment solver in tests and benchmarks.

The reference's DMET loop is `MyDMET.doexact` / `doselfconsistent`
(`/root/reference/mqc/third_party/qcdmet.py:28-235, 237-306`) on top of the external
QC-DMET package (`localintegrals`, `qcdmet_helper`), neither of which is importable
here.  The loop is OUT OF SCOPE for acceleration; this file is
the minimum restatement needed to *call* the hot path the way the reference does:

* mean-field 1-RDM in the localized basis  (QC-DMET ``construct1RDM_loc``, umat = 0)
* Schmidt bath from the environment block  (QC-DMET ``constructbath``;
  core-occupation snapping `qcdmet.py:51-59`)
* embedded integrals ``dmet_oei / dmet_fock / dmet_tei``   (`qcdmet.py:69-71`)
* per-fragment dispatch ``solve(dmetOEI, dmetFOCK, dmetTEI, Nelec_in_imp,
  numImpOrbs, chempot_imp)``                              (`qcdmet.py:108-110`)
* energy / electron-number bookkeeping                     (`qcdmet.py:139-156, 224-232`)
* secant search on the chemical potential                  (`qcdmet.py:257`,
  ``scipy.optimize.newton(..., tol=1e-8, maxiter=50)``), one pass (SCmethod 'NONE').

The fragment loop (`qcdmet.py:42`) is restructured as "collect inputs -> batch solve ->
scatter results" so a batch solver can place one fragment per GPU;
the DMET mathematics is unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Sequence

import numpy as np
from scipy import optimize

from .hydrogen import LocalIntegrals


def construct_1rdm_loc(ints: LocalIntegrals, umat=None):
    """QC-DMET helper.construct1RDM_loc with doSCF=False: diagonalise the localized
    Fock (+u) and doubly occupy the lowest N/2 orbitals."""
    F = ints.fock_loc if umat is None else ints.fock_loc + umat
    e, c = np.linalg.eigh(F)
    nocc = ints.nelec // 2
    return 2.0 * c[:, :nocc] @ c[:, :nocc].T


def construct_bath(one_rdm, impurity_orbs, n_bath, threshold=1e-13):
    """QC-DMET helper.constructbath: eigen-decompose the environment block of the 1-RDM,
    keep the ``n_bath`` orbitals whose occupations are farthest from 0/2, order the
    rest by decreasing occupation.  Returns (n_bath, loc2dmet, core_occupations)."""
    impurity_orbs = np.asarray(impurity_orbs).astype(bool)
    env = ~impurity_orbs
    n_imp = int(impurity_orbs.sum())
    n_tot = len(impurity_orbs)
    block = one_rdm[np.ix_(env, env)]
    lam, vec = np.linalg.eigh(block)
    key = np.maximum(-lam, lam - 2.0)
    idx = key.argsort()
    tokeep = int(np.sum(-key[idx] > threshold))
    n_bath = min(n_bath, tokeep)
    lam = lam[idx]
    vec = vec[:, idx]
    rest = (-lam[n_bath:]).argsort()
    vec[:, n_bath:] = vec[:, n_bath:][:, rest]
    core = np.hstack((np.zeros(n_imp + n_bath), lam[n_bath:][rest]))
    loc2dmet = np.zeros((n_tot, n_tot))
    imp_idx = np.where(impurity_orbs)[0]
    env_idx = np.where(env)[0]
    loc2dmet[imp_idx, np.arange(n_imp)] = 1.0
    loc2dmet[np.ix_(env_idx, n_imp + np.arange(len(env_idx)))] = vec
    return n_bath, loc2dmet, core


def _jk_loc(eri, dm):
    J = np.einsum("pqrs,rs->pq", eri, dm)
    K = np.einsum("prqs,rs->pq", eri, dm)
    return J - 0.5 * K


@dataclass
class FragmentProblem:
    """Inputs of one `solve()` call, exactly the positional arguments of qcdmet.py:108-110."""
    oei: np.ndarray
    fock: np.ndarray
    tei: np.ndarray
    nelec: int
    n_imp_orbs: int
    chempot: float


def fragment_problems(ints: LocalIntegrals, clusters: Sequence[np.ndarray], mu: float,
                      trans_inv: bool = False, umat=None) -> List[FragmentProblem]:
    """qcdmet.py:29-71 for every fragment (one if translation invariant)."""
    one_rdm = construct_1rdm_loc(ints, umat)
    out = []
    nfrag = 1 if trans_inv else len(clusters)
    for f in range(nfrag):
        imp = np.abs(np.asarray(clusters[f]))
        n_imp = int(imp.sum())
        n_bath, loc2dmet, core = construct_bath(one_rdm, imp, n_imp)
        core = core.copy()
        for k in range(len(core)):                     # qcdmet.py:51-59
            if core[k] < 0.01:
                core[k] = 0.0
            elif core[k] > 1.99:
                core[k] = 2.0
            else:
                raise AssertionError("bad DMET bath orbital selection: occupation %r" % core[k])
        n_act = n_imp + n_bath
        nelec = int(round(ints.nelec - core.sum()))
        core_loc = loc2dmet @ np.diag(core) @ loc2dmet.T
        A = loc2dmet[:, :n_act]
        oei = A.T @ ints.h_loc @ A
        fock = A.T @ (ints.h_loc + _jk_loc(ints.eri_loc, core_loc)) @ A
        tei = np.einsum("pqrs,pa,qb,rc,sd->abcd", ints.eri_loc, A, A, A, A, optimize=True)
        out.append(FragmentProblem(np.ascontiguousarray(oei), np.ascontiguousarray(fock),
                                   np.ascontiguousarray(tei), nelec, n_imp, mu))
    return out


Solver = Callable[..., tuple]


@dataclass
class DMETRun:
    """One-shot DMET (SCmethod 'NONE', as `mqc/dcalgo/dmet.py:7` defaults)."""
    ints: LocalIntegrals
    clusters: Sequence[np.ndarray]
    solver: Solver = None                  # solve(oei, fock, tei, nelec, n_imp_orbs, mu) -> (E_imp, rdm1)
    batch_solver: Callable = None          # list[FragmentProblem] -> list[(E_imp, rdm1)]
    trans_inv: bool = False
    energy: float = 0.0
    mu: float = 0.0
    history: list = field(default_factory=list)
    n_solves: int = 0

    def doexact(self, mu: float = 0.0) -> float:
        probs = fragment_problems(self.ints, self.clusters, mu, self.trans_inv)
        if self.batch_solver is not None:
            results = self.batch_solver(probs)
        else:
            results = [self.solver(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, p.chempot) for p in probs]
        self.n_solves += len(probs)
        energy = 0.0
        nelec = 0.0
        for p, (e_imp, rdm1) in zip(probs, results):
            energy += e_imp
            nelec += np.trace(np.asarray(rdm1, dtype=np.float64)[:p.n_imp_orbs, :p.n_imp_orbs])
        if self.trans_inv:
            energy *= len(self.clusters)
            nelec *= len(self.clusters)
        self.energy = energy + self.ints.e_nuc
        self.history.append((mu, nelec, self.energy))
        return nelec

    def run(self, mu0: float = 0.0) -> float:
        cost = lambda mu: self.doexact(float(mu)) - self.ints.nelec
        self.mu = float(optimize.newton(cost, mu0, tol=1e-8, maxiter=50, disp=False))
        return self.energy


def atom_clusters(norb: int, natom_per_fragment: int):
    """`SimpleFragment(natom_per_fragment)` + `get_impurity_clusters`
    (`mqc/dcalgo/dmet.py:49-56`) for one 1s orbital per H atom."""
    out = []
    for start in range(0, norb, natom_per_fragment):
        v = np.zeros(norb, dtype=int)
        v[start:start + natom_per_fragment] = 1
        out.append(v)
    return out
