"""ADAPT-VQE on top of the CUDA sweep.

The reference's default solver is not fixed UCCSD but ADAPT-VQE: `ADAPT(options['ansatz'], imp,
pool, ref)` with a generalized singles-and-doubles operator pool, grown `Nu` operators per macro
iteration by `ansatz.update(params, niter)` inside `imp_optimizer.run`
(`/root/reference/mqc/solver/dmet_solver_vqechem.py:35-41,54-65`,
`mqc/solver/dmet_interface.py:289-305`; `test/testlog_vqe:23-55` shows the macro iterations).
VQEChem is not in the tree, so this is the published ADAPT-VQE algorithm (Grimsley et al.,
Nat. Commun. 10, 3007) on the same surface:

* pool (`pools.py`): ``"sa"`` - the reference's spin-adapted fermionic pool (singlet generalized singles and
  doubles, 66 operators for four orbitals as in `test/testlog_vqe:23-24`; every operator is a linear combination
  of excitations that enters the ansatz in product form with one parameter, `mqc_set_ansatz_scaled`) - or
  ``"generalized"``: every spin-conserving generalized single and double over spin orbitals;
* pool gradient: dE/dtheta_k at theta_k = 0 of operator k appended AFTER the current ansatz.  The
  adjoint sweep yields all of them in ONE energy+gradient evaluation of the ansatz
  "chosen factors followed by the whole pool at zero angle" (zero-angle factors are identities, so
  their order does not matter and the state is unchanged);
* update: append the `nu` operators with the largest |gradient| (angle 0), re-optimise all
  angles; stop when the largest pool gradient falls below `eps`.

All state-vector arithmetic stays behind the C ABI (`mqc_set_ansatz` / `mqc_energy_grad`)."""
from __future__ import annotations

from itertools import combinations

import numpy as np

from .pools import generalized_spin_orbital_pool, singlet_gsd_pool
from .uccsd import FactorList, reference_index


def generalized_pool(n_orb: int):
    """(kind, idx4) arrays of all spin-conserving generalized singles and doubles."""
    nq = 2 * n_orb
    kind, idx = [], []
    for p, q in combinations(range(nq), 2):
        if p % 2 == q % 2:
            kind.append(1); idx.append((p, q, -1, -1))
    for quad in combinations(range(nq), 4):
        p, q, r, s = quad
        for (i, j), (a, b) in (((p, q), (r, s)), ((p, r), (q, s)), ((p, s), (q, r))):
            if sorted((i % 2, j % 2)) == sorted((a % 2, b % 2)):
                kind.append(2); idx.append((i, j, a, b))
    return np.array(kind, dtype=np.int32), np.array(idx, dtype=np.int32).reshape(-1, 4)


class ADAPTAnsatz:
    """Same surface as `ansatz.UCCSDAnsatz` (what `imp_optimizer.run` drives), with a growing
    factor list.  ``backend(factors)`` must return an object with ``energy_grad(theta)``; by
    default a `UCCSDAnsatz` on the CUDA library whose handle is re-programmed with
    `mqc_set_ansatz` whenever the ansatz grows."""

    def __init__(self, n_qubits, term_table, n_elec, device=0, options=None, nu=10, eps=1e-5,
                 max_ops=200, backend=None, pool="sa"):
        self.n_qubits, self.table, self.n_elec = n_qubits, term_table, n_elec
        self.n_orb = n_qubits // 2
        self.ref_index = reference_index(n_qubits, n_elec)
        if pool == "sa":
            self.pool_ops = singlet_gsd_pool(self.n_orb)
        elif pool == "generalized":
            self.pool_ops = generalized_spin_orbital_pool(self.n_orb)
        else:
            raise ValueError("unknown operator pool %r ('sa' or 'generalized')" % (pool,))
        self.pool_name = pool
        print_pool = (options or {}).get("verbose") if isinstance(options, dict) else False
        if print_pool:
            print(" Number of Operator: %6d" % len(self.pool_ops))
        self.nu, self.eps, self.max_ops = nu, eps, max_ops
        self.chosen = []
        self.converged = False
        self._params = np.zeros(0)
        self._niter = 0
        self._energy = None
        self._max_grad = None
        self.pool_grad_history = []
        self.n_evals = 0
        self._device, self._options = device, options
        self._backend_factory = backend
        self._ans = None
        self._program(self._factors(with_pool=False))

    # -- factor lists ---------------------------------------------------------------------
    def _factors(self, with_pool):
        """Chosen operators (one parameter each, possibly several factors) followed, for the gradient screen, by the
        whole pool at zero angle."""
        ids = list(self.chosen) + (list(range(len(self.pool_ops))) if with_pool else [])
        if not ids:
            return FactorList(self.n_qubits, np.zeros(0, np.int32), np.zeros((0, 4), np.int32), np.zeros(0, np.int32), 0,
                              self.ref_index)
        kind = np.concatenate([self.pool_ops[k][0] for k in ids])
        idx = np.concatenate([self.pool_ops[k][1] for k in ids])
        scale = np.concatenate([self.pool_ops[k][2] for k in ids])
        theta_index = np.concatenate([np.full(len(self.pool_ops[k][0]), t, dtype=np.int32) for t, k in enumerate(ids)])
        plain = bool(np.all(scale == 1.0))
        return FactorList(self.n_qubits, kind, idx, theta_index, len(ids), self.ref_index, None if plain else scale)

    @property
    def factors(self):
        return self._factors(with_pool=False)

    def _program(self, factors):
        if self._backend_factory is not None:
            self._ans = self._backend_factory(self.n_qubits, self.table, factors)
            return
        from .ansatz import UCCSDAnsatz
        if self._ans is None:
            self._ans = UCCSDAnsatz(self.n_qubits, self.table, factors, device=self._device, options=self._options)
        else:                                   # same handle, same Hamiltonian tables: only the ansatz changes
            self._ans.factors = factors
            self._ans.handle.set_ansatz(factors.kind, factors.idx, factors.theta_index, factors.n_theta, factors.ref_index,
                                        getattr(factors, "theta_scale", None))
            self._ans._cache_theta = None

    # -- reference surface ----------------------------------------------------------------
    def get_nparams_opt(self):
        return len(self.chosen)

    def get_number_of_hamiltonian_terms(self, params=None):
        return len(self.table[0])

    def pool_gradients(self, params):
        """dE/dtheta of every pool operator appended at zero angle after the current ansatz."""
        full = self._factors(with_pool=True)
        self._program(full)
        th = np.concatenate([np.asarray(params, dtype=np.float64), np.zeros(len(self.pool_ops))])
        e, g = self._ans._eval(th) if hasattr(self._ans, "_eval") else self._ans.energy_grad(th)
        self.n_evals += 1
        return e, np.array(g[len(self.chosen):])

    def update(self, params, niter):
        """`ansatz.update(params, niter)` of dmet_interface.py:290: grow by up to `nu` operators."""
        params = np.asarray(params, dtype=np.float64).flatten()
        e, pg = self.pool_gradients(params)
        self._energy = e
        self._max_grad = float(np.abs(pg).max())
        self.pool_grad_history.append(self._max_grad)
        if self._max_grad < self.eps or len(self.chosen) >= self.max_ops:
            self.converged = True
        else:
            order = np.argsort(-np.abs(pg))
            new = [int(k) for k in order[:self.nu] if abs(pg[k]) > 0.1 * self.eps]
            self.chosen += new
            params = np.concatenate([params, np.zeros(len(new))])
        self._program(self._factors(with_pool=False))
        self._params = params
        return params

    def energy(self, theta):
        e = self._ans.energy(theta)
        self._energy = e
        return e

    def gradient(self, theta):
        return self._ans.gradient(theta)

    def callback(self, theta):
        self._niter += 1
        self._params = np.array(theta, copy=True)

    def state(self, theta):
        return self._ans.state(theta)

    def prepare(self, theta):
        self._ans.prepare(theta)

    def rdm12(self, n_orb, n_alpha, n_beta, want_rdm2=True):
        return self._ans.rdm12(n_orb, n_alpha, n_beta, want_rdm2)

    def expectation(self, term_table):
        return self._ans.expectation(term_table)

    def close(self):
        if hasattr(self._ans, "close"):
            self._ans.close()
