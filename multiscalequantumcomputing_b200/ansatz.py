"""Inner boundary: the ansatz object SciPy drives (host side, Python like the reference).

Mirrors the surface `imp_optimizer.run` uses on VQEChem's ansatz
(`/root/reference/mqc/solver/dmet_interface.py:279-302`): ``energy(theta)``,
``gradient(theta)``, ``callback(theta)``, ``update(theta, niter)``, ``state(theta)``,
``_params / _niter / _energy / _max_grad``, ``get_nparams_opt()``,
``get_number_of_hamiltonian_terms()``.  All arithmetic happens in the CUDA library behind
the C ABI; there is no CPU implementation here."""
from __future__ import annotations

import numpy as np

from . import _lib
from .uccsd import FactorList


class UCCSDAnsatz:
    def __init__(self, n_qubits: int, term_table, factors: FactorList, device: int = 0, options=None, devices=None):
        """``devices``: list of 2, 4 or 8 GPU indices -> the state vector is sharded over them
        (`mqc_shard_group_create`; fragments beyond one GPU's memory).  A repeated index puts
        several shards on one GPU."""
        self.n_qubits = n_qubits
        self.factors = factors
        self.xmask, self.zmask, self.coeff = term_table
        self.handle = _lib.ShardGroup(n_qubits, list(devices)) if devices is not None and len(devices) > 1 \
            else _lib.Handle(n_qubits, device if devices is None else devices[0])
        for k, v in (options or {}).items():
            self.handle.set_option(k, v)
        self.handle.set_hamiltonian(self.xmask, self.zmask, self.coeff)
        self.handle.set_ansatz(factors.kind, factors.idx, factors.theta_index, factors.n_theta, factors.ref_index,
                               getattr(factors, "theta_scale", None))
        self._params = np.zeros(factors.n_theta)
        self._niter = 0
        self._energy = None
        self._max_grad = None
        self._cache_theta = None
        self._cache = None
        self.n_evals = 0

    def set_hamiltonian(self, term_table):
        """Next chemical-potential step / DMET iteration of the same fragment shape: keep the handle
        (ansatz, schedules, plans, buffers) and replace the Hamiltonian.  When the Pauli strings are
        unchanged only the coefficients travel (`mqc_update_hamiltonian_coeffs`)."""
        x, z, c = term_table
        same = len(x) == len(self.xmask) and np.array_equal(x, self.xmask) and np.array_equal(z, self.zmask)
        self.xmask, self.zmask, self.coeff = x, z, c
        if same:
            self.handle.update_hamiltonian_coeffs(c)
        else:
            self.handle.set_hamiltonian(x, z, c)
        self._cache_theta = None
        self._cache = None
        self._niter = 0
        self.n_evals = 0
        self.n_handle_reuses = getattr(self, "n_handle_reuses", 0) + 1

    # -- reference surface ------------------------------------------------------------
    def get_nparams_opt(self):
        return self.factors.n_theta

    def get_number_of_hamiltonian_terms(self, params=None):
        return len(self.xmask)

    def update(self, params, niter):
        """Fixed UCCSD: the ansatz does not grow (ADAPT's update appends operators)."""
        return params

    def _eval(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if self._cache_theta is None or not np.array_equal(theta, self._cache_theta):
            e, g = self.handle.energy_grad(theta)
            self._cache_theta = theta.copy()
            self._cache = (e, g.copy())
            self.n_evals += 1
        return self._cache

    def energy(self, theta):
        e, g = self._eval(theta)
        self._energy = e
        self._max_grad = float(np.abs(g).max()) if len(g) else 0.0
        return e

    def gradient(self, theta):
        return self._eval(theta)[1].copy()

    def energy_only(self, theta):
        return self.handle.energy(np.ascontiguousarray(theta, dtype=np.float64))

    def callback(self, theta):
        self._niter += 1
        self._params = np.array(theta, copy=True)

    def state(self, theta):
        """(2^N, 1) complex128 column in OpenFermion index order (`wfn[idx, 0]`,
        dmet_solver_vqechem.py:114)."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        if isinstance(self.handle, _lib.ShardGroup):
            return self.handle.gather_state(th).reshape(-1, 1)
        return self.handle.state(th).reshape(-1, 1)

    # -- extras -----------------------------------------------------------------------
    def prepare(self, theta):
        self.handle.prepare(np.ascontiguousarray(theta, dtype=np.float64))

    def rdm12(self, n_orb, n_alpha, n_beta, want_rdm2=True):
        return self.handle.rdm12(n_orb, n_alpha, n_beta, want_rdm2)

    def expectation(self, term_table):
        return self.handle.expectation(*term_table)

    def close(self):
        self.handle.close()
