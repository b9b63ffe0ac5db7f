"""Oracle-backed ansatz with the same surface as multiscalequantumcomputing_b200.ansatz.UCCSDAnsatz.
TEST INFRASTRUCTURE ONLY: lets the host logic of solve() run on a CPU-only box and gives the
CPU side of GPU-vs-oracle parity tests.  Never imported by the package."""
import numpy as np

from oracle.statevector import UCCSDOracle
from oracle import rdm as ordm
from oracle.hamiltonian import apply_terms


class OracleAnsatz:
    def __init__(self, n_qubits, term_table, factors):
        self.n_qubits = n_qubits
        self.factors = factors
        self.table = term_table
        self.orc = UCCSDOracle(n_qubits, factors.as_tuples(), factors.theta_index, factors.ref_index, term_table,
                               getattr(factors, "theta_scale", None))
        self._params = np.zeros(factors.n_theta)
        self._niter = 0
        self._energy = None
        self._max_grad = None
        self._ct = None
        self._c = None
        self.n_evals = 0
        self._psi = None

    def get_nparams_opt(self):
        return self.factors.n_theta

    def update(self, params, niter):
        return params

    def _eval(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        if self._ct is None or not np.array_equal(theta, self._ct):
            self._c = self.orc.energy_grad(theta)
            self._ct = theta.copy()
            self.n_evals += 1
        return self._c

    def energy(self, theta):
        e, g = self._eval(theta)
        self._energy = e
        self._max_grad = float(np.abs(g).max())
        return e

    def gradient(self, theta):
        return self._eval(theta)[1].copy()

    def callback(self, theta):
        self._niter += 1
        self._params = np.array(theta, copy=True)

    def state(self, theta):
        return self.orc.state(np.asarray(theta)).reshape(-1, 1)

    def prepare(self, theta):
        self._psi = self.orc.state(np.asarray(theta))

    def rdm12(self, n_orb, n_alpha, n_beta, want_rdm2=True):
        if n_alpha != n_beta:           # rdm.py's string tables are closed-shell (Sz = 0): literal ladder operators instead
            return ordm.rdm12_bruteforce(self._psi, n_orb)
        return ordm.get_rdm12(self._psi, n_alpha + n_beta, n_orb)

    def expectation(self, table):
        return complex(np.vdot(self._psi, apply_terms(*table, self._psi)))
