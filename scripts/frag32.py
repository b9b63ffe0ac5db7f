"""BASELINE.json configs[3] upper end: one 32-qubit fragment (16 orbitals, 16 electrons) on ONE B200:
UCCSD energy, energy + adjoint gradient and 1-/2-RDMs.  Synthetic seeded integrals (the POSCAR_acid
integrals need pyscf/openbabel)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.jw import fragment_term_table, number_of_x_groups
from multiscalequantumcomputing_b200.synth.hydrogen import random_fragment_integrals
from multiscalequantumcomputing_b200.uccsd import uccsd_factors
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
h1, eri = random_fragment_integrals(n, 1234)
tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
F = uccsd_factors(n, n, "blocked")
th = np.random.default_rng(4321).normal(0, 0.05, F.n_theta)
h = _lib.Handle(2 * n)
h.set_hamiltonian(*tab)
h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
out = {"n_qubits": 2 * n, "n_theta": F.n_theta, "n_terms": len(tab[0]), "n_x_groups": number_of_x_groups(tab[0])}
e = h.energy(th); t = h.last_timing()
e = h.energy(th); t = h.last_timing()
out["energy_ms"] = {k: t[k] for k in ("forward_ms", "hamiltonian_ms", "total_ms")}; out["energy"] = e
e2, g = h.energy_grad(th); t = h.last_timing()
out["energy_grad_ms"] = {k: t[k] for k in ("forward_ms", "hamiltonian_ms", "reverse_ms", "total_ms")}
out["passes"] = t["forward_passes"]; out["dE"] = abs(e2 - e); out["max_abs_grad"] = float(np.abs(g).max())
h.prepare(th)
t0 = time.perf_counter(); r1, _ = h.rdm12(n, n // 2, n // 2, want_rdm2=False); out["rdm1_s"] = time.perf_counter() - t0
out["tr_dm1"] = float(np.trace(r1))
if len(sys.argv) > 2:
    t0 = time.perf_counter(); r1, r2 = h.rdm12(n, n // 2, n // 2); out["rdm12_s"] = time.perf_counter() - t0
    out["sum_dm2_iijj"] = float(np.einsum("iijj->", r2))
print(json.dumps(out))
