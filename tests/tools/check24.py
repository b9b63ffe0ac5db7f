"""One-off check at BASELINE.json's full size: GPU vs the C oracle at 24 qubits (energy and
gradient).  Too slow for the test-suite (~2-4 min of CPU); result is recorded in profiles/."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from multiscalequantumcomputing_b200 import _lib
from oracle.cref import CRefUCCSD, num_threads
table, factors, theta, ng = bench.workload()
print("terms", len(table[0]), "groups", ng, "factors", len(factors), "cpu threads", num_threads(), flush=True)
h = _lib.Handle(24)
h.set_hamiltonian(*table)
h.set_ansatz(factors.kind, factors.idx, factors.theta_index, factors.n_theta, factors.ref_index)
e, g = h.energy_grad(theta)
print("gpu  E = %.14f" % e, h.last_timing(), flush=True)
hd = _lib.Handle(24); hd.set_option("sparse_tiles", 0)
hd.set_hamiltonian(*table)
hd.set_ansatz(factors.kind, factors.idx, factors.theta_index, factors.n_theta, factors.ref_index)
ed, gd = hd.energy_grad(theta)
print("gpu dense-tile E = %.14f  |dE| %.2e |dg| %.2e" % (ed, abs(ed - e), np.abs(gd - g).max()), hd.last_timing(), flush=True)
ref = CRefUCCSD(24, factors.kind, factors.idx, factors.theta_index, factors.n_theta, factors.ref_index, table)
t = time.time(); er, gr = ref.energy_grad(theta); dt = time.time() - t
print("cpu  E = %.14f  (%.1fs)" % (er, dt))
print("|dE| = %.3e   max|dg| = %.3e" % (abs(e - er), np.abs(g - gr).max()))
