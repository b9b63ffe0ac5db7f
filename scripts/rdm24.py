"""Time the RDM kernels (K4) at BASELINE.json's 24-qubit size and check the trace identities."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from multiscalequantumcomputing_b200 import _lib
table, factors, theta, ng = bench.workload()
h = _lib.Handle(24)
h.set_hamiltonian(*table)
h.set_ansatz(factors.kind, factors.idx, factors.theta_index, factors.n_theta, factors.ref_index)
h.prepare(theta)
for rep in range(2):
    t = time.perf_counter(); r1, _ = h.rdm12(12, 6, 6, want_rdm2=False); t1 = time.perf_counter() - t
    t = time.perf_counter(); r1, r2 = h.rdm12(12, 6, 6); t2 = time.perf_counter() - t
    print("rdm1 %.4f s   rdm1+rdm2 %.4f s   tr dm1 = %.12f   sum dm2[iijj] = %.10f" % (t1, t2, np.trace(r1), np.einsum("iijj->", r2)))
# energy from the RDMs = <H> (bare Hamiltonian): E = sum h dm1 + 1/2 sum eri dm2
