"""Time the 24-qubit bench workload under several library option sets (one handle each):
phase times of energy + adjoint gradient, CUDA events.  Usage:
    python scripts/sweep_opts.py "sp2_team=4,sp2_warps=4" "sp2_team=2,sp2_warps=8" ...
An empty string is the default configuration."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multiscalequantumcomputing_b200 import _lib  # noqa: E402


def main():
    table, factors, theta, _ = bench.workload()
    sets = sys.argv[1:] or [""]
    e0 = None
    for s in sets:
        h = _lib.Handle(factors.n_qubits, 0)
        try:
            for kv in filter(None, s.split(",")):
                k, v = kv.split("=")
                h.set_option(k, int(v))
            h.set_hamiltonian(*table)
            h.set_ansatz(factors.kind, factors.idx, factors.theta_index, factors.n_theta, factors.ref_index)
            for _ in range(3):
                e, g = h.energy_grad(theta)
            ph = np.zeros(4)
            n = 10
            for _ in range(n):
                e = h.energy_grad_resident()
                t = h.last_timing()
                ph += (t["forward_ms"], t["hamiltonian_ms"], t["reverse_ms"], t["total_ms"])
            ph /= n
            if e0 is None:
                e0 = e
            print(json.dumps({"opts": s, "fwd_ms": ph[0], "ham_ms": ph[1], "rev_ms": ph[2], "total_ms": ph[3],
                              "passes": t["forward_passes"], "launches": t["launches"], "dE": abs(e - e0)}), flush=True)
        except Exception as exc:
            print(json.dumps({"opts": s, "error": str(exc)}), flush=True)
        h.close()


if __name__ == "__main__":
    main()
