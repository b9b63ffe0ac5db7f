"""One large fragment on ONE GPU: energy + adjoint gradient, compact sector storage vs the dense
register (option compact=0).  python scripts/big_single.py N_ORB [key=value ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiscalequantumcomputing_b200 import _lib  # noqa: E402
from multiscalequantumcomputing_b200.jw import fragment_term_table  # noqa: E402
from multiscalequantumcomputing_b200.synth.hydrogen import random_fragment_integrals  # noqa: E402
from multiscalequantumcomputing_b200.uccsd import uccsd_factors  # noqa: E402


def main():
    n = int(sys.argv[1])
    energy_only = "energy" in sys.argv[2:]
    opts = [kv.split("=") for kv in sys.argv[2:] if "=" in kv]
    reps_opt = [int(v) for k, v in opts if k == "reps"]
    opts = [(k, v) for k, v in opts if k != "reps"]
    ne = n if n % 2 == 0 else n - 1
    t0 = time.perf_counter()
    tab = fragment_term_table(*random_fragment_integrals(n, 1234), 0, 0.0, 0.25)
    F = uccsd_factors(n, ne, "blocked")
    th = np.random.default_rng(4321).normal(0.0, 0.05, F.n_theta)
    t1 = time.perf_counter()
    h = _lib.Handle(2 * n, 0)
    for k, v in opts:
        h.set_option(k, int(v))
    h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    t2 = time.perf_counter()
    if energy_only:
        e, g = h.energy(th), np.zeros(1)
    else:
        e, g = h.energy_grad(th)
    t3 = time.perf_counter()
    reps = reps_opt[0] if reps_opt else (1 if energy_only else 2)
    ph = np.zeros(4)
    for _ in range(reps):
        e2 = h.energy(th) if energy_only else h.energy_grad_resident()
        t = h.last_timing()
        ph += (t["forward_ms"], t["hamiltonian_ms"], t["reverse_ms"], t["total_ms"])
    ph /= reps
    from math import comb
    ndet = comb(n, ne // 2) * comb(n, ne - ne // 2)
    info = h.cplan_info(0 if energy_only else 1)
    ab = h.amplitude_bytes() or 16
    norm = h.expectation(np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.ones(1)).real
    print(json.dumps({"n_qubits": 2 * n, "opts": dict(opts), "n_theta": F.n_theta, "n_terms": len(tab[0]), "n_det": ndet,
                      "amplitude_bytes": ab, "sector_MB": ab * 1e-6 * ndet, "compact_passes": info["n_passes"], "passes": t["forward_passes"],
                      "host_tables_s": t1 - t0, "setup_s": t2 - t1, "first_eval_s": t3 - t2,
                      "fwd_ms": ph[0], "ham_ms": ph[1], "rev_ms": ph[2], "total_ms": ph[3], "energy": e, "dE_repeat": abs(e - e2),
                      "fwd_GBps_per_pass": 2 * ab * 1e-9 * ndet / (ph[0] * 1e-3 / max(1, t["forward_passes"])),
                      "rev_GBps_per_pass": 4 * ab * 1e-9 * ndet / (ph[2] * 1e-3 / max(1, t["reverse_passes"])) if ph[2] > 0 else None,
                      "gnorm": float(np.linalg.norm(g)), "norm": norm, "what": "energy" if energy_only else "energy+gradient"}), flush=True)
    h.close()


if __name__ == "__main__":
    main()
