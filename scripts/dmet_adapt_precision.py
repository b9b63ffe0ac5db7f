"""DMET total energy of the H10 ring (2-atom fragments, translation invariant) with the ADAPT solver on the GPU
against the reference's FCI-solver log (test/testlog_fci:438), as a function of the ADAPT / optimiser thresholds."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiscalequantumcomputing_b200.solver import solve  # noqa: E402
from multiscalequantumcomputing_b200.synth.dmet_loop import DMETRun, atom_clusters  # noqa: E402
from multiscalequantumcomputing_b200.synth.hydrogen import hydrogen_ring, local_integrals  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_logs.json")))
E_REF = GOLD["ring_fci"]["dmet_energy"]
ints = local_integrals(hydrogen_ring(10, 1.0))
import itertools
for pool, (eps, gtol) in itertools.product(("generalized", "sa"), ((1e-5, 1e-7), (1e-6, 1e-7), (1e-7, 1e-9), (1e-8, 1e-10))):
    opts = {"ansatz": "adapt", "adapt": {"eps": eps, "nu": 10, "pool": pool}, "opt": {"gtol": gtol}}
    run = DMETRun(ints, atom_clusters(10, 2), trans_inv=True, solver=lambda *a: solve(*a, options=opts))
    t = time.perf_counter()
    e = run.run()
    print(json.dumps({"pool": pool, "adapt_eps": eps, "gtol": gtol, "E_dmet": e, "E_reference_fci_log": E_REF, "abs_diff": abs(e - E_REF),
                      "mu_steps": len(run.history), "wall_s": time.perf_counter() - t}), flush=True)
