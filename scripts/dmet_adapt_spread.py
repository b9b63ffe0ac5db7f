"""Run-to-run spread of the DMET total energy (H10 ring, ADAPT on the GPU) against the reference's FCI-solver log
(test/testlog_fci:438) for the two settings tests/test_gpu_parity.py asserts on.  python scripts/dmet_adapt_spread.py [runs]"""
import json
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiscalequantumcomputing_b200.solver import solve  # noqa: E402
from multiscalequantumcomputing_b200.synth.dmet_loop import DMETRun, atom_clusters  # noqa: E402
from multiscalequantumcomputing_b200.synth.hydrogen import hydrogen_ring, local_integrals  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E_REF = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_logs.json")))["ring_fci"]["dmet_energy"]
ints = local_integrals(hydrogen_ring(10, 1.0))
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 10
warnings.simplefilter("ignore")
for pool, eps, gtol in (("generalized", 1e-8, 1e-10), ("sa", 1e-7, 1e-9)):
    opts = {"ansatz": "adapt", "adapt": {"eps": eps, "nu": 10, "pool": pool}, "opt": {"gtol": gtol}}
    d = []
    for _ in range(runs):
        run = DMETRun(ints, atom_clusters(10, 2), trans_inv=True, solver=lambda *a: solve(*a, options=opts))
        d.append(abs(run.run() - E_REF))
    print(json.dumps({"pool": pool, "adapt_eps": eps, "gtol": gtol, "runs": runs, "abs_diff": d, "max": max(d), "min": min(d)}), flush=True)
