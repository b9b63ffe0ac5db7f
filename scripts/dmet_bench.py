"""Second BASELINE metric: DMET iteration wall time (configs[0]/[1]).
H10 chain (5 fragments x mu-steps, test/dmet_hchain.py shape) or H12 ring (6 fragments), 8 qubits
per fragment, fragments spread over the visible GPUs.  Reference: 745.97 s per iteration for the
H10 chain with its VQE solver (test/log_hchain_vqe:4039), 0.8-2.6 s with pyscf FCI."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.parallel import solve_fragments
from multiscalequantumcomputing_b200.solver import solve, clear_warm_start
from multiscalequantumcomputing_b200.synth.hydrogen import hydrogen_chain, hydrogen_ring, local_integrals
from multiscalequantumcomputing_b200.synth.dmet_loop import DMETRun, atom_clusters

which = sys.argv[1] if len(sys.argv) > 1 else "chain"
ngpu = int(sys.argv[2]) if len(sys.argv) > 2 else _lib.device_count()
ansatz = sys.argv[3] if len(sys.argv) > 3 else "uccsd"          # "adapt": the reference's default solver
geom = hydrogen_chain(10, 0.6) if which == "chain" else hydrogen_ring(12, 1.0)
n = len(geom)
ints = local_integrals(geom)
cl = atom_clusters(n, 2)
devs = list(range(ngpu))
for rep in range(3):
    clear_warm_start()
    run = DMETRun(ints, cl, batch_solver=lambda ps: solve_fragments(ps, devices=devs, warm_start=False,
                                                                     options={"ansatz": ansatz, "adapt": {"eps": 1e-6}}))
    t = time.perf_counter(); e = run.run(); dt = time.perf_counter() - t
    print(json.dumps({"metric": "dmet_iteration_wall_s", "system": which, "fragments": len(cl), "gpus": ngpu, "ansatz": ansatz,
                      "solves": run.n_solves, "mu_steps": len(run.history), "wall_s": dt, "energy": e, "mu": run.mu}), flush=True)
