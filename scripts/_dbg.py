import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.uccsd import uccsd_factors
from oracle.statevector import UCCSDOracle
EMPTY = (np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.complex128))
for n, ne, kb, opts in ((7, 6, 10, {}), (8, 8, 12, {}), (8, 8, 12, {"sp2_warps": 1}), (8, 8, 11, {}), (8, 8, 10, {}), (8, 6, 12, {})):
    F = uccsd_factors(n, ne, "blocked")
    h = _lib.Handle(2 * n)
    h.set_option("tile_bits", kb)
    for k, v in opts.items(): h.set_option(k, v)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    th = np.random.default_rng(5).normal(0, 0.2, F.n_theta)
    ref = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, EMPTY).state(th)
    psi = h.state(th)
    S = h.schedule(0)
    print(n, ne, kb, opts, "err", abs(psi - ref).max(), "passes", len(S["passes"]), "nf", [len(p["factors"]) for p in S["passes"]][:8])
