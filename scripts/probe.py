"""Scratch probe: time the hot path at a given size on the GPU (not a benchmark)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.jw import fragment_term_table
from multiscalequantumcomputing_b200.synth.hydrogen import random_fragment_integrals
from multiscalequantumcomputing_b200.uccsd import uccsd_factors

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
opts = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    opts[k] = int(v)
order = "lex" if opts.pop("lex", 0) else "blocked"
reps = opts.pop("reps", 3)
t = time.time()
h1, eri = random_fragment_integrals(n, 1234)
tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
F = uccsd_factors(n, n, order)
print("n", n, "terms", len(tab[0]), "factors", len(F), "host setup %.2fs" % (time.time() - t), flush=True)
h = _lib.Handle(2 * n)
for k, v in opts.items():
    h.set_option(k, v)
h.set_hamiltonian(*tab)
h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
for w in (0, 1):
    S = h.schedule(w)
    print("schedule", w, "tile_bits", S["tile_bits"], "passes", len(S["passes"]))
th = np.random.default_rng(4321).normal(0, 0.05, F.n_theta)
for r in range(reps):
    t = time.time(); e = h.energy(th); dt = time.time() - t
    print("energy      %.12f wall %.4fs" % (e, dt), h.last_timing(), flush=True)
for r in range(reps):
    t = time.time(); e, g = h.energy_grad(th); dt = time.time() - t
    print("energy_grad %.12f |g| %.6f wall %.4fs" % (e, np.linalg.norm(g), dt), h.last_timing(), flush=True)
t = time.time(); h.prepare(th); r1, r2 = h.rdm12(n, n // 2, n // 2); dt = time.time() - t
print("rdm12 wall %.4fs trace %.10f  sum2 %.8f" % (dt, np.trace(r1), np.einsum("iijj->", r2)))
