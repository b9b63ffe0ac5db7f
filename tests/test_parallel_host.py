"""Fragment-parallel dispatch on CPU: world_size-2 gloo run with the oracle behind the ansatz
surface; the assignment is fragment f -> worker f mod W and results equal the serial loop."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle_ansatz import OracleAnsatz
from multiscalequantumcomputing_b200.parallel import assign_round_robin, solve_fragments, solve_fragments_distributed
from multiscalequantumcomputing_b200.solver import solve
from multiscalequantumcomputing_b200.synth.hydrogen import hydrogen_chain, local_integrals
from multiscalequantumcomputing_b200.synth.dmet_loop import DMETRun, atom_clusters, fragment_problems


def _oracle_solver(*args, device=None, **kw):
    return solve(*args, warm_start=False, _ansatz_factory=lambda nq, tab, fac: OracleAnsatz(nq, tab, fac))


def test_round_robin():
    assert assign_round_robin(5, 2) == [[0, 2, 4], [1, 3]]
    assert assign_round_robin(2, 4) == [[0], [1], [], []]


def _problems():
    ints = local_integrals(hydrogen_chain(6, 0.9))
    return ints, atom_clusters(6, 2), fragment_problems(ints, atom_clusters(6, 2), 0.01, False)


def test_threaded_dispatch_matches_serial():
    ints, cl, probs = _problems()
    serial = [_oracle_solver(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, p.chempot) for p in probs]
    par = solve_fragments(probs, devices=[0, 1], solver=_oracle_solver)
    for (e1, r1), (e2, r2) in zip(serial, par):
        assert abs(e1 - e2) < 1e-12 and abs(r1 - r2).max() < 1e-12


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ints, cl, probs = _problems()
    res = solve_fragments_distributed(probs, solver=_oracle_solver)
    run = DMETRun(ints, cl, batch_solver=lambda ps: solve_fragments_distributed(ps, solver=_oracle_solver))
    nel = run.doexact(0.01)
    if rank == 0:
        q.put(([float(e) for e, _ in res], float(nel), float(run.energy)))
    dist.destroy_process_group()


def test_gloo_world2_matches_serial():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    energies, nel, etot = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ints, cl, probs = _problems()
    serial = [_oracle_solver(p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, p.chempot) for p in probs]
    assert np.allclose(energies, [e for e, _ in serial], atol=1e-10)
    run = DMETRun(ints, cl, solver=_oracle_solver)
    assert abs(run.doexact(0.01) - nel) < 1e-9 and abs(run.energy - etot) < 1e-9
