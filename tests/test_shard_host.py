"""Host side of the sharded state with one process per rank, on CPU (gloo, world size 2):
every rank builds the same schedule, the ranks' tile maps partition every pass, the CUDA IPC
blobs travel in rank order through `ShardRank.connect`, and RDM partial sums are all-reduced.
(The device side of the same protocol is covered by tests/test_gpu_sharded.py.)"""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.uccsd import uccsd_factors


def _deposit(t, pos):
    out = 0
    for j, p in enumerate(pos):
        out |= ((t >> j) & 1) << p
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 6
    F = uccsd_factors(n, 6, "blocked")
    h = _lib.ShardRank(2 * n, 0, rank, world)
    h.set_option("tile_bits", 8)
    h.set_option("low_bits", 2)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    S = h.schedule(0)
    n_local = h.shard_info()["n_local_bits"]
    # tiles this rank sweeps in every pass (its own view of the tile map)
    mine = []
    for P in S["passes"]:
        act = int(P["act_mask"])
        inactive_local = [b for b in range(n_local) if not (act >> b) & 1]
        tau_hi, fixed, ntiles = P["tile_map"][rank]
        mine.append(sorted(_deposit(t | tau_hi, inactive_local) | fixed for t in range(ntiles)))
    layouts = [None] * world
    dist.all_gather_object(layouts, (list(map(int, S["final_layout"])), [len(p["factors"]) for p in S["passes"]]))
    tiles = [None] * world
    dist.all_gather_object(tiles, mine)
    # connect(): blobs in rank order, no GPU needed when export / connect_blobs are stubbed
    seen = {}
    h.export = lambda want_lambda=True: bytes([rank]) * _lib.BLOB_BYTES
    h.connect_blobs = lambda blobs: seen.setdefault("blobs", blobs)
    h.connect(want_lambda=True)
    # rdm12(): partial sums of the ranks are added up
    _lib.Handle.rdm12 = lambda self, n_orb, na, nb, want_rdm2=True: (np.full((n_orb, n_orb), rank + 1.0), np.full((n_orb,) * 4, 10.0 * (rank + 1)))
    r1, r2 = h.rdm12(n, 3, 3)
    if rank == 0:
        q.put(dict(layouts=layouts, tiles=tiles, blobs=[b[0] for b in seen["blobs"]], r1=float(r1[0, 0]), r2=float(r2[0, 0, 0, 0]),
                   acts=[int(P["act_mask"]) for P in S["passes"]], n_local=n_local, k=S["tile_bits"]))
    dist.destroy_process_group()


def test_gloo_world2_sharded_host_protocol():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out["layouts"][0] == out["layouts"][1]                      # identical schedules on both ranks
    assert out["blobs"] == [0, 1]
    assert out["r1"] == 3.0 and out["r2"] == 30.0
    n_bits, K = 12, out["k"]
    n_glob = 0
    for p, act in enumerate(out["acts"]):
        t0, t1 = out["tiles"][0][p], out["tiles"][1][p]
        bases = t0 + t1
        assert len(set(bases)) == len(bases) == 1 << (n_bits - K)      # disjoint and complete
        assert all((b & act) == 0 for b in bases)
        if act >> out["n_local"]:
            n_glob += 1
            assert all((b >> out["n_local"]) == 0 for b in bases)      # the rank bit is inside the tile
        else:
            assert all((b >> out["n_local"]) == 0 for b in t0) and all((b >> out["n_local"]) == 1 for b in t1)
    assert n_glob > 0
