#!/usr/bin/env python
"""Sharded UCCSD state-vector scaling point (BASELINE.json configs[4]).

    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/shard_bench.py \
        --n-orb 17 --what energy [--dense] [--steps 2]

One process per GPU; the state of 2^(2 n_orb) complex128 amplitudes is split over the G ranks
(top log2 G physical index bits = rank), peers are mapped with CUDA IPC, and tile passes whose
active bits include a rank bit stream the peers' shards over NVLink.  Prints one JSON line on
rank 0: per-phase device times (CUDA events, max over ranks), pass statistics, algorithmic
HBM / NVLink bytes and the achieved aggregate bandwidth against G x MEASURED_PEAKS hbm_gbs.

--dense      dense 2^K tiles (k_tile_sweep): streams every amplitude, the HBM-roofline kernel
(default)    compressed-sector tiles (k_tile_sparse): touches only the (N_alpha, N_beta) sector
--what       prepare (forward sweep only) | energy | grad (energy + adjoint gradient)
--group      single process driving all GPUs (mqc_shard_group_create) instead of torchrun ranks
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pass_stats(h, which, n_local):
    import ctypes as C
    from multiscalequantumcomputing_b200 import _lib
    npass, kb, ntf = C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(h.lib.mqc_schedule_info(h.h, which, C.byref(npass), C.byref(kb), C.byref(ntf)))
    by_ga = {}
    for p in range(npass.value):
        act = C.c_uint64()
        _lib.check(h.lib.mqc_schedule_pass(h.h, which, p, None, None, C.byref(act), None, None, None))
        ga = bin(act.value >> n_local).count("1")
        by_ga[ga] = by_ga.get(ga, 0) + 1
    return npass.value, kb.value, by_ga


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-orb", type=int, default=14)
    ap.add_argument("--n-elec", type=int, default=None)
    ap.add_argument("--what", default="energy", choices=["prepare", "energy", "grad"])
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--group", type=int, default=0, help="single process, this many GPUs")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--max-factors", type=int, default=0, help="truncate the ansatz (0 = full UCCSD)")
    ap.add_argument("--check-norm", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value")
    args = ap.parse_args()

    import torch
    from multiscalequantumcomputing_b200 import _lib
    from multiscalequantumcomputing_b200.jw import fragment_term_table, number_of_x_groups
    from multiscalequantumcomputing_b200.synth.hydrogen import random_fragment_integrals
    from multiscalequantumcomputing_b200.uccsd import uccsd_factors

    n = args.n_orb
    ne = args.n_elec or (n if n % 2 == 0 else n - 1)
    nq = 2 * n
    dist = None
    if args.group:
        rank, world, local = 0, args.group, 0
    else:
        import torch.distributed as dist
        rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g_bits = world.bit_length() - 1
    n_local = nq - g_bits

    t0 = time.time()
    F = uccsd_factors(n, ne, "blocked")
    if args.max_factors:
        import dataclasses
        k = args.max_factors
        F = dataclasses.replace(F, kind=F.kind[:k], idx=F.idx[:k], theta_index=F.theta_index[:k], n_theta=k)
    th = np.random.default_rng(4321).normal(0, 0.05, F.n_theta)
    tab = None
    if args.what != "prepare":
        h1, eri = random_fragment_integrals(n, 1234)
        tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    if args.group:
        h = _lib.ShardGroup(nq, list(range(world))) if world > 1 else _lib.Handle(nq, 0)
    else:
        h = _lib.ShardRank(nq, local, rank, world) if world > 1 else _lib.Handle(nq, local)
    if args.dense:
        h.set_option("sparse_tiles", 0)
    for kv in args.opt:
        k, v = kv.split("=")
        h.set_option(k, int(v))
    if tab is not None:
        h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    if not args.group and world > 1:
        h.connect(want_lambda=(args.what == "grad"))
    setup_s = time.time() - t0

    def step():
        if args.what == "prepare":
            h.prepare(th)
            return None
        if args.what == "energy":
            return h.energy(th)
        return h.energy_grad(th)[0]

    def barrier():
        if dist is not None and world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    val = None
    for _ in range(args.warmup):
        val = step()
    barrier()
    t1 = time.perf_counter()
    phases = np.zeros(4)
    for _ in range(args.steps):
        h.event_record(0)
        val = step()
        h.event_record(1)
        ms = h.event_elapsed_ms()
        tm = h.last_timing()
        phases += np.array([tm["forward_ms"], tm["hamiltonian_ms"], tm["reverse_ms"], ms])
    barrier()
    wall = (time.perf_counter() - t1) / args.steps
    phases /= args.steps
    if dist is not None and world > 1:
        t = torch.tensor(phases, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = t.cpu().numpy()
    norm = None
    if args.check_norm:
        xz = (np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.ones(1))
        if args.what == "prepare":
            h.prepare(th)
        norm = h.expectation(*xz).real

    which = 1 if args.what == "grad" else 0
    npass, kb, by_ga = pass_stats(h, which, n_local)
    S_local = 16.0 * (1 << n_local)
    na, nb = (ne + 1) // 2, ne // 2
    from math import comb
    S_sec = 16.0 * comb(n, na) * comb(n, nb)
    if args.dense:
        hbm_fwd = npass * 2 * S_local * world              # every pass reads + writes every shard
    else:
        hbm_fwd = npass * 2 * S_sec
    nvl_fwd = sum(cnt * (1 - 0.5 ** ga) * 2 * (S_local if args.dense else S_sec / world) for ga, cnt in by_ga.items() if ga) * world
    peak = 6540.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    fwd_ms = phases[0] if args.what != "prepare" else phases[3]
    out = {
        "metric": "sharded_uccsd_%s" % args.what, "n_qubits": nq, "n_orb": n, "n_elec": ne, "n_gpus": world,
        "transport": "same-process group (events)" if args.group else "one process per GPU (CUDA IPC + flag barrier)",
        "tiles": "dense" if args.dense else "compressed-sector",
        "n_factors": len(F), "passes": npass, "tile_bits": kb, "passes_by_active_rank_bits": by_ga,
        "state_bytes_total": S_local * world, "sector_bytes_total": S_sec,
        "ms": {"forward": float(fwd_ms), "hamiltonian": float(phases[1]), "reverse": float(phases[2]), "total": float(phases[3])},
        "wall_ms_per_step": wall * 1e3, "steps": args.steps, "warmup": args.warmup,
        "forward_algorithmic_hbm_bytes": hbm_fwd, "forward_nvlink_bytes": nvl_fwd,
        "forward_aggregate_GBps": hbm_fwd / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None,
        "aggregate_hbm_peak_GBps": peak * world,
        "forward_frac_of_aggregate_hbm_peak": hbm_fwd / (fwd_ms * 1e-3) / 1e9 / (peak * world) if fwd_ms > 0 else None,
        "value": val, "norm": norm, "setup_s": setup_s, "options": args.opt,
    }
    if rank == 0:
        print(json.dumps(out))
    h.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
