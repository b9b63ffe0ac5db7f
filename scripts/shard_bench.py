#!/usr/bin/env python
"""Sharded UCCSD state-vector scaling point (BASELINE.json configs[4]).

    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/shard_bench.py \\
        --n-orb 17 --what energy [--dense] [--steps 2]

One process per GPU; see multiscalequantumcomputing_b200/shardbench.py (bench.py runs the same
points as its `sharded` block under torchrun).  Prints one JSON line on rank 0.

--dense      dense 2^K tiles (k_tile_sweep): streams every amplitude, the HBM-roofline kernel
(default)    compressed-sector tiles (k_tile_sp2): touches only the (N_alpha, N_beta) sector
--what       prepare (forward sweep only) | energy | grad (energy + adjoint gradient)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-orb", type=int, default=14)
    ap.add_argument("--what", default="energy", choices=["prepare", "energy", "grad"])
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--opt", action="append", default=[], help="library option key=value")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from multiscalequantumcomputing_b200.shardbench import sharded_point
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = sharded_point(args.n_orb, args.what, rank, world, local, steps=args.steps, warmup=args.warmup, dense=args.dense,
                        opts=[kv.split("=") for kv in args.opt])
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
