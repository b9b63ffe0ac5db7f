#!/usr/bin/env python
"""One process per GPU (torchrun): sharded UCCSD energy + gradient + RDMs through CUDA IPC peer
memory, checked against the CPU oracle on rank 0.  Used by tests/test_gpu_sharded.py and as a
stand-alone check:  torchrun --nproc-per-node 2 tests/tools/shard_ipc_check.py --n-orb 7"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-orb", type=int, default=7)
    ap.add_argument("--blocked", action="store_true", help="row-distributed sector matrix in H|psi>")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from multiscalequantumcomputing_b200 import _lib
    from multiscalequantumcomputing_b200.jw import fragment_term_table
    from multiscalequantumcomputing_b200.synth.hydrogen import random_fragment_integrals
    from multiscalequantumcomputing_b200.uccsd import uccsd_factors

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n_orb
    ne = n if n % 2 == 0 else n - 1
    h1, eri = random_fragment_integrals(n, 1234)
    tab = fragment_term_table(h1, eri, 0, 0.0, 0.25)
    F = uccsd_factors(n, ne, "blocked")
    th = np.random.default_rng(4321).normal(0, 0.05, F.n_theta)
    h = _lib.ShardRank(2 * n, local, rank, world)
    h.set_option("tile_bits", min(12, 2 * n - 4))
    h.set_option("grad_tile_bits", min(12, 2 * n - 4))
    if args.blocked:
        h.set_option("sigma_blocked", 1)
    h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    h.connect(want_lambda=True)
    e = h.energy(th)
    e2, g = h.energy_grad(th)
    h.prepare(th)
    r1, r2 = h.rdm12(n, ne // 2, ne // 2)
    ok = True
    if rank == 0:
        from oracle.statevector import UCCSDOracle
        from oracle import rdm as ordm
        orc = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, tab)
        e_ref, g_ref = orc.energy_grad(th)
        o1, o2 = ordm.get_rdm12(orc.state(th), ne, n)
        errs = (abs(e - e_ref), abs(e2 - e_ref), abs(g - g_ref).max(), abs(r1 - o1).max(), abs(r2 - o2).max())
        ok = errs[0] < 1e-10 and errs[1] < 1e-10 and errs[2] < 1e-9 and errs[3] < 1e-9 and errs[4] < 1e-9
        print("errors dE %.1e dE(grad call) %.1e dg %.1e drdm1 %.1e drdm2 %.1e" % errs)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    h.close()
    dist.destroy_process_group()
    if flag.item() != 1:
        sys.exit(1)
    if rank == 0:
        print("shard ipc ok: %d ranks, %d qubits" % (world, 2 * n))


if __name__ == "__main__":
    main()
