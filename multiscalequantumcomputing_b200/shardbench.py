"""One measured point of the SHARDED state vector (BASELINE.json configs[4]): the
2^N complex128 amplitudes are split over the ranks of a torch.distributed job (one process per
GPU, top log2 G physical index bits = rank), peers are mapped with CUDA IPC (`mqc_shard_export /
connect`), tile passes whose active bits include a rank bit read and write the peers' HBM over
NVLink inside the compute kernel, ranks are ordered by the flag barrier `k_rank_barrier`.

Used by bench.py (so that the driver's N > 1 runs exercise CUDA IPC, the flag barrier and
rank-bit passes on N physical GPUs) and by scripts/shard_bench.py."""
from __future__ import annotations

import json
import os
import time
from math import comb

import numpy as np

from . import _lib
from .jw import fragment_term_table
from .synth.hydrogen import random_fragment_integrals
from .uccsd import uccsd_factors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def problem(n_orb: int, want_ham: bool = True):
    """Seeded generic fragment: integrals rng 1234, angles rng 4321."""
    ne = n_orb if n_orb % 2 == 0 else n_orb - 1
    F = uccsd_factors(n_orb, ne, "blocked")
    th = np.random.default_rng(4321).normal(0.0, 0.05, F.n_theta)
    tab = fragment_term_table(*random_fragment_integrals(n_orb, 1234), 0, 0.0, 0.25) if want_ham else None
    return ne, F, th, tab


def nvlink_counters(index: int):
    """(tx KiB, rx KiB) summed over the GPU's NVLinks, from NVML field values; None if unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        tx = pynvml.nvmlDeviceGetFieldValues(h, [(138, 0xFFFFFFFF)])[0]      # NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, all links
        rx = pynvml.nvmlDeviceGetFieldValues(h, [(139, 0xFFFFFFFF)])[0]      # NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX
        if tx.nvmlReturn != 0 or rx.nvmlReturn != 0:
            return None
        return int(tx.value.ullVal), int(rx.value.ullVal)
    except Exception:
        return None


def pass_stats(h, which: int, n_local: int):
    import ctypes as C
    npass, kb, ntf = C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(h.lib.mqc_schedule_info(h.h, which, C.byref(npass), C.byref(kb), C.byref(ntf)))
    by_ga = {}
    for p in range(npass.value):
        act = C.c_uint64()
        _lib.check(h.lib.mqc_schedule_pass(h.h, which, p, None, None, C.byref(act), None, None, None))
        ga = bin(act.value >> n_local).count("1")
        by_ga[ga] = by_ga.get(ga, 0) + 1
    return npass.value, kb.value, by_ga


def sharded_point(n_orb: int, what: str, rank: int, world: int, local: int, steps: int = 1, warmup: int = 1,
                  dense: bool = False, opts=(), keep_gradient: bool = False):
    """Collective over the default torch.distributed group.  ``what``: 'prepare' | 'energy' | 'grad'.
    Returns a dict on every rank (times are the max over ranks)."""
    import torch
    import torch.distributed as dist
    nq = 2 * n_orb
    g_bits = world.bit_length() - 1
    n_local = nq - g_bits
    t0 = time.perf_counter()
    ne, F, th, tab = problem(n_orb, what != "prepare")
    h = _lib.ShardRank(nq, local, rank, world) if world > 1 else _lib.Handle(nq, local)
    if dense:
        h.set_option("sparse_tiles", 0)
    if world == 1:
        h.set_option("compact", 0)            # the single-rank point of a scaling curve uses the same kernels
    for k, v in opts:
        h.set_option(k, int(v))
    if tab is not None:
        h.set_hamiltonian(*tab)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    if world > 1:
        h.connect(want_lambda=(what == "grad"))
    setup_s = time.perf_counter() - t0

    grad = None

    def step():
        nonlocal grad
        if what == "prepare":
            h.prepare(th)
            return None
        if what == "energy":
            return h.energy(th)
        e, grad = h.energy_grad(th)
        return e

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    val = None
    first_ms = None
    for i in range(warmup):
        t = time.perf_counter()
        val = step()
        if i == 0:
            first_ms = (time.perf_counter() - t) * 1e3
    barrier()
    nv0 = nvlink_counters(local)
    phases = np.zeros(4)
    t1 = time.perf_counter()
    for _ in range(steps):
        h.event_record(0)
        val = step()
        h.event_record(1)
        ms = h.event_elapsed_ms()
        tm = h.last_timing()
        phases += np.array([tm["forward_ms"], tm["hamiltonian_ms"], tm["reverse_ms"], ms])
    barrier()
    wall_ms = (time.perf_counter() - t1) * 1e3 / steps
    nv1 = nvlink_counters(local)
    phases /= steps
    if world > 1:
        t = torch.tensor(phases, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = t.cpu().numpy()
    norm = h.expectation(np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.ones(1)).real if what != "prepare" else None
    which = 1 if what == "grad" else 0
    npass, kb, by_ga = pass_stats(h, which, n_local)
    S_local = 16.0 * (1 << n_local)
    S_sec = 16.0 * comb(n_orb, (ne + 1) // 2) * comb(n_orb, ne // 2)
    hbm_fwd = npass * 2 * (S_local * world if dense else S_sec)
    nvl_fwd = sum(cnt * (1 - 0.5 ** ga) * 2 * (S_local if dense else S_sec / world) for ga, cnt in by_ga.items() if ga) * world
    peak = 6540.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    fwd_ms = phases[0] if what != "prepare" else phases[3]
    # algorithmic bytes of the whole evaluation on the sector: forward 2 S per pass, reverse 4 S per pass, H apply 2 S
    n_rev = npass if what == "grad" else 0
    alg_total = (2 * npass + 4 * n_rev + (2 if what != "prepare" else 0)) * (S_local * world if dense else S_sec)
    out = {"metric": "sharded_uccsd_%s" % what, "n_qubits": nq, "n_orb": n_orb, "n_elec": ne, "n_gpus": world,
           "transport": "one process per GPU (CUDA IPC + flag barrier)" if world > 1 else "single GPU, dense register",
           "tiles": "dense" if dense else "compressed-sector", "n_factors": len(F), "passes": npass, "tile_bits": kb,
           "passes_by_active_rank_bits": by_ga, "state_bytes_total": S_local * world, "sector_bytes_total": S_sec,
           "ms": {"forward": float(fwd_ms), "hamiltonian": float(phases[1]), "reverse": float(phases[2]), "total": float(phases[3])},
           "wall_ms_per_step": wall_ms, "first_call_ms": first_ms, "steps": steps, "warmup": warmup,
           "forward_algorithmic_hbm_bytes": hbm_fwd, "forward_nvlink_bytes_computed": nvl_fwd,
           "nvlink_KiB_this_gpu_measured": None if (nv0 is None or nv1 is None) else {"tx": (nv1[0] - nv0[0]) / steps, "rx": (nv1[1] - nv0[1]) / steps},
           "forward_aggregate_GBps": hbm_fwd / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None,
           "aggregate_hbm_peak_GBps": peak * world,
           "forward_frac_of_aggregate_hbm_peak": hbm_fwd / (fwd_ms * 1e-3) / 1e9 / (peak * world) if fwd_ms > 0 else None,
           "evaluation_frac_of_aggregate_hbm_peak": alg_total / (phases[3] * 1e-3) / 1e9 / (peak * world) if phases[3] > 0 else None,
           "value": val, "norm": norm, "setup_s": setup_s}
    if keep_gradient:
        out["_gradient"] = grad
    h.close()
    return out
