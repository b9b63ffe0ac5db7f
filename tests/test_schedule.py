"""Host-side sweep scheduler (C++ in the CUDA library, reached through the C ABI without a
GPU): executing its passes with the numpy interpreter reproduces the oracle state."""
import numpy as np
import pytest

import schedule_interp as SI
from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.uccsd import uccsd_factors
from oracle.statevector import UCCSDOracle

EMPTY = (np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.complex128))


@pytest.mark.parametrize("n,ne,kb,lb,order", [
    (4, 4, 12, 3, "blocked"),     # whole register in one tile
    (4, 4, 6, 2, "lex"),
    (5, 4, 7, 3, "lex"),
    (6, 6, 8, 2, "blocked"),
    (6, 6, 7, 3, "lex"),
    (6, 4, 9, 4, "blocked"),
    (7, 6, 10, 3, "blocked"),
])
def test_schedule_reproduces_oracle_state(n, ne, kb, lb, order):
    F = uccsd_factors(n, ne, order)
    h = _lib.Handle(2 * n)
    h.set_option("tile_bits", kb)
    h.set_option("low_bits", lb)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    S = h.schedule(0)
    th = np.random.default_rng(5).normal(0, 0.2, F.n_theta)
    cs = [(np.cos(th[t]), np.sin(th[t])) for t in F.theta_index]
    psi = SI.to_logical(SI.run_forward(S, 2 * n, F.ref_index, cs), S["final_layout"])
    ref = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, EMPTY).state(th)
    assert abs(psi - ref).max() < 1e-14
    # every factor scheduled exactly once, in order
    slots = [f["slot"] for p in S["passes"] for f in p["factors"]]
    assert slots == list(range(len(F)))
    for p in S["passes"]:
        assert bin(p["act_mask"]).count("1") == S["tile_bits"]
        if S["tile_bits"] < 2 * n:
            assert p["act_mask"] & ((1 << lb) - 1) == (1 << lb) - 1     # low bits always active


def test_blocked_order_fuses_24_qubit_uccsd():
    F = uccsd_factors(12, 12, "blocked")
    assert len(F) == 1818
    h = _lib.Handle(24)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    S = h.schedule(0)
    assert len(S["passes"]) <= 48                           # ~40 factors per HBM pass on average
    F2 = uccsd_factors(12, 12, "lex")
    assert sorted(map(tuple, F.idx.tolist())) == sorted(map(tuple, F2.idx.tolist()))


def test_factor_counts():
    # n_theta = 2ov + 2C(o,2)C(v,2) + (ov)^2
    for n, cnt in ((4, 26), (8, 360)):
        assert len(uccsd_factors(n, n)) == cnt


@pytest.mark.parametrize("n,ne,kb,lb,ranks,order", [
    (5, 4, 6, 2, 2, "lex"),
    (6, 6, 7, 3, 2, "blocked"),
    (6, 6, 7, 2, 4, "lex"),
    (6, 4, 8, 3, 4, "blocked"),
    (7, 6, 8, 3, 8, "blocked"),
    (7, 6, 9, 3, 8, "lex"),
])
def test_sharded_schedule_reproduces_oracle_state(n, ne, kb, lb, ranks, order):
    """Rank bits are physical index bits like any other: a pass whose active set contains
    one straddles the ranks and its re-layout swaps the global qubit into local memory."""
    F = uccsd_factors(n, ne, order)
    h = _lib.ShardGroup(2 * n, [0] * ranks)
    h.set_option("tile_bits", kb)
    h.set_option("low_bits", lb)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    S = h.schedule(0)
    n_local = 2 * n - (ranks.bit_length() - 1)
    assert h.shard_info() == dict(rank=0, n_ranks=ranks, n_local_bits=n_local, members=ranks)
    th = np.random.default_rng(5).normal(0, 0.2, F.n_theta)
    cs = [(np.cos(th[t]), np.sin(th[t])) for t in F.theta_index]
    ref = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, EMPTY).state(th)
    shards = SI.run_forward_sharded(S, 2 * n, ranks, F.ref_index, cs)
    psi = _lib.assemble_state(shards, S["final_layout"])
    assert abs(psi - ref).max() < 1e-14
    slots = [f["slot"] for p in S["passes"] for f in p["factors"]]
    assert slots == list(range(len(F)))
    n_glob = sum(1 for p in S["passes"] if p["act_mask"] >> n_local)
    assert 0 < n_glob < len(S["passes"])            # some passes straddle ranks, not all


def test_sharded_schedule_global_pass_share():
    """34 qubits on 8 GPUs (BASELINE.json configs[4] shape): how many passes touch NVLink."""
    F = uccsd_factors(12, 12, "blocked")
    h = _lib.ShardGroup(24, [0] * 8)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    S = h.schedule(0)
    n_glob = sum(1 for p in S["passes"] if p["act_mask"] >> 21)
    assert len(S["passes"]) <= 60 and n_glob <= len(S["passes"]) // 2


@pytest.mark.parametrize("n,ne,kb,lb,order", [
    (4, 4, 12, 3, "blocked"),
    (4, 2, 6, 2, "lex"),
    (5, 4, 7, 3, "lex"),
    (6, 6, 8, 2, "blocked"),
    (6, 4, 9, 4, "blocked"),
    (7, 6, 10, 3, "blocked"),
    (8, 8, 12, 3, "blocked"),
])
def test_sparse_plan_reproduces_oracle_state(n, ne, kb, lb, order):
    """Compressed-sector tile plan (product-set slots, per-factor pair lists, re-layout offsets)
    executed on the host by the library = oracle state."""
    F = uccsd_factors(n, ne, order)
    h = _lib.Handle(2 * n)
    h.set_option("tile_bits", kb)
    h.set_option("grad_tile_bits", min(kb, 12))
    h.set_option("low_bits", lb)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    th = np.random.default_rng(5).normal(0, 0.2, F.n_theta)
    ref = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, EMPTY).state(th)
    for which in (0, 1):
        psi = h.sparse_plan_host_state(th, which)
        assert abs(psi - ref).max() < 1e-13


def test_sparse_plan_open_shell_generalized_factors():
    """N_alpha != N_beta reference, generalized singles/doubles: plan tables on the host = oracle."""
    from multiscalequantumcomputing_b200.adapt import generalized_pool
    from multiscalequantumcomputing_b200.uccsd import FactorList
    n, occ = 5, (0, 1, 2, 3, 4)
    nq = 2 * n
    kind, idx = generalized_pool(n)
    rng = np.random.default_rng(9)
    pick = rng.choice(len(kind), size=50, replace=False)
    ref_index = sum(1 << (nq - 1 - q) for q in occ)
    F = FactorList(nq, kind[pick], idx[pick], np.arange(50, dtype=np.int32), 50, ref_index)
    h = _lib.Handle(nq)
    h.set_option("tile_bits", 7)
    h.set_option("grad_tile_bits", 7)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    th = rng.normal(0, 0.3, F.n_theta)
    ref = UCCSDOracle(nq, F.as_tuples(), F.theta_index, F.ref_index, EMPTY).state(th)
    assert abs(h.sparse_plan_host_state(th, 0) - ref).max() < 1e-13
    # and the generic schedule through the numpy interpreter
    S = h.schedule(0)
    cs = [(np.cos(th[t]), np.sin(th[t])) for t in F.theta_index]
    psi = SI.to_logical(SI.run_forward(S, nq, F.ref_index, cs), S["final_layout"])
    assert abs(psi - ref).max() < 1e-13
