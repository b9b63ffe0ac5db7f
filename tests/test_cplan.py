"""Compact-sector sweep plan (csrc/mqc_cplan.h), executed on the host through the C ABI without a
GPU: row / column layouts per pass, tile classes, pair-list blobs and the row / column maps must
reproduce the oracle state.  k_csweep walks exactly these tables."""
import numpy as np
import pytest

from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.uccsd import uccsd_factors
from oracle.statevector import UCCSDOracle

EMPTY = (np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.complex128))


@pytest.mark.parametrize("n,ne,kb,order,which", [
    (4, 4, 14, "blocked", 0),     # whole register in one tile
    (4, 4, 6, "lex", 0),
    (5, 4, 7, "lex", 1),
    (6, 6, 8, "blocked", 0),
    (6, 6, 7, "lex", 1),
    (6, 4, 9, "blocked", 0),
    (6, 2, 8, "lex", 0),
    (7, 6, 10, "blocked", 1),
    (8, 8, 10, "blocked", 0),
])
def test_cplan_reproduces_oracle_state(n, ne, kb, order, which):
    """Both layouts of the class programs: 8-byte amplitudes (even row stride with spare slots, sixteen-lane bank
    order; the default) and 16-byte amplitudes (odd row stride)."""
    F = uccsd_factors(n, ne, order)
    th = np.random.default_rng(5).normal(0, 0.2, F.n_theta)
    ref = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, EMPTY).state(th)
    smem = {}
    for real in (1, 0):
        h = _lib.Handle(2 * n)
        h.set_option("tile_bits", kb)
        h.set_option("grad_tile_bits", kb)
        h.set_option("real_amplitudes", real)
        h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
        assert h.amplitude_bytes() == (8 if real else 16)
        info = h.cplan_info(which)
        assert info["n_passes"] == len(h.schedule(which)["passes"])
        from math import comb
        assert (info["n_rows"], info["n_cols"]) == (comb(n, ne // 2), comb(n, ne - ne // 2))
        psi = h.cplan_host_state(th, which)
        assert abs(psi - ref).max() < 1e-14
        smem[real] = info["smem_adjoint"]
        h.close()
    assert smem[1] <= smem[0]


def test_cplan_generalized_factors_and_tied_parameters():
    """de-excitation-like generalized factors (any i, a) and shared parameters"""
    n = 5
    rng = np.random.default_rng(11)
    kind, idx = [], []
    for _ in range(60):
        if rng.random() < 0.4:
            s = int(rng.integers(0, 2))
            p, q = rng.choice(n, 2, replace=False)
            kind.append(1); idx.append((2 * p + s, 2 * q + s, -1, -1))
        else:
            s1, s2 = int(rng.integers(0, 2)), int(rng.integers(0, 2))
            while True:
                p, q, r, t = (int(v) for v in rng.integers(0, n, 4))
                qs = {2 * p + s1, 2 * q + s2, 2 * r + s1, 2 * t + s2}
                if len(qs) == 4:
                    break
            kind.append(2); idx.append((2 * p + s1, 2 * q + s2, 2 * r + s1, 2 * t + s2))
    kind = np.array(kind, np.int32); idx = np.array(idx, np.int32)
    ti = (np.arange(len(kind)) // 3).astype(np.int32)
    nth = int(ti.max()) + 1
    ref_index = 0b1111000000
    h = _lib.Handle(2 * n)
    h.set_option("tile_bits", 6)
    h.set_ansatz(kind, idx, ti, nth, ref_index)
    assert h.cplan_info(0)["n_passes"] > 1
    th = rng.normal(0, 0.3, nth)
    tup = [(int(k), tuple(int(v) for v in (ix[:2] if k == 1 else ix))) for k, ix in zip(kind, idx)]
    ref = UCCSDOracle(2 * n, tup, ti, ref_index, EMPTY).state(th)
    assert abs(h.cplan_host_state(th, 0) - ref).max() < 1e-14
    h.close()


def test_no_compact_plan_for_non_conserving_ansatz():
    kind = np.array([1], np.int32); idx = np.array([[0, 3, -1, -1]], np.int32)      # alpha -> beta: S_z changes
    h = _lib.Handle(8)
    h.set_ansatz(kind, idx, np.zeros(1, np.int32), 1, 0b11110000)
    assert h.cplan_info(0)["n_passes"] == -1
    with pytest.raises(_lib.MqcError):
        h.cplan_host_state(np.zeros(1))
    h.close()


def test_24_qubit_plan_shape():
    F = uccsd_factors(12, 12, "blocked")
    h = _lib.Handle(24)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    i0, i1 = h.cplan_info(0), h.cplan_info(1)
    assert (i0["n_rows"], i0["n_cols"]) == (924, 924)
    assert i0["n_passes"] == len(h.schedule(0)["passes"]) and i1["n_passes"] == len(h.schedule(1)["passes"])
    assert i0["smem_forward"] < 64 * 1024 and i1["smem_adjoint"] < 100 * 1024       # several CTAs per SM
    h.close()


@pytest.mark.parametrize("n,ne", [(4, 3), (5, 5), (6, 3)])
def test_odd_electron_reference(n, ne):
    """noa = N // 2, nob = N - noa (`mqc/solver/dmet_interface.py:84-85`): open-shell reference
    determinant, UCCSD singles / doubles from it, compact plan against the oracle state."""
    from multiscalequantumcomputing_b200.uccsd import electron_split, occupied_qubits
    F = uccsd_factors(n, ne, "blocked")
    noa, nob = electron_split(ne)
    assert (noa, nob) == (ne // 2, ne - ne // 2) and nob == noa + 1
    occ = occupied_qubits(ne)
    assert sum(1 for q in occ if q % 2 == 0) == noa and sum(1 for q in occ if q % 2 == 1) == nob
    assert F.ref_index == sum(1 << (2 * n - 1 - q) for q in occ)
    for k, ix in zip(F.kind, F.idx):                 # every factor excites occupied -> virtual, spin conserved
        qs = [int(v) for v in (ix[:2] if k == 1 else ix)]
        src, dst = (qs[:1], qs[1:]) if k == 1 else (qs[:2], qs[2:])
        assert all(q in occ for q in src) and all(q not in occ for q in dst)
        assert sorted(q % 2 for q in src) == sorted(q % 2 for q in dst)
    h = _lib.Handle(2 * n)
    h.set_option("tile_bits", 6)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    from math import comb
    info = h.cplan_info(0)
    assert (info["n_rows"], info["n_cols"]) == (comb(n, noa), comb(n, nob))
    th = np.random.default_rng(3).normal(0, 0.2, F.n_theta)
    ref = UCCSDOracle(2 * n, F.as_tuples(), F.theta_index, F.ref_index, EMPTY).state(th)
    assert abs(h.cplan_host_state(th, 0) - ref).max() < 1e-14
    h.close()


def test_dressed_singles_and_scaled_tied_parameters():
    """kind 3 = n_c (a_a^+ a_i - h.c.) and per-factor angle scales (the building blocks of a spin-adapted pool
    operator in product form): compact plan on the host against the oracle, several passes."""
    n = 5
    rng = np.random.default_rng(21)
    kind, idx = [], []
    for _ in range(70):
        r = rng.random()
        s1, s2 = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        if r < 0.35:
            p, q = rng.choice(n, 2, replace=False)
            c = int(rng.integers(0, 2 * n))
            while c in (2 * p + s1, 2 * q + s1):
                c = int(rng.integers(0, 2 * n))
            kind.append(3); idx.append((2 * p + s1, 2 * q + s1, c, -1))
        elif r < 0.55:
            p, q = rng.choice(n, 2, replace=False)
            kind.append(1); idx.append((2 * p + s1, 2 * q + s1, -1, -1))
        else:
            while True:
                p, q, r2, t = (int(v) for v in rng.integers(0, n, 4))
                if len({2 * p + s1, 2 * q + s2, 2 * r2 + s1, 2 * t + s2}) == 4:
                    break
            kind.append(2); idx.append((2 * p + s1, 2 * q + s2, 2 * r2 + s1, 2 * t + s2))
    kind = np.array(kind, np.int32); idx = np.array(idx, np.int32)
    ti = (np.arange(len(kind)) // 4).astype(np.int32)
    nth = int(ti.max()) + 1
    scale = rng.normal(0, 1, len(kind))
    ref_index = 0b1111000000
    h = _lib.Handle(2 * n)
    h.set_option("tile_bits", 6)
    h.set_ansatz(kind, idx, ti, nth, ref_index, theta_scale=scale)
    assert h.cplan_info(0)["n_passes"] > 1
    th = rng.normal(0, 0.3, nth)
    tup = [(int(k), tuple(int(v) for v in ix[:{1: 2, 2: 4, 3: 3}[int(k)]])) for k, ix in zip(kind, idx)]
    ref = UCCSDOracle(2 * n, tup, ti, ref_index, EMPTY, scale).state(th)
    assert abs(h.cplan_host_state(th, 0) - ref).max() < 1e-14
    # a dressed single needs the compact storage
    hd = _lib.Handle(2 * n)
    hd.set_option("compact", 0)
    with pytest.raises(_lib.MqcError):
        hd.set_ansatz(kind, idx, ti, nth, ref_index)
    h.close(); hd.close()
