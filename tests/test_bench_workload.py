"""bench.py measures BASELINE.json's metric on BASELINE.json's configuration: the workload
builder must produce configs[2] (24 qubits, 12 orbitals, 12 electrons, full UCCSD)."""
import json
import os

import numpy as np

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_is_baseline_config_2():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "24 qubits (12 orbitals, 12 e" in base["configs"][2]
    table, factors, theta, n_groups = bench.workload()
    assert factors.n_qubits == 24 and len(factors) == 1818 and factors.n_theta == 1818
    assert factors.ref_index == (1 << 24) - (1 << 12)                                        # 12 electrons in qubits 0..11
    x, z, c = table
    assert len(x) == len(z) == len(c) == 15169 and n_groups == 2767                          # H12 chain STO-3G + 0.25 S^2
    assert np.abs(np.asarray(c).imag).max() == 0.0
    assert theta.shape == (1818,) and abs(theta.std() - 0.05) < 0.005
    assert bench.CONFIG["n_qubits"] == 24 and "24q" in bench.CONFIG["workload"]
    # every factor conserves N_alpha and N_beta: the sector kernels are the measured path
    k1 = factors.idx[factors.kind == 1]
    assert ((k1[:, 0] ^ k1[:, 1]) & 1).max() == 0
