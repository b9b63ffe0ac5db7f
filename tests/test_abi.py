"""The C-ABI library loads, exports every symbol include/mqc_b200.h declares, validates its
arguments and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from multiscalequantumcomputing_b200 import _lib
from multiscalequantumcomputing_b200.uccsd import uccsd_factors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mqc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mqc_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    lib = _lib.load()
    names = _declared()
    assert "mqc_energy_grad" in names and "mqc_rdm12" in names
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.SIGNATURES) | {"mqc_last_error"}
    assert lib.mqc_version() >= 100


def test_argument_validation():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.mqc_create(1, 0, ctypes.byref(h)) != 0
    assert b"n_qubits" in lib.mqc_last_error()
    hd = _lib.Handle(8)
    with pytest.raises(_lib.MqcError):
        hd.set_option("no_such_option", 1)
    with pytest.raises(_lib.MqcError):
        hd.set_ansatz([2], [[0, 0, 4, 5]], [0], 1, 240)          # repeated orbital
    with pytest.raises(_lib.MqcError):
        hd.set_ansatz([1], [[0, 9, -1, -1]], [0], 1, 240)         # index outside register
    with pytest.raises(_lib.MqcError):
        hd.set_hamiltonian([1 << 9], [0], [1.0])                   # mask outside register


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    F = uccsd_factors(4, 4)
    hd = _lib.Handle(8)
    hd.set_hamiltonian([0], [0], [1.0])
    hd.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    with pytest.raises(_lib.MqcError, match="no CUDA device"):
        hd.energy(np.zeros(F.n_theta))
    with pytest.raises(_lib.MqcError, match="no CUDA device"):
        hd.energy_grad(np.zeros(F.n_theta))
    with pytest.raises(_lib.MqcError, match="no CUDA device"):
        hd.state(np.zeros(F.n_theta))


def test_schedule_options_are_frozen_by_set_ansatz():
    """Options that shape schedules / layouts / plans cannot change under an existing ansatz (they would
    silently desynchronise the launch path from the tables); launch-time options still can."""
    import numpy as np
    from multiscalequantumcomputing_b200.uccsd import uccsd_factors
    F = uccsd_factors(4, 4)
    h = _lib.Handle(8)
    h.set_option("tile_bits", 6)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    assert h.amplitude_bytes() == 8            # real compact storage is decided by mqc_set_ansatz (no Hamiltonian yet: real)
    for key in ("mode", "sector", "sparse_tiles", "sp2", "tile_bits", "grad_tile_bits", "low_bits", "compact", "real_amplitudes"):
        with pytest.raises(_lib.MqcError):
            h.set_option(key, 0 if key != "tile_bits" else 7)
    h.set_option("csweep_threads", 64)
    h.set_option("csweep_adj_threads", 32)
    h.set_option("sigma_stage", 0)
    h.set_option("sigma_split", 0)
    h.close()
    # the storage form is a set_ansatz-time choice: option, and the dense register for non-conserving / compact = 0
    h = _lib.Handle(8)
    h.set_option("real_amplitudes", 0)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    assert h.amplitude_bytes() == 16
    h.close()
    h = _lib.Handle(8)
    h.set_option("compact", 0)
    h.set_ansatz(F.kind, F.idx, F.theta_index, F.n_theta, F.ref_index)
    assert h.amplitude_bytes() == 0
    h.close()
