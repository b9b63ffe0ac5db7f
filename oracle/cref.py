"""ORACLE (test infrastructure, not product code) — ctypes binding of oracle/c/uccsd_ref.c.

The C restatement is the fast CPU checker (20-24 qubits) and bench.py's CPU baseline
("kind": "port").  Validated against the numpy oracle in tests/test_oracle_c.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "liborc.so")
_lib = None

_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_c128p = C.c_void_p


def build(force=False):
    src = os.path.join(HERE, "c", "uccsd_ref.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(HERE, "c")], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        lib.orc_num_threads.restype = C.c_int
        lib.orc_apply_excitation.argtypes = [_c128p, C.c_int, C.c_int, _i32p, C.c_double]
        lib.orc_apply_excitation.restype = None
        lib.orc_adjoint_step.argtypes = [_c128p, _c128p, C.c_int, C.c_int, _i32p, C.c_double]
        lib.orc_adjoint_step.restype = C.c_double
        lib.orc_happly.argtypes = [_c128p, _c128p, C.c_int, C.c_int64, _u64p, _u64p, _f64p, _f64p, C.c_int64, C.c_int64]
        lib.orc_happly.restype = None
        lib.orc_reference_state.argtypes = [_c128p, C.c_int, C.c_uint64]
        lib.orc_reference_state.restype = None
        lib.orc_energy_grad.argtypes = [_c128p, _c128p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _f64p, C.c_uint64,
                                        C.c_int64, _u64p, _u64p, _f64p, _f64p, _f64p, C.c_int, C.c_int]
        lib.orc_energy_grad.restype = C.c_double
        lib.orc_sector_list.argtypes = [C.c_int, C.c_int, C.c_int, _u64p]
        lib.orc_sector_list.restype = C.c_int64
        lib.orc_energy_grad_sector.argtypes = [_c128p, _c128p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _f64p, C.c_uint64,
                                               C.c_int64, _u64p, _u64p, _f64p, _f64p, _f64p, C.c_int, C.c_int,
                                               _u64p, C.c_int64, _f64p]
        lib.orc_energy_grad_sector.restype = C.c_double
        _lib = lib
    return _lib


def num_threads():
    return load().orc_num_threads()


def use_all_cores():
    """The CPU baseline uses every host core, also under torchrun (which exports OMP_NUM_THREADS=1)."""
    import os
    lib = load()
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    lib.orc_set_num_threads(C.c_int(n))
    return lib.orc_num_threads()


class CRefUCCSD:
    """Same role as oracle.statevector.UCCSDOracle, in C."""

    def __init__(self, n_qubits, kind, idx, theta_index, n_theta, ref_index, ham_terms):
        self.lib = load()
        self.n = n_qubits
        self.kind = np.ascontiguousarray(kind, dtype=np.int32)
        self.idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1, 4)
        self.ti = np.ascontiguousarray(theta_index, dtype=np.int32)
        self.n_theta = int(n_theta)
        self.ref_index = int(ref_index)
        xm, zm, cf = ham_terms
        self.xm = np.ascontiguousarray(xm, dtype=np.uint64)
        self.zm = np.ascontiguousarray(zm, dtype=np.uint64)
        cf = np.asarray(cf, dtype=np.complex128)
        self.cre = np.ascontiguousarray(cf.real)
        self.cim = np.ascontiguousarray(cf.imag)
        self.psi = np.zeros(1 << n_qubits, dtype=np.complex128)
        self.lam = np.zeros(1 << n_qubits, dtype=np.complex128)

    def _p(self, a, t):
        return a.ctypes.data_as(t)

    def energy_grad(self, theta, want_grad=True):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        g = np.zeros(max(1, self.n_theta))
        e = self.lib.orc_energy_grad(self.psi.ctypes.data, self.lam.ctypes.data, self.n, len(self.kind),
                                     self._p(self.kind, _i32p), self._p(self.idx, _i32p), self._p(self.ti, _i32p),
                                     self._p(th, _f64p), self.ref_index, len(self.xm), self._p(self.xm, _u64p),
                                     self._p(self.zm, _u64p), self._p(self.cre, _f64p), self._p(self.cim, _f64p),
                                     self._p(g, _f64p), self.n_theta, 1 if want_grad else 0)
        return e, g[:self.n_theta]

    def energy(self, theta):
        return self.energy_grad(theta, False)[0]

    def sector_list(self):
        """Ascending basis indices of the (N_alpha, N_beta) sector of the reference determinant
        (alpha = even qubits, beta = odd qubits)."""
        if getattr(self, "_sector", None) is None:
            na = sum((self.ref_index >> (self.n - 1 - q)) & 1 for q in range(0, self.n, 2))
            nb = sum((self.ref_index >> (self.n - 1 - q)) & 1 for q in range(1, self.n, 2))
            cnt = self.lib.orc_sector_list(self.n, na, nb, None)
            lst = np.zeros(cnt, dtype=np.uint64)
            self.lib.orc_sector_list(self.n, na, nb, self._p(lst, _u64p))
            self._sector = lst
        return self._sector

    def energy_grad_sector(self, theta, want_grad=True):
        """One FULL evaluation with the sector-aware routines (sweeps over the sector's
        determinants only, H applied X-group by X-group).  Returns (E, grad, phase seconds)."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        g = np.zeros(max(1, self.n_theta))
        lst = self.sector_list()
        ph = np.zeros(3)
        e = self.lib.orc_energy_grad_sector(self.psi.ctypes.data, self.lam.ctypes.data, self.n, len(self.kind),
                                            self._p(self.kind, _i32p), self._p(self.idx, _i32p), self._p(self.ti, _i32p),
                                            self._p(th, _f64p), self.ref_index, len(self.xm), self._p(self.xm, _u64p),
                                            self._p(self.zm, _u64p), self._p(self.cre, _f64p), self._p(self.cim, _f64p),
                                            self._p(g, _f64p), self.n_theta, 1 if want_grad else 0,
                                            self._p(lst, _u64p), len(lst), self._p(ph, _f64p))
        return e, g[:self.n_theta], ph

    def state(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        self.lib.orc_reference_state(self.psi.ctypes.data, self.n, self.ref_index)
        for k in range(len(self.kind)):
            self.lib.orc_apply_excitation(self.psi.ctypes.data, self.n, int(self.kind[k]),
                                          self.idx[k].ctypes.data_as(_i32p), float(th[self.ti[k]]))
        return self.psi.copy()

    # pieces, for bounded-sample timing
    def time_sample(self, theta, n_factors_sample, n_terms_sample, sector_index=None):
        """Time a bounded sample: forward over the first ``n_factors_sample`` factors, H apply
        over the first ``n_terms_sample`` terms, reverse over the same factors.
        Returns (t_forward, t_happly, t_reverse) seconds."""
        import time
        th = np.ascontiguousarray(theta, dtype=np.float64)
        nf = min(n_factors_sample, len(self.kind))
        nt = min(n_terms_sample, len(self.xm))
        self.lib.orc_reference_state(self.psi.ctypes.data, self.n, self.ref_index)
        t0 = time.perf_counter()
        for k in range(nf):
            self.lib.orc_apply_excitation(self.psi.ctypes.data, self.n, int(self.kind[k]),
                                          self.idx[k].ctypes.data_as(_i32p), float(th[self.ti[k]]))
        t_fwd = time.perf_counter() - t0
        # after a few factors the state is still almost a single determinant; fill the whole
        # (N_alpha, N_beta) sector so that H apply and the reverse sweep see the populated state
        # of a full evaluation (untimed)
        if sector_index is not None:
            rng = np.random.default_rng(7)
            v = rng.standard_normal(len(sector_index))
            self.psi[:] = 0.0
            self.psi[sector_index] = v / np.linalg.norm(v)
        t1 = time.perf_counter()
        self.lib.orc_happly(self.psi.ctypes.data, self.lam.ctypes.data, self.n, len(self.xm), self._p(self.xm, _u64p),
                            self._p(self.zm, _u64p), self._p(self.cre, _f64p), self._p(self.cim, _f64p), 0, nt)
        t2 = time.perf_counter()
        for k in range(nf - 1, -1, -1):
            self.lib.orc_adjoint_step(self.psi.ctypes.data, self.lam.ctypes.data, self.n, int(self.kind[k]),
                                      self.idx[k].ctypes.data_as(_i32p), float(th[self.ti[k]]))
        t3 = time.perf_counter()
        return t_fwd, t2 - t1, t3 - t2, nf, nt
