"""ctypes binding of the C ABI in include/mqc_b200.h.

The library is the only compute path: if ``libmqc_b200.so`` is missing, or no CUDA device is
visible when a compute entry point is called, this raises — there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_H = C.c_void_p

# name -> (argtypes); every function returns int except mqc_last_error
SIGNATURES = {
    "mqc_version": [],
    "mqc_device_count": [_i32p],
    "mqc_create": [C.c_int, C.c_int, C.POINTER(_H)],
    "mqc_shard_group_create": [C.c_int, C.c_int, _i32p, C.POINTER(_H)],
    "mqc_shard_create": [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_H)],
    "mqc_shard_export": [_H, C.c_int32, C.c_void_p, C.c_int64],
    "mqc_shard_connect": [_H, C.c_void_p, C.c_int64],
    "mqc_shard_info": [_H, _i32p, _i32p, _i32p, _i32p],
    "mqc_shard_read": [_H, _f64p, C.c_int32, _f64p, C.c_int64, _i32p],
    "mqc_schedule_tile_map": [_H, C.c_int32, C.c_int32, C.c_int32, _u64p, _u64p, _u64p],
    "mqc_destroy": [_H],
    "mqc_set_option": [_H, C.c_char_p, C.c_int64],
    "mqc_set_hamiltonian": [_H, C.c_int64, _u64p, _u64p, _f64p, _f64p],
    "mqc_set_ansatz": [_H, C.c_int32, _i32p, _i32p, _i32p, C.c_int32, C.c_uint64],
    "mqc_set_ansatz_scaled": [_H, C.c_int32, _i32p, _i32p, _i32p, _f64p, C.c_int32, C.c_uint64],
    "mqc_update_hamiltonian_coeffs": [_H, C.c_int64, _f64p, _f64p],
    "mqc_energy": [_H, _f64p, _f64p],
    "mqc_energy_grad": [_H, _f64p, _f64p, _f64p],
    "mqc_energy_grad_resident": [_H, _f64p],
    "mqc_event_record": [_H, C.c_int32],
    "mqc_event_elapsed_ms": [_H, _f64p],
    "mqc_prepare": [_H, _f64p],
    "mqc_state": [_H, _f64p, _f64p, C.c_int64],
    "mqc_expectation": [_H, C.c_int64, _u64p, _u64p, _f64p, _f64p, _f64p, _f64p],
    "mqc_rdm12": [_H, C.c_int32, C.c_int32, C.c_int32, _f64p, _f64p],
    "mqc_last_timing": [_H, _f64p, _i64p],
    "mqc_sparse_plan_host_state": [_H, C.c_int32, _f64p, _f64p, C.c_int64],
    "mqc_cplan_host_state": [_H, C.c_int32, _f64p, _f64p, C.c_int64],
    "mqc_cplan_info": [_H, C.c_int32, _i32p, _i64p, _i64p, _i64p, _i64p],
    "mqc_amplitude_bytes": [_H, _i32p],
    "mqc_load_state": [_H, _f64p, C.c_int64],
    "mqc_sigma_host_apply": [C.c_int32, C.c_int32, C.c_int32, C.c_int64, _u64p, _u64p, _f64p, _f64p, _f64p, _f64p,
                             _i64p, _i64p, _i64p],
    "mqc_sigma_ab_info": [C.c_int32, C.c_int32, C.c_int32, C.c_int64, _u64p, _u64p, _f64p, _f64p, _i64p, _i64p],
    "mqc_schedule_info": [_H, C.c_int32, _i32p, _i32p, _i32p],
    "mqc_schedule_pass": [_H, C.c_int32, C.c_int32, _i32p, _i32p, _u64p, _i32p, _i32p, _i32p],
    "mqc_schedule_factor": [_H, C.c_int32, C.c_int32, _u32p, _u32p, _u32p, _u64p, _i32p, _i32p],
    "mqc_schedule_final_layout": [_H, C.c_int32, _i32p],
    "mqc_schedule_initial_layout": [_H, C.c_int32, _i32p],
}


class MqcError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building in-tree first if the sources are newer) and type the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.needs_build():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on the box and no prebuilt library
            if not os.path.exists(path):
                raise MqcError("CUDA extension libmqc_b200.so is missing and could not be built: %s" % exc)
            import warnings
            warnings.warn("libmqc_b200.so is OLDER than its sources and the rebuild failed (%s): running the stale binary"
                          % str(exc).splitlines()[0])
    lib = C.CDLL(path)
    lib.mqc_last_error.restype = C.c_char_p
    lib.mqc_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise MqcError(load().mqc_last_error().decode())


def ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def device_count() -> int:
    n = C.c_int32(0)
    check(load().mqc_device_count(C.byref(n)))
    return n.value


def sigma_host_apply(n_qubits, n_alpha, n_beta, table, c_sector=None):
    """Host-only evaluation of the link tables (see include/mqc_b200.h).  Returns
    (L or None, (n_rows, n_cols), n_fallback_terms)."""
    lib = load()
    x = np.ascontiguousarray(table[0], dtype=np.uint64)
    z = np.ascontiguousarray(table[1], dtype=np.uint64)
    c = np.asarray(table[2], dtype=np.complex128)
    cr, ci = np.ascontiguousarray(c.real), np.ascontiguousarray(c.imag)
    nr, nc, nf = C.c_int64(), C.c_int64(), C.c_int64()
    check(lib.mqc_sigma_host_apply(n_qubits, n_alpha, n_beta, len(x), ptr(x, _u64p), ptr(z, _u64p), ptr(cr, _f64p),
                                   ptr(ci, _f64p), None, None, C.byref(nr), C.byref(nc), C.byref(nf)))
    out = None
    if c_sector is not None:
        cs = np.ascontiguousarray(c_sector, dtype=np.complex128)
        assert cs.shape == (nr.value, nc.value)
        out = np.zeros_like(cs)
        check(lib.mqc_sigma_host_apply(n_qubits, n_alpha, n_beta, len(x), ptr(x, _u64p), ptr(z, _u64p), ptr(cr, _f64p),
                                       ptr(ci, _f64p), C.cast(cs.ctypes.data, _f64p), C.cast(out.ctypes.data, _f64p),
                                       C.byref(nr), C.byref(nc), C.byref(nf)))
    return out, (nr.value, nc.value), nf.value


def sigma_ab_info(n_qubits, n_alpha, n_beta, table):
    """(Pauli terms taken over by the opposite-spin doubles kernel, non-zero weights W[ij][kl])."""
    lib = load()
    x = np.ascontiguousarray(table[0], dtype=np.uint64)
    z = np.ascontiguousarray(table[1], dtype=np.uint64)
    c = np.asarray(table[2], dtype=np.complex128)
    cr, ci = np.ascontiguousarray(c.real), np.ascontiguousarray(c.imag)
    nt, nw = C.c_int64(), C.c_int64()
    check(lib.mqc_sigma_ab_info(n_qubits, n_alpha, n_beta, len(x), ptr(x, _u64p), ptr(z, _u64p), ptr(cr, _f64p), ptr(ci, _f64p),
                                C.byref(nt), C.byref(nw)))
    return nt.value, nw.value


class Handle:
    """Thin RAII wrapper over ``mqc_solver_t*``."""

    def __init__(self, n_qubits: int, device: int = 0, _create=True):
        self.lib = load()
        self.n_qubits = n_qubits
        self.h = _H()
        if _create:
            check(self.lib.mqc_create(n_qubits, device, C.byref(self.h)))
        self.n_theta = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.mqc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        check(self.lib.mqc_set_option(self.h, key.encode(), int(value)))

    def set_hamiltonian(self, xmask, zmask, coeff):
        x = np.ascontiguousarray(xmask, dtype=np.uint64)
        z = np.ascontiguousarray(zmask, dtype=np.uint64)
        c = np.asarray(coeff, dtype=np.complex128)
        cr = np.ascontiguousarray(c.real)
        ci = np.ascontiguousarray(c.imag)
        check(self.lib.mqc_set_hamiltonian(self.h, len(x), ptr(x, _u64p), ptr(z, _u64p), ptr(cr, _f64p), ptr(ci, _f64p)))

    def update_hamiltonian_coeffs(self, coeff):
        """New coefficients for the Pauli strings of the last set_hamiltonian (same order)."""
        c = np.asarray(coeff, dtype=np.complex128)
        cr = np.ascontiguousarray(c.real)
        ci = np.ascontiguousarray(c.imag)
        check(self.lib.mqc_update_hamiltonian_coeffs(self.h, len(c), ptr(cr, _f64p), ptr(ci, _f64p)))

    def set_ansatz(self, kind, idx, theta_index, n_theta, ref_index, theta_scale=None):
        k = np.ascontiguousarray(kind, dtype=np.int32)
        ix = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1, 4)
        ti = np.ascontiguousarray(theta_index, dtype=np.int32)
        if theta_scale is None:
            check(self.lib.mqc_set_ansatz(self.h, len(k), ptr(k, _i32p), ptr(ix, _i32p), ptr(ti, _i32p), int(n_theta), int(ref_index)))
        else:
            sc = np.ascontiguousarray(theta_scale, dtype=np.float64)
            if sc.shape != (len(k),):
                raise ValueError("theta_scale must have one entry per factor")
            check(self.lib.mqc_set_ansatz_scaled(self.h, len(k), ptr(k, _i32p), ptr(ix, _i32p), ptr(ti, _i32p), ptr(sc, _f64p),
                                                 int(n_theta), int(ref_index)))
        self.n_theta = int(n_theta)

    def _theta(self, theta):
        t = np.ascontiguousarray(theta, dtype=np.float64)
        if t.shape != (self.n_theta,):
            raise ValueError("theta must have shape (%d,)" % self.n_theta)
        return t

    def energy(self, theta) -> float:
        t = self._theta(theta)
        e = C.c_double(0.0)
        check(self.lib.mqc_energy(self.h, ptr(t, _f64p), C.byref(e)))
        return e.value

    def energy_grad(self, theta):
        t = self._theta(theta)
        e = C.c_double(0.0)
        g = np.zeros(max(1, self.n_theta))
        check(self.lib.mqc_energy_grad(self.h, ptr(t, _f64p), C.byref(e), ptr(g, _f64p)))
        return e.value, g[:self.n_theta]

    def energy_grad_resident(self) -> float:
        e = C.c_double(0.0)
        check(self.lib.mqc_energy_grad_resident(self.h, C.byref(e)))
        return e.value

    def energy_grad_into(self, theta: np.ndarray, grad_out: np.ndarray) -> float:
        """Host buffers supplied by the caller (e.g. pinned memory): no allocation per call."""
        e = C.c_double(0.0)
        check(self.lib.mqc_energy_grad(self.h, ptr(theta, _f64p), C.byref(e), ptr(grad_out, _f64p)))
        return e.value

    def event_record(self, which: int):
        check(self.lib.mqc_event_record(self.h, which))

    def event_elapsed_ms(self) -> float:
        ms = C.c_double(0.0)
        check(self.lib.mqc_event_elapsed_ms(self.h, C.byref(ms)))
        return ms.value

    def prepare(self, theta):
        t = self._theta(theta)
        check(self.lib.mqc_prepare(self.h, ptr(t, _f64p)))

    def state(self, theta=None) -> np.ndarray:
        out = np.empty(1 << self.n_qubits, dtype=np.complex128)
        tp = ptr(self._theta(theta), _f64p) if theta is not None else None
        check(self.lib.mqc_state(self.h, tp, C.cast(out.ctypes.data, _f64p), out.shape[0]))
        return out

    def sparse_plan_host_state(self, theta, which=0) -> np.ndarray:
        """CPU execution of the compressed-sector tile plan (test of the plan tables)."""
        out = np.empty(1 << self.n_qubits, dtype=np.complex128)
        t = self._theta(theta)
        check(self.lib.mqc_sparse_plan_host_state(self.h, which, ptr(t, _f64p), C.cast(out.ctypes.data, _f64p), out.shape[0]))
        return out

    def cplan_host_state(self, theta, which=0) -> np.ndarray:
        """CPU execution of the compact-sector plan (test of the tables k_csweep walks)."""
        out = np.empty(1 << self.n_qubits, dtype=np.complex128)
        t = self._theta(theta)
        check(self.lib.mqc_cplan_host_state(self.h, which, ptr(t, _f64p), C.cast(out.ctypes.data, _f64p), out.shape[0]))
        return out

    def cplan_info(self, which=0):
        npass = C.c_int32()
        nr, nc, sf, sa = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self.lib.mqc_cplan_info(self.h, which, C.byref(npass), C.byref(nr), C.byref(nc), C.byref(sf), C.byref(sa)))
        return dict(n_passes=npass.value, n_rows=nr.value, n_cols=nc.value, smem_forward=sf.value, smem_adjoint=sa.value)

    def amplitude_bytes(self) -> int:
        """8: real compact storage, 16: complex compact storage, 0: dense register."""
        b = C.c_int32()
        check(self.lib.mqc_amplitude_bytes(self.h, C.byref(b)))
        return b.value

    def load_state(self, psi):
        """Make ``psi`` (OpenFermion order, 2^N complex amplitudes) the prepared state."""
        v = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1)
        check(self.lib.mqc_load_state(self.h, C.cast(v.ctypes.data, _f64p), v.shape[0]))

    def expectation(self, xmask, zmask, coeff) -> complex:
        x = np.ascontiguousarray(xmask, dtype=np.uint64)
        z = np.ascontiguousarray(zmask, dtype=np.uint64)
        c = np.asarray(coeff, dtype=np.complex128)
        cr = np.ascontiguousarray(c.real)
        ci = np.ascontiguousarray(c.imag)
        vr, vi = C.c_double(0.0), C.c_double(0.0)
        check(self.lib.mqc_expectation(self.h, len(x), ptr(x, _u64p), ptr(z, _u64p), ptr(cr, _f64p), ptr(ci, _f64p),
                                       C.byref(vr), C.byref(vi)))
        return complex(vr.value, vi.value)

    def rdm12(self, n_orb, n_alpha, n_beta, want_rdm2=True):
        r1 = np.zeros((n_orb, n_orb))
        r2 = np.zeros((n_orb,) * 4) if want_rdm2 else None
        check(self.lib.mqc_rdm12(self.h, n_orb, n_alpha, n_beta, ptr(r1, _f64p),
                                 ptr(r2, _f64p) if want_rdm2 else None))
        return r1, r2

    def last_timing(self):
        ms = np.zeros(4)
        cnt = np.zeros(3, dtype=np.int64)
        check(self.lib.mqc_last_timing(self.h, ptr(ms, _f64p), ptr(cnt, _i64p)))
        return dict(forward_ms=ms[0], hamiltonian_ms=ms[1], reverse_ms=ms[2], total_ms=ms[3],
                    launches=int(cnt[0]), forward_passes=int(cnt[1]), reverse_passes=int(cnt[2]))

    # ---- host-only schedule introspection -------------------------------------------
    def schedule(self, which=0):
        npass, kb, ntf = C.c_int32(), C.c_int32(), C.c_int32()
        check(self.lib.mqc_schedule_info(self.h, which, C.byref(npass), C.byref(kb), C.byref(ntf)))
        passes = []
        for p in range(npass.value):
            f0, nf, hp = C.c_int32(), C.c_int32(), C.c_int32()
            act = C.c_uint64()
            pos = np.zeros(16, dtype=np.int32)
            perm = np.zeros(16, dtype=np.int32)
            check(self.lib.mqc_schedule_pass(self.h, which, p, C.byref(f0), C.byref(nf), C.byref(act),
                                             ptr(pos, _i32p), ptr(perm, _i32p), C.byref(hp)))
            facs = []
            for t in range(f0.value, f0.value + nf.value):
                m, v, z = C.c_uint32(), C.c_uint32(), C.c_uint32()
                zo = C.c_uint64()
                s0, slot = C.c_int32(), C.c_int32()
                check(self.lib.mqc_schedule_factor(self.h, which, t, C.byref(m), C.byref(v), C.byref(z),
                                                   C.byref(zo), C.byref(s0), C.byref(slot)))
                facs.append(dict(m=m.value, v=v.value, z=z.value, z_out=zo.value, sign0=s0.value, slot=slot.value))
            tile_map = []
            for r in range(getattr(self, "n_ranks", 1)):
                th, fb, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
                check(self.lib.mqc_schedule_tile_map(self.h, which, p, r, C.byref(th), C.byref(fb), C.byref(nt)))
                tile_map.append((th.value, fb.value, nt.value))
            passes.append(dict(tile_map=tile_map, act_mask=act.value, pos=pos[:kb.value].copy(), perm=perm[:kb.value].copy(),
                               has_perm=bool(hp.value), factors=facs))
        lay = np.zeros(self.n_qubits, dtype=np.int32)
        check(self.lib.mqc_schedule_final_layout(self.h, which, ptr(lay, _i32p)))
        lay0 = np.zeros(self.n_qubits, dtype=np.int32)
        check(self.lib.mqc_schedule_initial_layout(self.h, which, ptr(lay0, _i32p)))
        return dict(tile_bits=kb.value, passes=passes, final_layout=lay, initial_layout=lay0)


BLOB_BYTES = 512


class _ShardBase(Handle):
    def shard_info(self):
        r, n, nl, m = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(self.lib.mqc_shard_info(self.h, C.byref(r), C.byref(n), C.byref(nl), C.byref(m)))
        return dict(rank=r.value, n_ranks=n.value, n_local_bits=nl.value, members=m.value)

    def state(self, theta=None):
        raise MqcError("a sharded state is read shard by shard: use read_shard() / gather_state()")

    def read_shard(self, member=0, theta=None):
        """(raw physical-order shard, phys_of_logical layout)"""
        nl = self.shard_info()["n_local_bits"]
        out = np.empty(1 << nl, dtype=np.complex128)
        lay = np.zeros(self.n_qubits, dtype=np.int32)
        tp = ptr(self._theta(theta), _f64p) if theta is not None else None
        check(self.lib.mqc_shard_read(self.h, tp, member, C.cast(out.ctypes.data, _f64p), out.shape[0], ptr(lay, _i32p)))
        return out, lay


def assemble_state(shards, layout):
    """Logical-order state vector from the physical-order shards of all ranks (rank = top
    physical index bits); small registers only (tests, checkpoints)."""
    phys = np.concatenate(shards)
    n = len(layout)
    i = np.arange(1 << n, dtype=np.uint64)
    p = np.zeros_like(i)
    for l, pos in enumerate(layout):
        p |= ((i >> np.uint64(l)) & np.uint64(1)) << np.uint64(int(pos))
    return phys[p.astype(np.int64)]


class ShardGroup(_ShardBase):
    """All ranks of a sharded state driven by this process (`mqc_shard_group_create`).
    ``devices[r]`` is the GPU of rank r; repeating a device puts several ranks on one GPU."""

    def __init__(self, n_qubits: int, devices):
        super().__init__(n_qubits, _create=False)
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        self.n_ranks = len(dv)
        check(self.lib.mqc_shard_group_create(n_qubits, len(dv), ptr(dv, _i32p), C.byref(self.h)))

    def gather_state(self, theta=None):
        shards = []
        lay = None
        for m in range(self.n_ranks):
            sh, lay = self.read_shard(m, theta if m == 0 else None)
            shards.append(sh)
        return assemble_state(shards, lay)


class ShardRank(_ShardBase):
    """One rank of a sharded state, one process per GPU (`mqc_shard_create`); ``connect``
    all-gathers the CUDA IPC blobs over ``torch.distributed`` (any backend)."""

    def __init__(self, n_qubits: int, device: int, rank: int, n_ranks: int):
        super().__init__(n_qubits, _create=False)
        self.rank, self.n_ranks = rank, n_ranks
        check(self.lib.mqc_shard_create(n_qubits, device, rank, n_ranks, C.byref(self.h)))

    def export(self, want_lambda=True) -> bytes:
        buf = C.create_string_buffer(BLOB_BYTES)
        check(self.lib.mqc_shard_export(self.h, int(bool(want_lambda)), buf, BLOB_BYTES))
        return buf.raw

    def connect_blobs(self, blobs):
        raw = b"".join(blobs)
        check(self.lib.mqc_shard_connect(self.h, raw, len(raw)))

    def connect(self, want_lambda=True, group=None):
        import torch.distributed as dist
        mine = self.export(want_lambda)
        blobs = [None] * self.n_ranks
        dist.all_gather_object(blobs, mine, group=group)
        self.connect_blobs(blobs)
        dist.barrier(group=group)

    def rdm12(self, n_orb, n_alpha, n_beta, want_rdm2=True, group=None):
        """Partial sums of this rank, all-reduced over the ranks."""
        import torch
        import torch.distributed as dist
        r1, r2 = super().rdm12(n_orb, n_alpha, n_beta, want_rdm2)
        for a in (r1, r2):
            if a is not None and self.n_ranks > 1:
                t = torch.from_numpy(a)
                if dist.get_backend(group) == "nccl":
                    t = t.cuda()
                    dist.all_reduce(t, group=group)
                    a[...] = t.cpu().numpy()
                else:
                    dist.all_reduce(t, group=group)
        return r1, r2
