"""Fragment-parallel multi-GPU dispatch (mode "fragment-parallel").

The fragments of one `MyDMET.doexact(mu)` are independent
(`/root/reference/mqc/third_party/qcdmet.py:42-146`: each loop iteration reads only the
mean-field 1-RDM, its own cluster and mu); the reference solves them sequentially.  Here the
loop body's solver call becomes "collect inputs -> solve one fragment per GPU -> scatter
results"; no data-path collective is needed, only the (E_imp, rdm1) results travel.

Two transports:
* ``solve_fragments``             one process, one worker thread per GPU (ctypes releases the
                                  GIL while a CUDA call runs; every handle owns its own stream);
* ``solve_fragments_distributed`` one process per GPU under ``torch.distributed`` (torchrun):
                                  rank r solves fragments r, r+W, ...; results are exchanged with
                                  ``all_gather_object`` (NCCL ranks use a gloo side group, CPU
                                  tests run it on gloo directly).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Sequence

from . import _lib
from .solver import solve


def _as_args(p):
    if hasattr(p, "oei"):
        return (p.oei, p.fock, p.tei, p.nelec, p.n_imp_orbs, p.chempot)
    return tuple(p)


def assign_round_robin(n_items: int, n_workers: int):
    """Item i -> worker i mod W (fragment f -> GPU f mod G)."""
    return [list(range(w, n_items, n_workers)) for w in range(n_workers)]


def solve_fragments(problems: Sequence, devices: Sequence[int] = None, solver: Callable = None, fragment_ids=None, **kw):
    """Solve every fragment problem, one per GPU at a time.  Returns [(E_imp, rdm1), ...] in input
    order.  ``solver(*args, device=d, **kw)`` defaults to the CUDA `solve`.

    ``fragment_ids``: ``True`` (problem i is fragment i) or a list of hashable ids: the solver keeps one
    GPU handle and one warm-start vector per fragment across calls (successive chemical-potential
    steps and DMET iterations re-use them, `solver.solve(fragment_id=...)`)."""
    solver = solver or solve
    if fragment_ids is True:
        fragment_ids = list(range(len(problems)))
    if devices is None:
        n = _lib.device_count()
        if n == 0:
            raise _lib.MqcError("no CUDA device available: the B200 kernels are required (there is no CPU fallback)")
        devices = list(range(n))
    devices = list(devices)[:max(1, len(problems))]
    plan = assign_round_robin(len(problems), len(devices))
    out = [None] * len(problems)

    def work(w):
        for i in plan[w]:
            extra = {} if fragment_ids is None else {"fragment_id": ("frag", fragment_ids[i])}
            out[i] = solver(*_as_args(problems[i]), device=devices[w], **extra, **kw)

    if len(devices) == 1:
        work(0)
    else:
        with ThreadPoolExecutor(max_workers=len(devices)) as ex:
            list(ex.map(work, range(len(devices))))
    return out


def solve_fragments_distributed(problems: Sequence, solver: Callable = None, device: int = None, group=None, **kw):
    """Collective call: every rank passes the same problem list and gets the full result list."""
    import torch.distributed as dist
    solver = solver or solve
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = assign_round_robin(len(problems), world)[rank]
    local = {}
    for i in mine:
        args = _as_args(problems[i])
        local[i] = solver(*args, device=device, **kw) if device is not None else solver(*args, **kw)
    gathered = [None] * world
    dist.all_gather_object(gathered, local, group=group)
    out = [None] * len(problems)
    for part in gathered:
        for i, v in part.items():
            out[i] = v
    return out
