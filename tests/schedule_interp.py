"""numpy interpreter of the sweep schedule produced by the C ABI (tests only).

Mirrors the semantics of ``k_tile_sweep`` (csrc/mqc_kernels.cuh) pass by pass so that the
host-side scheduler — mask translation, tile-local coordinates, re-layout permutations —
can be validated against the oracle on a CPU-only box.  It is NOT a compute path of the
product; nothing in the package imports it."""
import numpy as np


def _deposit(t, pos):
    out = np.zeros_like(t)
    for j, p in enumerate(pos):
        out |= ((t >> np.uint64(j)) & np.uint64(1)) << np.uint64(p)
    return out


def _parity(v):
    v = v.copy()
    for s in (32, 16, 8, 4, 2, 1):
        v ^= v >> np.uint64(s)
    return (v & np.uint64(1)).astype(np.int64)


def run_forward(sched, n_bits, ref_index, cos_sin):
    """Execute all passes on a 2^n state that starts as |ref> in the identity layout.
    Returns the state in PHYSICAL order after the last pass."""
    dim = 1 << n_bits
    psi = np.zeros(dim, dtype=np.complex128)
    ref_phys = 0
    for l, p in enumerate(sched["initial_layout"]):
        ref_phys |= ((ref_index >> l) & 1) << int(p)
    psi[ref_phys] = 1.0
    K = sched["tile_bits"]
    T = 1 << K
    t = np.arange(T, dtype=np.uint64)
    for P in sched["passes"]:
        pos = [int(p) for p in P["pos"]]
        act = int(P["act_mask"])
        inactive = [b for b in range(n_bits) if not (act >> b) & 1]
        ntiles = 1 << (n_bits - K)
        tau = np.arange(ntiles, dtype=np.uint64)
        base = _deposit(tau, inactive)                        # (ntiles,)
        g = base[:, None] | _deposit(t, pos)[None, :]          # physical index of (tile, local)
        tile = psi[g.astype(np.int64)]                         # plain load
        for F in P["factors"]:
            m, v, z = np.uint64(F["m"]), np.uint64(F["v"]), np.uint64(F["z"])
            c, sn = cos_sin[F["slot"]]
            x0 = t[(t & m) == v]
            x1 = x0 ^ m
            par_in = _parity(x0 & z)                           # (npairs,)
            par_tile = (_parity(base & np.uint64(F["z_out"])) + F["sign0"]) & 1   # (ntiles,)
            s = 1.0 - 2.0 * ((par_in[None, :] + par_tile[:, None]) & 1)
            a = tile[:, x0.astype(np.int64)]
            b = tile[:, x1.astype(np.int64)]
            tile[:, x0.astype(np.int64)] = c * a - s * sn * b
            tile[:, x1.astype(np.int64)] = c * b + s * sn * a
        if P["has_perm"]:
            perm = [int(p) for p in P["perm"]]
            src = _deposit(t, perm).astype(np.int64)           # old local index for new local t'
            psi[g.astype(np.int64)] = tile[:, src]
        else:
            psi[g.astype(np.int64)] = tile
    return psi


def to_logical(psi_phys, layout):
    """out[logical index] = psi[physical index]; layout[l] = physical position of logical bit l."""
    n = len(layout)
    i = np.arange(1 << n, dtype=np.uint64)
    p = _deposit(i, [int(v) for v in layout])
    return psi_phys[p.astype(np.int64)]


def run_forward_sharded(sched, n_bits, n_ranks, ref_index, cos_sin):
    """The same passes executed the way the ranks of a sharded state do (csrc `launch_tile` +
    `peer_psi`): every rank sweeps the tiles of its tile map, an amplitude with physical index g
    lives in shard g >> n_local at offset g & (2^n_local - 1).  Also checks that the ranks'
    tiles partition the register in every pass.  Returns the list of shards."""
    g_bits = n_ranks.bit_length() - 1
    n_local = n_bits - g_bits
    lmask = (1 << n_local) - 1
    shards = [np.zeros(1 << n_local, dtype=np.complex128) for _ in range(n_ranks)]
    ref_phys = 0
    for l, p in enumerate(sched["initial_layout"]):
        ref_phys |= ((ref_index >> l) & 1) << int(p)
    shards[ref_phys >> n_local][ref_phys & lmask] = 1.0
    K = sched["tile_bits"]
    t = np.arange(1 << K, dtype=np.uint64)
    flat = np.concatenate(shards)                               # index = physical index (rank bits on top)
    for P in sched["passes"]:
        pos = [int(p) for p in P["pos"]]
        act = int(P["act_mask"])
        inactive_local = [b for b in range(n_local) if not (act >> b) & 1]
        touched = np.zeros(1 << n_bits, dtype=np.int32)
        new = flat.copy()
        for r in range(n_ranks):
            tau_hi, fixed, ntiles = P["tile_map"][r]
            assert ntiles == 1 << (n_local - K)
            tau = np.arange(ntiles, dtype=np.uint64) | np.uint64(tau_hi)
            base = _deposit(tau, inactive_local) | np.uint64(fixed)
            g = (base[:, None] | _deposit(t, pos)[None, :]).astype(np.int64)
            np.add.at(touched, g.ravel(), 1)
            tile = flat[g]
            for F in P["factors"]:
                m, v, z = np.uint64(F["m"]), np.uint64(F["v"]), np.uint64(F["z"])
                c, sn = cos_sin[F["slot"]]
                x0 = t[(t & m) == v]
                x1 = x0 ^ m
                par_in = _parity(x0 & z)
                par_tile = (_parity(base & np.uint64(F["z_out"])) + F["sign0"]) & 1
                s = 1.0 - 2.0 * ((par_in[None, :] + par_tile[:, None]) & 1)
                a = tile[:, x0.astype(np.int64)]
                b = tile[:, x1.astype(np.int64)]
                tile[:, x0.astype(np.int64)] = c * a - s * sn * b
                tile[:, x1.astype(np.int64)] = c * b + s * sn * a
            if P["has_perm"]:
                src = _deposit(t, [int(p) for p in P["perm"]]).astype(np.int64)
                new[g] = tile[:, src]
            else:
                new[g] = tile
        assert (touched == 1).all(), "the ranks' tiles must partition the register"
        flat = new
    return [flat[r << n_local:(r + 1) << n_local] for r in range(n_ranks)]
