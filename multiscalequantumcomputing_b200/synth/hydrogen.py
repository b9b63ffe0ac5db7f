"""Synthetic H_n / STO-3G inputs for the fragment-solver hot path (host side, numpy only).

This module is the *input generator* on the caller's side of the hot path: it
produces the integrals that the reference obtains from pyscf
(`/root/reference/mqc/system/fragment.py:49-63` ``run_pyscf_rhf``) for the two
geometries its tests use

* ``hydrogen_chain`` — `/root/reference/mqc/system/structure.py:17-23`
* ``hydrogen_ring``  — `/root/reference/mqc/system/structure.py:8-14`

pyscf is not available in this image, so the s-type Gaussian integrals are
written out in closed form.  STO-3G exponents and
contraction coefficients for H and the Bohr radius are the ones printed in the
reference's own molden dump (`/root/reference/test/tmp.molden:5,16-19`).

Pinned by tests/test_synth_inputs.py against the RHF energies in the reference
logs (`test/log_hchain_vqe:2`, `test/atestlog_vqe:2`, `test/log_hchain_fci`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy.special import erf

BOHR = 0.52917721092  # Angstrom; test/tmp.molden:5 (3.0 A = 5.66917837369518 a.u.)

# H STO-3G; test/tmp.molden:16-19
_STO3G_EXP = np.array([3.42525091, 0.62391373, 0.1688554])
_STO3G_COEF = np.array([0.15432897, 0.53532814, 0.44463454])


def hydrogen_chain(natom: int = 10, bondlength: float = 1.0):
    """Same geometry list as structure.py:17-23 (atoms on the z axis)."""
    return [("H", (0.0, 0.0, i * bondlength)) for i in range(natom)]


def hydrogen_ring(natom: int = 10, bondlength: float = 1.0):
    """Same geometry list as structure.py:8-14 (regular polygon in the xy plane)."""
    r = 0.5 * bondlength / np.sin(np.pi / natom)
    out = []
    for i in range(natom):
        theta = i * (2 * np.pi / natom)
        out.append(("H", (r * np.cos(theta), r * np.sin(theta), 0.0)))
    return out


def _boys0(t):
    """F0(t) = 1/2 sqrt(pi/t) erf(sqrt t), with the t->0 series."""
    t = np.asarray(t, dtype=np.float64)
    small = t < 1e-12
    ts = np.where(small, 1.0, t)
    val = 0.5 * np.sqrt(np.pi / ts) * erf(np.sqrt(ts))
    return np.where(small, 1.0 - t / 3.0, val)


@dataclass
class AOIntegrals:
    coords: np.ndarray   # (natom,3) Bohr
    S: np.ndarray        # overlap (nao,nao)
    T: np.ndarray        # kinetic
    V: np.ndarray        # nuclear attraction
    eri: np.ndarray      # (pq|rs) chemist notation, (nao,)*4
    e_nuc: float

    @property
    def hcore(self):
        return self.T + self.V

    @property
    def nao(self):
        return self.S.shape[0]


def ao_integrals(geometry) -> AOIntegrals:
    """One- and two-electron integrals over contracted 1s STO-3G functions on H atoms."""
    coords = np.array([xyz for _, xyz in geometry], dtype=np.float64) / BOHR
    nat = len(coords)
    nprim = len(_STO3G_EXP)
    # primitive table: (natom*nprim) entries
    ctr = np.repeat(coords, nprim, axis=0)                       # (P,3)
    alpha = np.tile(_STO3G_EXP, nat)                             # (P,)
    coef = np.tile(_STO3G_COEF, nat) * (2.0 * alpha / np.pi) ** 0.75
    P = len(alpha)
    owner = np.repeat(np.arange(nat), nprim)

    a = alpha[:, None]
    b = alpha[None, :]
    p = a + b
    mu = a * b / p
    AB2 = ((ctr[:, None, :] - ctr[None, :, :]) ** 2).sum(-1)
    Kab = np.exp(-mu * AB2)
    Pc = (a[..., None] * ctr[:, None, :] + b[..., None] * ctr[None, :, :]) / p[..., None]

    Sp = (np.pi / p) ** 1.5 * Kab
    Tp = mu * (3.0 - 2.0 * mu * AB2) * Sp
    Vp = np.zeros_like(Sp)
    for C in coords:  # Z=1
        PC2 = ((Pc - C) ** 2).sum(-1)
        Vp += -(2.0 * np.pi / p) * Kab * _boys0(p * PC2)

    cc = coef[:, None] * coef[None, :]
    contract = np.zeros((P, nat))
    contract[np.arange(P), owner] = 1.0

    def c2(M):
        return contract.T @ (cc * M) @ contract

    S = c2(Sp)
    T = c2(Tp)
    V = c2(Vp)

    # ERI over primitive pairs
    pp = p.reshape(-1)
    Kp = (Kab * cc).reshape(-1)
    Pp = Pc.reshape(-1, 3)
    PQ2 = ((Pp[:, None, :] - Pp[None, :, :]) ** 2).sum(-1)
    pq = pp[:, None] * pp[None, :]
    ps = pp[:, None] + pp[None, :]
    eri_prim = 2.0 * np.pi ** 2.5 / (pq * np.sqrt(ps)) * _boys0(pq / ps * PQ2)
    eri_prim *= Kp[:, None] * Kp[None, :]
    # contract pair index (i_prim, j_prim) -> (i_ao, j_ao)
    pair_c = np.einsum("pa,qb->pqab", contract, contract).reshape(P * P, nat * nat)
    eri = (pair_c.T @ eri_prim @ pair_c).reshape(nat, nat, nat, nat)

    # renormalise contracted functions to <phi|phi> = 1
    nrm = 1.0 / np.sqrt(np.diag(S))
    S = S * nrm[:, None] * nrm[None, :]
    T = T * nrm[:, None] * nrm[None, :]
    V = V * nrm[:, None] * nrm[None, :]
    eri = eri * nrm[:, None, None, None] * nrm[None, :, None, None] \
        * nrm[None, None, :, None] * nrm[None, None, None, :]

    e_nuc = 0.0
    for i in range(nat):
        for j in range(i):
            e_nuc += 1.0 / math.sqrt(((coords[i] - coords[j]) ** 2).sum())
    return AOIntegrals(coords, S, T, V, eri, e_nuc)


def _fock(h, eri, dm):
    J = np.einsum("pqrs,rs->pq", eri, dm)
    K = np.einsum("prqs,rs->pq", eri, dm)
    return h + J - 0.5 * K


def rhf(h, eri, nelec, S=None, dm0=None, conv=1e-13, max_cycle=200, diis_space=8):
    """Closed-shell RHF with Pulay DIIS.  Returns (e_elec, mo_energy, mo_coeff, dm, converged).

    Used (a) on the AO integrals as the stand-in for pyscf RHF
    (`mqc/system/fragment.py:49-63`) and (b) inside the embedding space, the way
    `mqc/solver/dmet_solver_fci.py:28-42` does, to get an occupied/virtual
    splitting for the UCCSD ansatz.
    """
    n = h.shape[0]
    nocc = nelec // 2
    if S is None:
        S = np.eye(n)
    s_e, s_v = np.linalg.eigh(S)
    X = s_v @ np.diag(s_e ** -0.5) @ s_v.T

    def diag(F):
        e, c = np.linalg.eigh(X.T @ F @ X)
        return e, X @ c

    if dm0 is None:
        e, c = diag(h)
        dm = 2.0 * c[:, :nocc] @ c[:, :nocc].T
    else:
        dm = dm0
    fs, es = [], []
    e_old = 0.0
    converged = False
    for it in range(max_cycle):
        F = _fock(h, eri, dm)
        e_el = 0.5 * np.einsum("pq,pq->", dm, h + F)
        err = X.T @ (F @ dm @ S - S @ dm @ F) @ X
        fs.append(F)
        es.append(err)
        if len(fs) > diis_space:
            fs.pop(0)
            es.pop(0)
        if np.abs(err).max() < 1e-10 and abs(e_el - e_old) < conv:
            converged = True
            break
        e_old = e_el
        if len(fs) >= 2:
            m = len(fs)
            B = -np.ones((m + 1, m + 1))
            B[m, m] = 0.0
            for i in range(m):
                for j in range(m):
                    B[i, j] = np.einsum("pq,pq->", es[i], es[j])
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            try:
                w = np.linalg.solve(B, rhs)[:m]
                F = sum(wi * Fi for wi, Fi in zip(w, fs))
            except np.linalg.LinAlgError:
                pass
        e, c = diag(F)
        dm = 2.0 * c[:, :nocc] @ c[:, :nocc].T
    F = _fock(h, eri, dm)
    e, c = diag(F)
    e_el = 0.5 * np.einsum("pq,pq->", dm, h + F)
    return e_el, e, c, dm, converged


@dataclass
class LocalIntegrals:
    """What QC-DMET's ``localintegrals`` holds for the DMET loop
    (`mqc/dcalgo/dmet.py:27-30`): everything in the orthonormal localized basis.
    For H/STO-3G ``meta_lowdin`` reduces to symmetric S^(-1/2)
    (`test/tmp.molden:78-219`)."""
    h_loc: np.ndarray
    eri_loc: np.ndarray
    fock_loc: np.ndarray     # RHF Fock in the localized basis
    e_nuc: float
    e_rhf: float
    nelec: int
    ao2loc: np.ndarray

    @property
    def norb(self):
        return self.h_loc.shape[0]


def local_integrals(geometry, nelec=None) -> LocalIntegrals:
    ao = ao_integrals(geometry)
    n = ao.nao
    if nelec is None:
        nelec = n
    e_el, mo_e, mo_c, dm, ok = rhf(ao.hcore, ao.eri, nelec, ao.S)
    if not ok:
        raise RuntimeError("synthetic RHF did not converge")
    s_e, s_v = np.linalg.eigh(ao.S)
    X = s_v @ np.diag(s_e ** -0.5) @ s_v.T            # Lowdin
    h_loc = X.T @ ao.hcore @ X
    eri_loc = np.einsum("pqrs,pa,qb,rc,sd->abcd", ao.eri, X, X, X, X, optimize=True)
    fock_ao = _fock(ao.hcore, ao.eri, dm)
    fock_loc = X.T @ fock_ao @ X
    return LocalIntegrals(h_loc, eri_loc, fock_loc, ao.e_nuc, e_el + ao.e_nuc, nelec, X)


def mo_integrals(geometry, nelec=None):
    """Full-molecule RHF-MO integrals (config C3: H12 chain, 12 orbitals, no embedding).
    Returns (h_mo, eri_mo, e_nuc, e_rhf)."""
    ao = ao_integrals(geometry)
    if nelec is None:
        nelec = ao.nao
    e_el, mo_e, C, dm, ok = rhf(ao.hcore, ao.eri, nelec, ao.S)
    if not ok:
        raise RuntimeError("synthetic RHF did not converge")
    h_mo = C.T @ ao.hcore @ C
    eri_mo = np.einsum("pqrs,pa,qb,rc,sd->abcd", ao.eri, C, C, C, C, optimize=True)
    return h_mo, eri_mo, ao.e_nuc, e_el + ao.e_nuc


def random_fragment_integrals(n_orb: int, seed: int = 1234):
    """Seeded generic fragment Hamiltonian:
    h1 = A + A^T with A ~ N(0,1)/n; ERI = L L^T over symmetric pair index with
    8-fold symmetry, scaled 1/n^2."""
    rng = np.random.default_rng(seed)
    n = n_orb
    A = rng.standard_normal((n, n)) / n
    h1 = A + A.T
    npair = n * (n + 1) // 2
    Lf = rng.standard_normal((npair, npair)) / n
    M = Lf @ Lf.T / npair
    iu = np.triu_indices(n)
    pair = np.zeros((n, n), dtype=np.int64)
    pair[iu] = np.arange(npair)
    pair = np.maximum(pair, pair.T)
    eri = M[pair[:, :, None, None], pair[None, None, :, :]]
    return h1, np.ascontiguousarray(eri)
