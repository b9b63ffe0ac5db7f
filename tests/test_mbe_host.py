"""Host logic of the second consumer (multiscalequantumcomputing_b200/mbe.py): molecule-level VQE with the
reference's active-space rule (`mqc/solver/mbe_solver.py:179-217`) and the measurement-style RDM path
(`mqc/solver/dmet_solver_vqechem.py:119-179`), with the CPU oracle plugged in behind the ansatz surface."""
import numpy as np
import pytest

from oracle_ansatz import OracleAnsatz
from multiscalequantumcomputing_b200 import mbe
from multiscalequantumcomputing_b200.jw import fragment_term_table
from multiscalequantumcomputing_b200.synth.hydrogen import hydrogen_chain, mo_integrals
from oracle.fci import sector_indices
from oracle.hamiltonian import dense_matrix

ORACLE = dict(_ansatz_factory=lambda nq, tab, fac: OracleAnsatz(nq, tab, fac))


def _fci(h, g, n, ne):
    H = dense_matrix(*fragment_term_table(h, g, 0, 0.0, 0.0), 2 * n)
    idx = sector_indices(n, ne)[1]
    return float(np.linalg.eigvalsh(H[np.ix_(idx, idx)].real)[0])


def test_active_space_rule():
    assert mbe.active_space(6, 6) == (0, [0, 1, 2, 3, 4, 5])
    assert mbe.active_space(6, 6, ncas=4) == (1, [1, 2, 3, 4])            # ncore = nocc - floor(ncas / 2)
    assert mbe.active_space(6, 6, ncas=3) == (2, [2, 3, 4])               # 1 occupied + 2 virtual
    with pytest.raises(ValueError):
        mbe.active_space(4, 8, ncas=2)


def test_frozen_core_reproduces_the_hf_energy():
    h, g, e_nuc, e_rhf = mo_integrals(hydrogen_chain(6, 1.0), 6)
    nocc = 3
    e_full = 2 * sum(h[i, i] for i in range(nocc)) + sum(2 * g[i, i, j, j] - g[i, j, j, i] for i in range(nocc) for j in range(nocc))
    nc, mo = mbe.active_space(6, 6, ncas=4)
    ha, ga, ec = mbe.frozen_core(h, g, nc, mo)
    na = nocc - nc                                                         # doubly occupied active orbitals
    e_act = 2 * sum(ha[i, i] for i in range(na)) + sum(2 * ga[i, i, j, j] - ga[i, j, j, i] for i in range(na) for j in range(na))
    assert abs(ec + e_act - e_full) < 1e-12


def test_vqechem_full_space_and_cas():
    h, g, e_nuc, e_rhf = mo_integrals(hydrogen_chain(4, 1.0), 4)
    e = mbe.vqechem(h, g, 4, e_nuc, **ORACLE)
    e_fci = _fci(h, g, 4, 4) + e_nuc
    assert e_fci - 1e-10 < e < e_fci + 5e-4                                # UCCSD: variational, close to FCI
    nc, mo = mbe.active_space(4, 4, ncas=2)
    ha, ga, ec = mbe.frozen_core(h, g, nc, mo)
    e2 = mbe.vqechem(h, g, 4, e_nuc, ncas=2, **ORACLE)
    assert abs(e2 - (_fci(ha, ga, 2, 2) + ec + e_nuc)) < 1e-8              # 2 electrons in 2 orbitals: UCCSD is exact
    assert e < e2 + 1e-12


def test_measurement_style_rdms():
    h, g, e_nuc, e_rhf = mo_integrals(hydrogen_chain(3, 1.1), 2)
    e, ans, th = mbe.vqechem(h, g, 2, e_nuc, return_ansatz=True, **ORACLE)
    r1, r2 = mbe.get_rdm12_openfermion(ans, 3)
    o1, o2 = ans.rdm12(3, 1, 1)
    assert abs(r2 - o2).max() < 1e-12                                     # same convention as make_rdm12 + reorder_rdm
    assert abs(np.trace(r1) - 2.0) < 1e-6 and abs(r1 - o1).max() < 1e-5   # 2 <alpha alpha> vs spin-summed
    assert abs(np.einsum("pq,pq->", h, o1) + 0.5 * np.einsum("pqrs,pqrs->", g, r2) + e_nuc - e) < 1e-10
