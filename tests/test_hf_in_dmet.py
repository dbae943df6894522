"""CPU: the HF-in-DMET identities of libdmet/test/test_mfd.py on a seeded synthetic cell, oracle routines only.
Pins the restatement of the lattice mean field (oracle/mfd.py <- libdmet/routine/mfd.py:33-108, 235-427, 862-957)
and the self-consistent synthetic mean field the GPU energy tests start from (tests/hf_in_dmet.py):
    lattice HF reproduces the SCF density matrix and energy          (test_mfd.py:107-113: 1e-8 / 1e-10)
    the folded density matrix is a fixed point of the impurity HF    (test_mfd.py:138: 1e-8)
    the fragment energy equals the k-point HF energy per cell        (test_mfd.py:153: 1e-8 Ha)"""
import numpy as np
import pytest

from helpers import OracleLattice
from libdmet_preview_b200 import synthetic
import hf_in_dmet as hd
from oracle import mfd as o_mfd


def build(kmesh, nao, naux, nocc, seed=0, sym=4):
    gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=40 + seed)
    spin = 2 if np.ndim(nocc) else 1
    hcore = hd.gapped_hcore(kmesh, nao, nocc, seed=seed)
    mf = hd.lattice_scf(gdf, hcore, nocc)
    C = synthetic.make_C_ao_lo(kmesh, nao, seed=50 + seed, spin=(2 if spin == 2 else None))
    nk = len(gdf.kpts_scaled)
    ovlp = np.asarray([np.eye(nao, dtype=np.complex128)] * nk)
    O = OracleLattice(gdf, C, hcore, ovlp, mf["rdm1"], mf["vhf"], eri_symmetry=sym, H0=0.125)
    return gdf, mf, O


@pytest.mark.parametrize("kmesh,nao,naux,nocc,sym", [([1, 1, 3], 4, 12, 2, 4), ([2, 1, 2], 5, 14, 2, 1),
                                                      ([1, 1, 3], 5, 12, (3, 2), 4)])
def test_identities_oracle(kmesh, nao, naux, nocc, sym):
    gdf, mf, O = build(kmesh, nao, naux, nocc, sym=sym)
    restricted = mf["spin"] == 1
    # impurity = the whole cell, every orbital contributes to the bath: the embedding space then holds exactly nao
    # electrons per spin channel (the reference test has IAO valence orbitals spanning the occupied bands instead,
    # test_mfd.py:76-80, 127; the seeded local orbitals here are generic rotations of the AOs)
    O.val_idx, O.virt_idx = list(range(nao)), []
    filling = nocc / float(nao) if restricted else [nocc[0] / float(nao), nocc[1] / float(nao)]
    res = hd.dmet_cycle(O, hd.OracleMods, filling, restricted, nelec_emb=nao if restricted else [nao, nao])
    assert res["basis"].shape[-1] == 2 * nao
    assert abs(res["E_lattice_HF"] - (mf["e_cell"] + 0.125)) < 1e-10
    assert res["rdm_diff"] < 1e-8
    assert res["fixed_point_diff"] < 1e-8
    assert abs(res["E_frag"] - (mf["e_cell"] + 0.125)) < 1e-8, (res["E_frag"], mf["e_cell"] + 0.125)


def test_assignocc_degenerate_homo():
    """zero-temperature fractional filling of a degenerate HOMO (mfd.py:936-944)"""
    ew = np.asarray([[[-1.0, 0.0, 0.0, 2.0]]])
    occ, mu, _ = o_mfd.assignocc(ew, 2, np.inf, mu0=0.0)
    assert np.allclose(occ, [[[1.0, 0.5, 0.5, 0.0]]]) and mu == 0.0
    occ, mu, _ = o_mfd.assignocc(ew, 3, np.inf, mu0=-5.0)        # mu0 outside: mid-gap of the sorted levels
    assert np.allclose(occ, [[[1.0, 1.0, 1.0, 0.0]]]) and mu == 1.0
    occ, mu, _ = o_mfd.assignocc(np.asarray([ew[0], ew[0] + 0.5]), [2, 1], np.inf, mu0=[0.0, 0.0])
    assert np.allclose(occ[1], [[1.0, 0.0, 0.0, 0.0]])
