"""CPU: host-side logic of the LO-basis GDF tensor (`transform_gdf_to_lo`, eri_transform.py:1312-1427): the
time-reversal map over the stored pairs, the storage rules and the way `LoGDF.load` serves blocks (PySCF `_load3c` +
`sr_loop(compact=False)` semantics: packed blocks unpacked Hermitian, missing (ki, kj) = conj-transpose of (kj, ki))."""
import numpy as np
import pytest

from libdmet_preview_b200 import synthetic
from libdmet_preview_b200 import eri_transform as et
from oracle import eri_transform as oe


@pytest.mark.parametrize("kmesh", [[1, 1, 3], [2, 2, 1], [2, 1, 3]])
def test_time_reversal_mask_matches_oracle(kmesh):
    g = synthetic.SyntheticGDF(kmesh, 3, 4, seed=1)
    pairs = et.stored_pairs(g)
    assert pairs == oe.stored_pairs(g) and pairs[0] == (0, 0) and all(j <= i for i, j in pairs)
    ks = g.kpts_scaled
    kptij = [np.concatenate([ks[i], ks[j]]) for i, j in pairs]
    m_ours = et.get_mask_kptij_lst(None, kptij, scaled=True)
    m_ref = oe.get_mask_kptij_lst(kptij)
    assert np.array_equal(m_ours, m_ref)
    # through the cell (absolute k-points), as the reference calls it
    kabs = np.asarray([(g.kpts[i], g.kpts[j]) for i, j in pairs])
    assert np.array_equal(et.get_mask_kptij_lst(g.cell, kabs), m_ref)
    for a, b in enumerate(m_ref):
        if b >= 0:      # pair b is minus pair a and is marked as filled
            assert m_ref[b] == -2
            s = kptij[a] + kptij[b]
            assert np.abs(s - np.round(s)).max() < 1e-9


@pytest.mark.parametrize("nlo", [5, 3])
def test_lo_gdf_serves_blocks_like_load3c(nlo):
    kmesh, nao, naux = [2, 1, 3], 5, 6
    g = synthetic.SyntheticGDF(kmesh, nao, naux, seed=9)
    C = synthetic.make_C_ao_lo(kmesh, nao, nlo, seed=4)
    stored = oe.transform_gdf_to_lo(g, C)
    lo = et.LoGDF(g, nlo, et.stored_pairs(g), stored)
    assert lo.nao == nlo and lo.naux == naux and lo.cell.nao_nr() == nlo and g.cell.nao_nr() == nao
    nk = len(g.kpts_scaled)
    for ki in range(nk):
        for kj in range(nk):
            want = np.einsum("pm,Lpq,qn->Lmn", C[ki].conj(), g.load(ki, kj), C[kj])
            got = lo.load(ki, kj)
            assert got.shape == (naux, nlo, nlo) and got.dtype == np.complex128
            assert np.abs(got - want).max() < 1e-13, (ki, kj)
    # storage rules: real packed at (Gamma, Gamma), complex packed on the diagonal, full otherwise
    npk = nlo * (nlo + 1) // 2
    for pos, (i, j) in enumerate(lo.kptij_idx):
        x = lo.j3c[pos]
        if i == j == 0:
            assert x.dtype == np.float64 and x.shape == (naux, npk)
        elif i == j:
            assert x.dtype == np.complex128 and x.shape == (naux, npk)
        else:
            assert x.dtype == np.complex128 and x.shape == (naux, nlo * nlo)


def test_lo_gdf_npz_round_trip(tmp_path):
    g = synthetic.SyntheticGDF([1, 1, 3], 4, 5, seed=2)
    C = synthetic.make_C_ao_lo([1, 1, 3], 4, seed=3)
    lo = et.LoGDF(g, 4, et.stored_pairs(g), oe.transform_gdf_to_lo(g, C))
    f = str(tmp_path / "lo.npz")
    lo.save(f)
    z = np.load(f)
    assert z["j3c-kptij"].shape == (6, 2, 3)
    assert all(np.array_equal(z["j3c/%d/0" % k], v) for k, v in lo.j3c.items())
    # HDF5 in the reference's layout (written by h5lite, no h5py needed), served back through the file provider
    from libdmet_preview_b200.gdf_file import GDFFile
    h5 = str(tmp_path / "lo.h5")
    lo.save(h5)
    back = GDFFile(h5, cell=lo.cell, kpts=g.kpts)
    assert back.version == "v1" and back.nao == 4 and back.naux == 5 and back.kptij_idx == lo.kptij_idx
    for ki in range(3):
        for kj in range(3):
            assert np.array_equal(back.load(ki, kj), lo.load(ki, kj))
