"""CPU: the oracle's `get_eri_7d` convention (oracle/pbc_helper.py <- libdmet/routine/pbc_helper.py:276-294 + PySCF's
`GDF.get_eri`) pinned through the reference's own consumer of those integrals, `get_jk_from_eri_7d`
(pbc_helper.py:314-351): J - K/2 built from the 7-d integrals equals the effective potential of the self-consistent
lattice mean field obtained from the supercell integrals of `get_emb_eri` (tests/hf_in_dmet.py)."""
import numpy as np
import pytest

from libdmet_preview_b200 import synthetic
import hf_in_dmet as hd
from oracle import pbc_helper as o_pbc


@pytest.mark.parametrize("kmesh,nao,naux,nocc", [([1, 1, 3], 4, 12, 2), ([2, 1, 2], 3, 10, 1)])
def test_jk_from_eri_7d_equals_supercell_jk(kmesh, nao, naux, nocc):
    gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=71)
    hcore = hd.gapped_hcore(kmesh, nao, nocc, seed=5)
    mf = hd.lattice_scf(gdf, hcore, nocc)
    eri = o_pbc.get_eri_7d(gdf.cell, gdf)
    nk = len(gdf.kpts_scaled)
    assert eri.shape == (nk, nk, nk, nao, nao, nao, nao)
    vj, vk = o_pbc.get_jk_from_eri_7d(eri, mf["rdm1"])
    assert np.abs(vj - 0.5 * vk - mf["vhf"]).max() < 1e-10
    # permutational symmetry of the k-space integrals: (pq|rs)(ki,kj,kk,kl) = (rs|pq)(kk,kl,ki,kj)
    kc = o_pbc.get_kconserv(gdf.kpts_scaled)
    for (i, j, k) in [(0, 1, 2 % nk), (1, 1, 0), (nk - 1, 0, 1)]:
        l = kc[i, j, k]
        assert np.abs(eri[i, j, k] - eri[k, l, i].transpose(2, 3, 0, 1)).max() < 1e-12
