"""GPU: `pbc_helper.get_eri_7d` against the oracle, and on the LO-basis GDF tensor of `transform_gdf_to_lo` against
the AO integrals rotated with C_ao_lo -- the comparison of libdmet/basis_transform/test/test_transform_gdf.py:108-121
(threshold 1e-10)."""
import numpy as np
import pytest

from helpers import problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kmesh,nao,naux", [([1, 1, 3], 4, 30), ([2, 2, 1], 7, 19), ([3, 1, 1], 26, 40)])
def test_eri_7d_vs_oracle(dev, kmesh, nao, naux):
    from libdmet_preview_b200 import pbc_helper
    from oracle import pbc_helper as o_pbc
    gdf, _, _ = problem(kmesh, nao, naux, 2)
    got = pbc_helper.get_eri_7d(gdf.cell, gdf)
    ref = o_pbc.get_eri_7d(gdf.cell, gdf)
    assert got.shape == ref.shape and got.dtype == np.complex128
    assert np.abs(got - ref).max() < 1e-10
    with pytest.raises(NotImplementedError):
        pbc_helper.get_eri_7d(gdf.cell, gdf, compact=True)


def test_eri_7d_of_lo_gdf(dev):
    from libdmet_preview_b200 import pbc_helper, eri_transform as et
    from oracle import pbc_helper as o_pbc
    gdf, C, _ = problem([1, 2, 2], 6, 17, 2)
    nk = len(gdf.kpts_scaled)
    lo = et.transform_gdf_to_lo(gdf, C, fname=None)
    got = pbc_helper.get_eri_7d(lo.cell, lo)
    ao = o_pbc.get_eri_7d(gdf.cell, gdf)
    kc = o_pbc.get_kconserv(gdf.kpts_scaled)
    for i in range(nk):
        for j in range(nk):
            for k in range(nk):
                l = kc[i, j, k]
                ref = np.einsum("pqrs,pi,qj,rk,sl->ijkl", ao[i, j, k], C[i].conj(), C[j], C[k].conj(), C[l],
                                optimize=True)
                assert np.abs(got[i, j, k] - ref).max() < 1e-10
