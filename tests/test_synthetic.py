"""CPU: the synthetic GDF provider has PySCF's storage symmetries and time-reversal structure."""
import numpy as np

from libdmet_preview_b200 import synthetic


def test_gdf_symmetries():
    g = synthetic.SyntheticGDF([2, 1, 3], 5, 7, seed=3)
    for (i, j) in [(0, 0), (1, 4), (5, 2), (3, 3)]:
        A = g.load(i, j)
        assert A.shape == (7, 5, 5) and A.dtype == np.complex128
        assert np.array_equal(g.load(j, i), A.conj().transpose(0, 2, 1))
        assert np.array_equal(g.load(g.minus[i], g.minus[j]), A.conj())
        assert np.array_equal(g.load(i, j, aux_slice=(2, 5)), A[2:5])
    assert np.abs(g.load(0, 0).imag).max() == 0.0            # Gamma-Gamma block is real
    assert 0.05 < np.abs(g.load(1, 2)).max() < 2.0
    assert not np.array_equal(g.load(1, 2), synthetic.SyntheticGDF([2, 1, 3], 5, 7, seed=4).load(1, 2))


def test_coefficients_and_basis():
    C = synthetic.make_C_ao_lo([1, 2, 3], 6, 4, seed=1, spin=2)
    assert C.shape == (2, 6, 6, 4)
    ks = synthetic.make_kpts_scaled([1, 2, 3])
    minus = [int(synthetic.kpt_member(-k, ks)[0]) for k in ks]
    for s in range(2):
        for k in range(6):
            assert np.allclose(C[s, k].conj().T @ C[s, k], np.eye(4), atol=1e-13)
            assert np.allclose(C[s, minus[k]], C[s, k].conj())
    b = synthetic.make_emb_basis([1, 2, 3], 4, 7, nimp=3)
    flat = b.reshape(1, 24, 7)[0]
    assert np.allclose(flat.T @ flat, np.eye(7), atol=1e-13) and np.array_equal(flat[:3, :3], np.eye(3))
