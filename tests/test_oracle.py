"""CPU: the oracle satisfies the invariants the reference's own tests assert (thresholds are the reference's):
libdmet/basis_transform/test/test_eri_transform_gdf.py:54 (C_ao_eo re-entry < 1e-14), :28-29 (chunking),
:94 (time reversal vs plain < 1e-10); test_eri_transform_uhf.py:36-40 (aa, ab, bb order);
libdmet/system/test/test_fourier.py:213-218 (k2R/R2k round trip < 1e-11); test_slater.py:235-266 (unit2emb)."""
import numpy as np
import pytest

from helpers import problem
from oracle import eri_transform as o_eri, fourier as o_f, pyscf_lib as olib, slater as o_sl, make_basis as o_mb
from oracle.fourier import max_abs


@pytest.mark.parametrize("kmesh", [[1, 1, 3], [1, 2, 2], [3, 1, 2]])
def test_trs_equals_plain(kmesh):
    gdf, C, basis = problem(kmesh, 5, 12, 6)
    e_trs = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, t_reversal_symm=True)
    e_pln = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, t_reversal_symm=False)
    assert e_trs.shape == (1, 21, 21)
    assert max_abs(e_trs - e_pln) < 1e-10
    assert max_abs(e_trs[0] - e_trs[0].T) < 1e-12       # (ij|kl) = (kl|ij)


def test_c_ao_eo_reentry_and_chunking():
    gdf, C, basis = problem([1, 1, 3], 6, 40, 7)
    e0 = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    C_ao_eo = o_eri.build_C_ao_emb(gdf, C, basis) * (gdf.nkpts ** 0.75)
    e1 = o_eri.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_eo=C_ao_eo)
    assert max_abs(e0 - e1) < 1e-14
    e2 = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, max_memory=0.02)   # blksize 16 -> 3 chunks
    assert max_abs(e0 - e2) < 1e-12
    with pytest.raises(ValueError):
        o_eri.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_lo=C, C_ao_eo=C_ao_eo)


def test_unrestricted_order_and_symmetries():
    gdf, C, basis = problem([1, 1, 3], 4, 10, 5, spin=2)
    e = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    assert e.shape == (3, 15, 15)
    # aa and bb equal the restricted result with that spin's orbitals; ab is the mixed product
    for s, blk in ((0, 0), (1, 2)):
        er = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C[s], basis=basis[s:s + 1])
        assert max_abs(er[0] - e[blk]) < 1e-12
    assert max_abs(e[1] - e[1].T) > 1e-6                 # ab is not symmetric under pair exchange
    e1 = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, symmetry=1)
    assert e1.shape == (3, 5, 5, 5, 5)
    assert max_abs(e1[1] - e1[1].transpose(1, 0, 2, 3)) < 1e-12
    with pytest.raises(ValueError):
        o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, symmetry=8)


def test_restore_roundtrip_and_unit2emb():
    gdf, C, basis = problem([1, 1, 2], 4, 9, 4)
    e4 = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, symmetry=4)
    e1 = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, symmetry=1)
    e8 = o_eri.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, symmetry=8)
    assert e1.shape == (1, 4, 4, 4, 4) and e8.shape == (1, 55)
    assert max_abs(olib.restore(4, e1[0], 4) - e4[0]) < 1e-14
    assert max_abs(olib.restore(4, e8[0], 4) - e4[0]) < 1e-14
    assert max_abs(e1[0] - e1[0].transpose(2, 3, 0, 1)) < 1e-12
    eu = o_eri.get_unit_eri(gdf.cell, gdf, C_ao_lo=C, symmetry=1)
    big = o_sl.unit2emb(eu, 6)
    assert big.shape == (1, 6, 6, 6, 6) and max_abs(big[:, :4, :4, :4, :4] - eu) == 0.0
    assert max_abs(big[:, 4:]) == 0.0


def test_pyscf_lib_semantics():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 5, 5)) + 1j * rng.standard_normal((3, 5, 5))
    p = olib.pack_tril(a)
    assert p.shape == (3, 15) and p[1, 4 * 5 // 2 + 2] == a[1, 4, 2]
    u = olib.unpack_tril(p)
    assert max_abs(np.tril(u) - np.tril(a)) == 0 and max_abs(u[:, 1, 3] - a[:, 3, 1].conj()) == 0
    s = olib.hermi_sum(a.copy(), axes=(0, 2, 1), hermi=olib.SYMMETRIC)
    assert max_abs(s - (a + a.transpose(0, 2, 1))) == 0
    ci = rng.standard_normal((5, 3)) + 1j * rng.standard_normal((5, 3))
    cj = rng.standard_normal((5, 4)) + 1j * rng.standard_normal((5, 4))
    mo, sl = olib.conc_mos(ci, cj)
    out = olib.r_e2(a.reshape(3, 25), mo, sl).reshape(3, 3, 4)
    ref = np.einsum("Lpq,pi,qj->Lij", a, ci.conj(), cj)
    assert max_abs(out - ref) < 1e-13
    # dot_eri_dm conventions (libdmet/solver/scf.py:269-271) on an s1 tensor with 8-fold symmetry
    n = 4
    x = rng.standard_normal((n * (n + 1) // 2,) * 2)
    e4 = x + x.T
    e1 = olib.restore(1, e4, n)
    d = rng.standard_normal((n, n))
    d = d + d.T
    vj, vk = olib.dot_eri_dm(e4, d)
    assert max_abs(vj - np.einsum("ijkl,kl->ij", e1, d)) < 1e-13
    assert max_abs(vk - np.einsum("ijkl,il->jk", e1, d)) < 1e-13
    vj8, vk8 = olib.dot_eri_dm(olib.restore(8, e4, n), d)
    assert max_abs(vj8 - vj) < 1e-13 and max_abs(vk8 - vk) < 1e-13


def test_fourier_conventions():
    kmesh = [2, 1, 3]
    nk = 6
    rng = np.random.default_rng(1)
    A = rng.standard_normal((nk, 3, 4))
    ks = o_f.make_kpts_scaled(kmesh)
    ph = o_f.get_phase_R2k_scaled(kmesh, ks)
    Ak = o_f.R2k(A, kmesh)
    assert max_abs(Ak - np.einsum("Rim,Rk->kim", A, ph)) < 1e-13          # eri_transform.py:125 == FFTtoK
    back = o_f.k2R(Ak, kmesh)
    assert back.dtype == np.float64 and max_abs(back - A) < 1e-11        # test_fourier.py:213-218
    A4 = rng.standard_normal((2, nk, 3, 3))
    assert max_abs(o_f.k2R(o_f.R2k(A4, kmesh), kmesh) - A4) < 1e-11
    with pytest.raises(ValueError):
        o_f.R2k(A[0], kmesh)
    # kpt_member / round_to_FBZ conventions (libdmet/system/test/test_fourier.py:9-41)
    assert list(o_f.kpt_member(np.array([0.5, 0.0, 1.0 / 3 + 1.0]), ks)) == [list(map(tuple, np.round(ks, 8))).index(
        (-0.5, 0.0, round(1.0 / 3, 8)))]
    r = o_f.round_to_FBZ(np.array([[0.5, 0.75, -0.5]]))
    assert np.allclose(r, [[-0.5, -0.25, -0.5]])


def test_h1_to_lo_identities():
    gdf, C, basis = problem([1, 1, 3], 5, 4, 4, spin=2)
    rng = np.random.default_rng(2)
    h = rng.standard_normal((3, 5, 5)) + 1j * rng.standard_normal((3, 5, 5))
    out = o_mb.transform_h1_to_lo(h, C)
    assert out.shape == (2, 3, 5, 5)
    for s in range(2):
        for k in range(3):
            assert max_abs(out[s, k] - C[s, k].conj().T @ h[k] @ C[s, k]) < 1e-13
    assert o_mb.transform_h1_to_lo(h, C[0]).shape == (3, 5, 5)
    assert max_abs(o_mb.transform_h1_to_lo(0.0, C[0])) == 0.0
    assert o_mb.transform_h1_to_lo(np.array([0.0, 0.0]), C).shape == (2, 3, 5, 5)
    m = o_mb.multiply_basis(C, C[0])
    assert m.shape == (2, 3, 5, 5) and max_abs(m[1, 2] - C[1, 2] @ C[0, 2]) < 1e-13
