"""GPU: get_emb_basis / embHam / transform_trans_inv_k / get_veff against the oracle.
Bath orbitals are compared as subspaces (the reference's own check, libdmet/routine/test/test_slater.py:46-54,
SVD gauge freedom); integrals are compared after feeding the SAME basis to both implementations."""
import numpy as np
import pytest

from helpers import problem, mean_field, OracleLattice

pytestmark = pytest.mark.gpu
TOL = 1e-10


def make_lattices(kmesh, nao, naux, nval, spin, eri_symmetry):
    from libdmet_preview_b200 import lattice as lat
    gdf, C, _ = problem(kmesh, nao, naux, 2, spin=spin)
    if spin == 1:
        C = C if C.ndim == 3 else C[0]
    hcore, ovlp, vhf, rdm1 = mean_field(kmesh, nao, nval // 2 + 1, spin=spin)
    L = lat.Lattice(gdf.cell, kmesh)
    L.set_val_virt_core(nval, nao - nval, 0)
    L.set_Ham(None, gdf, C, eri_symmetry=eri_symmetry, ovlp=ovlp, hcore=hcore, rdm1=rdm1, vhf=vhf, H0=1.5)
    O = OracleLattice(gdf, C, hcore, ovlp, rdm1, vhf, eri_symmetry=eri_symmetry, H0=1.5)
    O.val_idx, O.virt_idx = list(range(nval)), list(range(nval, nao))
    return L, O


@pytest.mark.parametrize("spin", [1, 2])
def test_set_ham_matches_oracle(dev, spin):
    L, O = make_lattices([1, 2, 3], 6, 12, 4, spin, 4)
    for name in ("hcore_lo_k", "ovlp_lo_k", "vhf_lo_k", "fock_lo_k", "rdm1_lo_k"):
        a, b = getattr(L, name), getattr(O, name)
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-12, name
    assert L.rdm1_lo_R.dtype == np.float64 and np.abs(L.rdm1_lo_R - O.rdm1_lo_R).max() < 1e-12
    A = np.random.default_rng(0).standard_normal((6, 6, 6))
    assert np.array_equal(L.expand(A), O.expand(A)) and np.array_equal(L.extract_stripe(L.expand(A)), A)


@pytest.mark.parametrize("spin", [1, 2])
def test_get_emb_basis(dev, spin):
    from libdmet_preview_b200 import slater
    from oracle import slater as osl
    L, O = make_lattices([1, 2, 3], 6, 12, 4, spin, 4)
    rho = L.rdm1_lo_R * (0.5 if spin == 1 else 1.0)
    b1 = slater.get_emb_basis(L, rho)
    b2 = osl.get_emb_basis(O, rho)
    assert b1.shape == b2.shape and b1.shape[0] == spin and b1.shape[1:3] == (6, 6)
    nimp = 6
    for s in range(spin):
        f1, f2 = b1[s].reshape(36, -1), b2[s].reshape(36, -1)
        assert np.array_equal(f1[:, :nimp], f2[:, :nimp])                      # impurity columns: identity
        assert osl.check_span_same_space(f1[:, nimp:], f2[:, nimp:])            # bath: same span
        assert np.abs(f1.T @ f1 - np.eye(f1.shape[1])).max() < 1e-12
    b3 = slater.get_emb_basis(L, rho, valence_bath=False, nbath=2)
    assert b3.shape[-1] == nimp + 2
    with pytest.raises(ValueError):
        slater.get_emb_basis(L, rho, kind="nope")


@pytest.mark.parametrize("spin,sym", [(1, 4), (1, 1), (1, 8), (2, 4), (2, 1)])
def test_embham(dev, spin, sym):
    from libdmet_preview_b200 import slater
    from oracle import slater as osl
    L, O = make_lattices([1, 1, 3], 5, 11, 3, spin, sym)
    rho = L.rdm1_lo_R * (0.5 if spin == 1 else 1.0)
    basis = slater.get_emb_basis(L, rho)
    Ham, none = slater.embHam(L, basis, None)
    Ref, _ = osl.embHam(O, basis, None)
    assert none is None and Ham.norb == Ref.norb == basis.shape[-1]
    assert Ham.restricted == (spin == 1) and Ham.bogoliubov is False and Ham.H0 == 1.5
    assert Ham.H2["ccdd"].shape == Ref.H2["ccdd"].shape
    assert np.abs(Ham.H2["ccdd"] - Ref.H2["ccdd"]).max() < TOL                   # aa, bb, ab order for spin 2
    assert Ham.H1["cd"].shape == Ref.H1["cd"].shape == (spin, Ham.norb, Ham.norb)
    assert np.abs(Ham.H1["cd"] - Ref.H1["cd"]).max() < TOL
    assert Ham.ovlp.shape == Ref.ovlp.shape and np.abs(Ham.ovlp - Ref.ovlp).max() < TOL
    assert np.abs(L.JK_core - O.JK_core).max() < TOL                               # side effect, slater.py:640-643
    # H2_given re-entry
    Ham2, _ = slater.embHam(L, basis, None, H2_given=Ref.H2["ccdd"])
    assert np.abs(Ham2.H1["cd"] - Ref.H1["cd"]).max() < TOL
    # H2_fname re-entry: integrals stored by an earlier run in the dataset 'emb_eri' (slater.py:349-355)
    import tempfile, os
    from libdmet_preview_b200 import h5lite
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "emb_eri.h5")
        with h5lite.Writer(fn) as w:
            w["emb_eri"] = Ham.H2["ccdd"]
        Ham3, _ = slater.embHam(L, basis, None, H2_fname=fn)
    assert np.array_equal(Ham3.H2["ccdd"], Ham.H2["ccdd"])
    assert np.abs(Ham3.H1["cd"] - Ham.H1["cd"]).max() < 1e-12
    # branches that are not mirrored are refused before any ERI work (UnsupportedBranch <: NotImplementedError)
    from libdmet_preview_b200._lib import UnsupportedBranch
    n0 = dev.launch_count()
    for kw in (dict(int_bath=False), dict(dft=True), dict(qsgw=True)):
        with pytest.raises(UnsupportedBranch):
            slater.embHam(L, basis, None, **kw)
    assert dev.launch_count() == n0
    # non-interacting bath ERI: unit ERI zero-padded (slater.py:464-472)
    H2u, _ = slater._embHam2e(L, basis, None, True, int_bath=False)
    Ru = osl._embHam2e(O, basis, None, True, int_bath=False)
    assert H2u.shape == Ru.shape and np.abs(H2u - Ru).max() < TOL


def test_transform_helpers(dev):
    from libdmet_preview_b200 import slater
    from oracle import slater as osl
    rng = np.random.default_rng(5)
    bk = rng.standard_normal((2, 6, 7, 9)) + 1j * rng.standard_normal((2, 6, 7, 9))
    H = rng.standard_normal((6, 7, 7)) + 1j * rng.standard_normal((6, 7, 7))
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = slater.transform_h1(H, bk)
        one = slater.transform_trans_inv_k(bk[1], H)
    assert np.abs(got - osl.transform_h1(H, bk)).max() < 1e-11
    assert np.abs(one - osl.transform_trans_inv_k(bk[1], H)).max() < 1e-11
    assert any("imag part" in str(x.message) for x in w)                          # slater_helper.py:47-48
    e = rng.standard_normal((1, 45, 45))
    e = e + e.transpose(0, 2, 1)
    d = rng.standard_normal((1, 9, 9))
    d = d + d.transpose(0, 2, 1)
    assert np.abs(slater.get_veff(d, e) - osl.get_veff(d, e)).max() < 1e-11
    u = rng.standard_normal((3, 6, 6))
    assert np.array_equal(slater.unit2emb(u, 5), osl.unit2emb(u, 5))


def test_h2_scaling_kernels(dev):
    """get_H1_scaled / get_H2_scaled (slater.py:1716-1778) against the oracle for s4 and s1 layouts"""
    from libdmet_preview_b200 import slater
    from oracle import slater as osl, pyscf_lib as olib
    rng = np.random.default_rng(8)
    n, imp = 7, [0, 1, 2, 5]
    x = rng.standard_normal((3, 28, 28))
    a, b = x.copy(), x.copy()
    assert np.abs(slater.get_H2_scaled(a, imp) - osl.get_H2_scaled(b, imp)).max() < 1e-15
    y = rng.standard_normal((1, n, n, n, n))
    a, b = y.copy(), y.copy()
    assert np.abs(slater.get_H2_scaled(a, imp) - osl.get_H2_scaled(b, imp)).max() < 1e-15
    h = rng.standard_normal((2, n, n))
    assert np.array_equal(slater.get_H1_scaled(h.copy(), imp), osl.get_H1_scaled(h.copy(), imp))
    with pytest.raises(ValueError):
        slater.get_H2_scaled(rng.standard_normal((4, 4)), imp)


def test_gso_embham_and_helpers(dev):
    """spinless helpers and the GSO embedding Hamiltonian at a less trivial size against the oracle"""
    from helpers import GSOLattice, gso_basis
    from libdmet_preview_b200 import spinless, fourier
    from oracle import spinless as osp, fourier as of
    kmesh, nao, naux, nemb = [2, 1, 3], 5, 13, 9
    gdf, C, _ = problem(kmesh, nao, naux, 2, spin=2)
    L1 = GSOLattice(gdf, C, fourier)
    L2 = GSOLattice(gdf, C, of)
    basis = gso_basis(kmesh, nao, nemb)
    bk = of.R2k(basis, kmesh)
    ka, kb = osp.separate_basis(bk)
    for H in (L1.hcore_lo_k, L1.hcore_lo_k[:2]):
        assert np.abs(spinless.transform_trans_inv_k(ka, kb, H) - osp.transform_trans_inv_k(ka, kb, H)).max() < 1e-11
    Ra, Rb = osp.separate_basis(basis)
    v = np.random.default_rng(1).standard_normal((3, nao, nao))
    assert np.abs(spinless.transform_local(Ra, Rb, v) - osp.transform_local(Ra, Rb, v)).max() < 1e-13
    assert np.abs(spinless.transform_imp(Ra, Rb, v) - osp.transform_imp(Ra, Rb, v)).max() < 1e-13
    Ham, _ = spinless.embHam(L1, basis, None, -0.2, hcore_add=v)
    Ref, _ = osp.embHam(L2, basis, None, -0.2, hcore_add=v)
    assert np.abs(Ham.H2["ccdd"] - Ref.H2["ccdd"]).max() < TOL and np.abs(Ham.H1["cd"] - Ref.H1["cd"]).max() < TOL
    assert np.abs(L1.JK_core - L2.JK_core).max() < TOL
    with pytest.raises(NotImplementedError):
        spinless.embHam(L1, basis, None, 0.0, int_bath=False)
