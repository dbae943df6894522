"""GPU: k2R / R2k / FFTtoK / FFTtoT, transform_h1_to_lo and its siblings, multiply_basis -- public functions with
the reference's signatures against the oracle (scipy FFT / numpy), thresholds from the reference's tests
(libdmet/system/test/test_fourier.py:213-218 1e-11; basis_transform/test/test_make_basis.py:112-139 1e-12)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _z(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.mark.parametrize("kmesh", [[1, 1, 1], [1, 1, 3], [3, 3, 1], [2, 2, 2], [4, 3, 2], [4, 4, 4]])
def test_r2k_k2r_match_fft(dev, kmesh):
    from libdmet_preview_b200 import fourier as f
    from oracle import fourier as of
    rng = np.random.default_rng(sum(kmesh))
    nk = int(np.prod(kmesh))
    A = rng.standard_normal((nk, 5, 7))
    Ak = f.R2k(A, kmesh)
    assert Ak.dtype == np.complex128 and Ak.shape == A.shape
    assert np.abs(Ak - of.R2k(A, kmesh)).max() < 1e-12
    assert np.abs(f.FFTtoK(A, kmesh) - of.FFTtoK(A, kmesh)).max() < 1e-12
    B = _z(rng, 2, nk, 6, 6)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = f.k2R(B, kmesh)
    assert got.dtype == np.float64
    assert np.abs(got - of.k2R(B, kmesh)).max() < 1e-12
    if nk > 1:
        assert any("non-zero imaginary part" in str(x.message) for x in w)     # fourier.py:174-175
    # round trip on physical (real in R) data, < 1e-11
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert np.abs(f.k2R(f.R2k(A, kmesh), kmesh) - A).max() < 1e-11
    Bc = _z(rng, nk, 3, 3)
    assert np.abs(f.R2k(Bc, kmesh) - of.R2k(Bc, kmesh)).max() < 1e-12           # complex input
    with pytest.raises(ValueError):
        f.R2k(A[0], kmesh)
    with pytest.raises(ValueError):
        f.k2R(A[0], kmesh)


def test_phase_conventions(dev):
    from libdmet_preview_b200 import fourier as f, synthetic
    from oracle import fourier as of
    kmesh = [2, 3, 1]
    ks = f.make_kpts_scaled(kmesh)
    assert np.abs(f.get_phase_R2k_scaled(kmesh) - of.get_phase_R2k_scaled(kmesh, ks)).max() < 1e-14
    cell = synthetic.SyntheticCell(3)
    kabs = cell.get_abs_kpts(ks)
    assert np.abs(f.get_phase_R2k(cell, kabs) - of.get_phase_R2k_scaled(kmesh, ks)).max() < 1e-13
    assert np.abs(f.get_phase_k2R(cell, kabs) - of.get_phase_R2k_scaled(kmesh, ks).conj().T / 6).max() < 1e-13
    assert f.get_kmesh(cell, kabs) == kmesh


@pytest.mark.parametrize("spin_h,spin_c", [(0, 0), (0, 2), (2, 0), (2, 2), (1, 2)])
def test_transform_h1_family(dev, spin_h, spin_c):
    from libdmet_preview_b200 import make_basis as mb, synthetic
    from oracle import make_basis as omb
    kmesh, nao, nlo = [1, 2, 3], 9, 7
    nk = 6
    rng = np.random.default_rng(10 * spin_h + spin_c)
    C = synthetic.make_C_ao_lo(kmesh, nao, nlo, seed=5, spin=spin_c if spin_c else None)
    h = _z(rng, *( (spin_h,) if spin_h else ()), nk, nao, nao)
    got = mb.transform_h1_to_lo(h, C)
    ref = omb.transform_h1_to_lo(h, C)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.abs(got - ref).max() < 1e-12
    # round trip h1_to_lo o h1_to_ao (test_make_basis.py:112-139) with a non-trivial overlap
    Cf = synthetic.make_C_ao_lo(kmesh, nao, nao, seed=6, spin=spin_c if spin_c else None)
    X = _z(rng, nk, nao, nao) * 0.1 + np.eye(nao)
    S = np.einsum("kpq,krq->kpr", X, X.conj())
    # C orthonormal w.r.t. S:  C = X^{-dagger} Q
    Cs = np.einsum("kpq,...kqr->...kpr", np.linalg.inv(X.conj().transpose(0, 2, 1)), Cf)
    hlo = _z(rng, *((spin_h,) if spin_h else ()), nk, nao, nao)
    hao = mb.transform_h1_to_ao(hlo, Cs, S)
    back = mb.transform_h1_to_lo(hao, Cs)
    assert np.abs(back - (hlo if back.ndim == hlo.ndim else np.broadcast_to(hlo, back.shape))).max() < 1e-11
    dlo = mb.transform_rdm1_to_lo(mb.transform_rdm1_to_ao(hlo, Cs), Cs, S)
    assert np.abs(dlo - (hlo if dlo.ndim == hlo.ndim else np.broadcast_to(hlo, dlo.shape))).max() < 1e-11


def test_h1_shortcuts_real_and_multiply_basis(dev):
    from libdmet_preview_b200 import make_basis as mb, synthetic
    from oracle import make_basis as omb
    kmesh = [1, 1, 3]
    C = synthetic.make_C_ao_lo(kmesh, 5, 4, seed=1, spin=2)
    assert np.array_equal(mb.transform_h1_to_lo(0.0, C[0]), omb.transform_h1_to_lo(0.0, C[0]))
    assert np.array_equal(mb.transform_h1_to_lo(np.array([0.0, 2.0]), C), omb.transform_h1_to_lo(np.array([0.0, 2.0]), C))
    rng = np.random.default_rng(0)
    hr, Cr = rng.standard_normal((3, 5, 5)), rng.standard_normal((3, 5, 4))
    got = mb.transform_h1_to_lo(hr, Cr)
    assert got.dtype == np.float64 and np.abs(got - omb.transform_h1_to_lo(hr, Cr)).max() < 1e-13
    b = rng.standard_normal((3, 4, 6))
    for a_, b_ in ((C, b), (C[0], b), (C, np.stack([b, 2 * b])), (C[0], np.stack([b, 2 * b]))):
        got = mb.multiply_basis(a_, b_)
        ref = omb.multiply_basis(a_, b_)
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-13
    with pytest.raises(ValueError):
        mb.multiply_basis(C[0, 0], b)
