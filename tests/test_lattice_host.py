"""CPU: host logic of `lattice.Lattice` that needs no device (cell arithmetic, `transpose`, `expand_orb`) and of
`update_Ham` (lattice.py:565-589) with the device transforms replaced by numpy restatements, against golden results
of the reference's own Lattice (tests/golden/make_golden.py: gen_lattice_misc)."""
import os

import numpy as np
import pytest

from libdmet_preview_b200 import synthetic
from libdmet_preview_b200 import lattice as plat, fourier as pf, make_basis as pmb
from oracle import fourier as of, make_basis as omb

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _lat(d, tag):
    km, nao = [int(x) for x in d["kmesh_" + tag]], int(d["nao_" + tag])
    return plat.Lattice(synthetic.SyntheticCell(nao), km), km, nao


@pytest.mark.parametrize("tag", ["r", "u"])
def test_transpose_and_expand_orb_vs_reference_python(tag):
    d = np.load(os.path.join(G, "lattice_misc.npz"))
    L, km, nao = _lat(d, tag)
    A = d["new_R_" + tag]
    assert np.array_equal(L.transpose(A), d["transpose_" + tag])
    assert np.array_equal(L.transpose(A[0]), d["transpose_" + tag][0])
    assert np.array_equal(L.expand_orb(A), d["expand_orb_" + tag])
    assert np.array_equal(L.expand_orb(A[0]), d["expand_orb_" + tag][0])
    assert np.array_equal(L.transpose(L.transpose(A)), A)
    # expand_orb of a square stripe is `expand` (lattice.py:304-337 vs 353-377)
    assert np.array_equal(L.expand_orb(A), L.expand(A, dense=True))
    with pytest.raises(ValueError):
        L.transpose(A[0, 0])
    with pytest.raises(ValueError):
        L.expand_orb(A[0, 0])


def _inv_transform(x, C, S):
    """C^-1 x C^-dagger with C^-1 = C^dagger S (make_basis.py:560-620), spin broadcast like the reference"""
    x, C = np.asarray(x), np.asarray(C)
    x4 = x if x.ndim == 4 else x[None]
    C4 = C if C.ndim == 4 else C[None]
    spin = max(x4.shape[0], C4.shape[0])
    out = np.zeros((spin, x4.shape[1], C4.shape[-1], C4.shape[-1]), dtype=np.complex128)
    for s in range(spin):
        for k in range(x4.shape[1]):
            Ci = C4[min(s, C4.shape[0] - 1), k].conj().T.dot(S[k])
            out[s, k] = Ci.dot(x4[min(s, x4.shape[0] - 1), k]).dot(Ci.conj().T)
    return out[0] if (x.ndim == 3 and C.ndim == 3) else out


def _to_ao(x, C):
    x, C = np.asarray(x), np.asarray(C)
    x4 = x if x.ndim == 4 else x[None]
    C4 = C if C.ndim == 4 else C[None]
    spin = max(x4.shape[0], C4.shape[0])
    out = np.asarray([[C4[min(s, C4.shape[0] - 1), k].dot(x4[min(s, x4.shape[0] - 1), k]).dot(
        C4[min(s, C4.shape[0] - 1), k].conj().T) for k in range(x4.shape[1])] for s in range(spin)])
    return out[0] if (x.ndim == 3 and C.ndim == 3) else out


@pytest.mark.parametrize("tag", ["r", "u"])
def test_update_Ham_host_logic_vs_reference_python(tag, monkeypatch):
    """which quantities `update_Ham` replaces, which it keeps (hcore, ovlp, the stored HF potential when only veff
    is passed), shapes and spin axes -- device transforms restated with numpy; the device versions run the same
    fixture in tests/test_zz_gpu_gdf_file.py"""
    d = np.load(os.path.join(G, "lattice_misc.npz"))
    L, km, nao = _lat(d, tag)
    monkeypatch.setattr(pf, "R2k", of.R2k)
    monkeypatch.setattr(pf, "k2R", lambda A, kmesh, tol=1e-7: of.k2R(A, kmesh, tol=np.inf))
    monkeypatch.setattr(pmb, "transform_h1_to_lo", omb.transform_h1_to_lo)
    monkeypatch.setattr(pmb, "transform_rdm1_to_lo", _inv_transform)
    monkeypatch.setattr(pmb, "transform_rdm1_to_ao", _to_ao)
    L.set_Ham(None, None, d["C_" + tag], ovlp=d["ovlp_" + tag], hcore=d["hcore_" + tag], rdm1=d["rdm1_" + tag],
              vhf=d["vhf_" + tag], H0=1.5)
    hcore_before = L.hcore_lo_k.copy()
    L.update_Ham(d["new_R_" + tag], vhf=d["new_vhf_" + tag])
    for k in ("rdm1_ao_k", "rdm1_lo_k", "rdm1_lo_R", "fock_lo_k", "fock_hf_lo_k", "vhf_lo_R", "veff_lo_k",
              "hcore_lo_k"):
        got, want = np.asarray(getattr(L, k)), d["%s_%s" % (k, tag)]
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-12, k
    assert np.array_equal(L.hcore_lo_k, hcore_before) and L.H0 == 1.5 and L.has_Ham
    # only veff supplied: the HF potential of the last set_Ham stays, fock follows veff
    L.update_Ham(d["new_R_" + tag], veff=d["vhf_" + tag] * 0.5)
    assert np.abs(L.vhf_ao_k - d["new_vhf_" + tag]).max() == 0
    assert np.abs(L.fock_ao_k - (d["hcore_" + tag] + d["vhf_" + tag] * 0.5)).max() < 1e-15
    with pytest.raises(ValueError):
        L.update_Ham(d["new_R_" + tag])             # J and K would have to be rebuilt: needs the mean-field object
