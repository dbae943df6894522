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


class _HostStandIn(object):
    """stand-in for `device.Device` with the five entry points `transform_gdf_to_lo` drives, each restated with
    torch CPU tensors from the C-ABI documentation (include/ldm_b200.h) -- exercises the HOST logic of that function
    (pair loop, time-reversal filling, storage rules, stored-entry route, HDF5 output) without a GPU.  The CUDA
    kernels themselves are checked in tests/test_gpu_*.py."""
    import torch as _torch
    torch_device = _torch.device("cpu")

    def empty(self, shape, dtype):
        import torch
        return torch.empty(shape, dtype=dtype)

    def ztranspose(self, x, conj=False, scale=1.0):
        return x.transpose(-1, -2).contiguous()

    def synchronize(self):
        pass

    def to_host(self, t):
        return t.numpy().copy()

    def pack_tril(self, x, out_real=False):
        """ldm_pack_tril: out[L][m(m+1)/2 + c] = x[L][m][c], c <= m; real part + max|imag| when out_real"""
        import torch
        r, c = np.tril_indices(x.shape[-1])
        p = x.numpy()[:, r, c]
        if out_real:
            return torch.from_numpy(np.ascontiguousarray(p.real)), float(np.abs(p.imag).max())
        return torch.from_numpy(np.ascontiguousarray(p)), None

    def unpack_stored(self, src, naux, nao, flags=0, out=None):
        import torch
        from libdmet_preview_b200.gdf_file import StoredEntry
        if src.dtype == torch.float64:
            flags |= 2
        out.copy_(torch.from_numpy(StoredEntry(src.numpy(), flags).expand(naux, nao)))
        return out

    def zgemm_tn(self, A, B, segs, C_out, rdiv=1, s_outer=None, s_inner=0, s_col=1, **kw):
        """C[(r / rdiv) * s_outer + (r % rdiv) * s_inner + n * s_col] = sum_k A[za][r][k] * op(B[zb][n][k])"""
        import torch
        (za, zb, _, conj), = segs
        Bm = B[zb].conj() if conj else B[zb]
        res = A[za] @ Bm.T
        r = torch.arange(A.shape[1])
        n = torch.arange(Bm.shape[0])
        idx = ((r // rdiv) * s_outer + (r % rdiv) * s_inner)[:, None] + n[None, :] * s_col
        C_out.reshape(-1)[idx.reshape(-1)] = res.reshape(-1)


@pytest.mark.parametrize("device_unpack", [True, False])
def test_transform_gdf_to_lo_host_logic_from_a_file(tmp_path, monkeypatch, device_unpack):
    import torch
    from libdmet_preview_b200 import h5lite
    from libdmet_preview_b200.gdf_file import GDFFile, write_gdf_file
    monkeypatch.setattr(et, "get_device", lambda *a: _HostStandIn())
    monkeypatch.setattr(et, "_to_z", lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128)))
    monkeypatch.setattr(et, "DEVICE_UNPACK", device_unpack)
    kmesh, nao, nlo, naux = [2, 1, 2], 6, 5, 13
    gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=21)
    C = synthetic.make_C_ao_lo(kmesh, nao, nlo, seed=22)
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, nsegments=2, naux_of={(2, 1): 9})
    f = GDFFile(path, cell=gdf.cell, kpts=gdf.kpts)
    out = str(tmp_path / "lo.h5")
    lo = et.transform_gdf_to_lo(f, C, fname=out)
    ref = oe.transform_gdf_to_lo(f, C)
    assert isinstance(lo, et.LoGDF) and sorted(lo.j3c) == sorted(ref)
    with h5lite.File(out) as h:
        for k, v in ref.items():
            x = h["j3c/%d/0" % k][...]
            assert x.dtype == v.dtype and x.shape == v.shape and np.abs(x - v).max() < 1e-13
    # a GDF-like object (path + kpts + cell) comes back as a GDF-like object pointing at the new file
    import types
    mydf = types.SimpleNamespace(_cderi=path, kpts=gdf.kpts, cell=gdf.cell)
    mydf_lo = et.transform_gdf_to_lo(mydf, C, fname=str(tmp_path / "lo2.h5"))
    assert mydf_lo._cderi.endswith("lo2.h5") and mydf_lo.cell.nao_nr() == nlo and mydf.cell.nao_nr() == nao
