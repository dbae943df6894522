"""GPU: `get_emb_eri` (GDF) through the public API against the oracle on the same seeded inputs; <= 1e-10 max-abs
 on O(1) integrals.  Cases follow the reference's own tests:
test_eri_transform_gdf.py (restricted, C_ao_eo re-entry, tiny chunks, time reversal vs plain, unit ERI),
test_eri_transform_uhf.py (aa, ab, bb), t_eri_transform_gdf_mpi.py (sharded == serial)."""
import numpy as np
import pytest

from helpers import problem

pytestmark = pytest.mark.gpu
TOL = 1e-10


def both(gdf, **kw):
    from libdmet_preview_b200 import eri_transform as et
    from oracle import eri_transform as oe
    okw = {k: v for k, v in kw.items() if k not in ("group", "kl_group", "source", "stats")}
    return et.get_emb_eri(gdf.cell, gdf, **kw), oe.get_emb_eri(gdf.cell, gdf, **okw)


@pytest.mark.parametrize("kmesh,nao,naux,neo", [
    ([1, 1, 3], 4, 10, 6),        # C1-like H chain
    ([3, 3, 1], 8, 30, 10),       # C2-like 2-D mesh
    ([2, 2, 2], 12, 40, 17),      # C3-like
    ([2, 2, 1], 7, 19, 5),        # odd sizes, neo < 8
    ([1, 1, 1], 5, 8, 5),         # Gamma only
    ([1, 2, 4], 26, 75, 33),
])
@pytest.mark.parametrize("trs", [True, False])
def test_restricted_s4(dev, kmesh, nao, naux, neo, trs):
    gdf, C, basis = problem(kmesh, nao, naux, neo)
    got, ref = both(gdf, C_ao_lo=C, basis=basis, t_reversal_symm=trs)
    assert got.shape == ref.shape == (1, neo * (neo + 1) // 2, neo * (neo + 1) // 2)
    assert got.dtype == np.float64 and got.flags.c_contiguous
    assert np.abs(got - ref).max() < TOL
    assert np.array_equal(got[0], got[0].T)             # both triangles returned, exactly mirrored


@pytest.mark.parametrize("symmetry", [1, 4, 8])
def test_symmetries_and_sources(dev, symmetry):
    gdf, C, basis = problem([1, 2, 3], 9, 21, 11)
    got, ref = both(gdf, C_ao_lo=C, basis=basis, symmetry=symmetry)
    assert got.shape == ref.shape and np.abs(got - ref).max() < TOL
    got_h, _ = both(gdf, C_ao_lo=C, basis=basis, symmetry=symmetry, source="host")
    assert np.array_equal(got_h, got)                   # host-staged blocks == device-generated blocks, bit for bit


@pytest.mark.parametrize("group,kl_group", [(1, 1), (2, 1), (3, 2), (7, 16)])
def test_grouping_is_only_a_schedule(dev, group, kl_group):
    gdf, C, basis = problem([2, 2, 2], 10, 24, 9)
    got, ref = both(gdf, C_ao_lo=C, basis=basis, group=group, kl_group=kl_group)
    assert np.abs(got - ref).max() < TOL


def test_unrestricted_blocks(dev):
    gdf, C, basis = problem([1, 1, 3], 6, 14, 7, spin=2)
    for trs in (True, False):
        got, ref = both(gdf, C_ao_lo=C, basis=basis, t_reversal_symm=trs)
        assert got.shape == ref.shape == (3, 28, 28)
        assert np.abs(got - ref).max() < TOL                          # order aa, ab, bb
    got1, ref1 = both(gdf, C_ao_lo=C, basis=basis, symmetry=1)
    assert got1.shape == (3, 7, 7, 7, 7) and np.abs(got1 - ref1).max() < TOL
    # restricted C with unrestricted basis and vice versa (add_spin_dim, eri_transform.py:283-286)
    got2, ref2 = both(gdf, C_ao_lo=C[0], basis=basis)
    assert got2.shape == (3, 28, 28) and np.abs(got2 - ref2).max() < TOL
    got3, ref3 = both(gdf, C_ao_lo=C, basis=basis[:1])
    assert np.abs(got3 - ref3).max() < TOL
    with pytest.raises(ValueError):
        both(gdf, C_ao_lo=C, basis=basis, symmetry=8)


def test_entry_variants(dev):
    from libdmet_preview_b200 import eri_transform as et
    from oracle import eri_transform as oe
    gdf, C, basis = problem([1, 1, 3], 6, 33, 8)
    e0 = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    # C_ao_eo re-entry (test_eri_transform_gdf.py:54, 1e-14 between two runs of the same algorithm)
    C_ao_eo = oe.build_C_ao_emb(gdf, C, basis) * (gdf.nkpts ** 0.75)
    e1 = et.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_eo=C_ao_eo)
    assert np.abs(e1 - e0).max() < 1e-13
    with pytest.raises(ValueError):
        et.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_lo=C, C_ao_eo=C_ao_eo)
    # the reference's chunk length does not change results (max_memory=0.02 -> 16-row chunks in the oracle)
    ref = oe.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, max_memory=0.02)
    assert np.abs(e0 - ref).max() < TOL
    # unit ERI (test_eri_transform_gdf.py:86-95)
    u = et.get_unit_eri(gdf.cell, gdf, C_ao_lo=C, symmetry=4)
    assert np.abs(u - oe.get_unit_eri(gdf.cell, gdf, C_ao_lo=C, symmetry=4)).max() < TOL
    u2 = et.get_unit_eri(gdf.cell, gdf, C_ao_lo=C, symmetry=4, t_reversal_symm=False)
    assert np.abs(u - u2).max() < TOL
    # k2gamma AO transformation: C_ao_lo and basis omitted (eri_transform.py:272-282)
    small, _, _ = problem([1, 1, 2], 3, 5, 3)
    g = et.get_emb_eri(small.cell, small)
    assert g.shape == (1, 21, 21) and np.abs(g - oe.get_emb_eri(small.cell, small)).max() < TOL
    # k-point centre shift only relabels the conservation test
    sh = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, kscaled_center=np.zeros(3))
    assert np.array_equal(sh, e0)
    with pytest.raises(ValueError):
        et.get_emb_eri(gdf.cell, object(), C_ao_lo=C, basis=basis)                  # unknown DF type
    with pytest.raises(NotImplementedError):
        et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, incore=False, t_reversal_symm=False)   # l.326-327


def test_device_tensors_in_and_out(dev):
    import torch
    from libdmet_preview_b200 import eri_transform as et
    gdf, C, basis = problem([2, 1, 2], 8, 16, 6)
    e0 = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    e1 = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=dev.to_device(C, torch.complex128),
                        basis=dev.to_device(basis, torch.float64), return_device=True)
    assert e1.is_cuda and np.array_equal(e1.cpu().numpy(), e0)


def test_properties_at_bench_shape_sample(dev):
    """size-independent properties at the target block shape (nao=200, naux=1000, neo=150) on a 1x1x2 mesh:
    pair-exchange symmetry, positive semi-definite diagonal, linearity in the GDF tensor (eri scales as scale^2),
    and a checksum of the blocks fed to the oracle on a slice."""
    from libdmet_preview_b200 import eri_transform as et, synthetic
    gdf, C, basis = problem([1, 1, 2], 200, 1000, 150)
    st = {}
    e = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, stats=st)
    assert e.shape == (1, 11325, 11325)
    assert np.array_equal(e[0], e[0].T)
    assert e[0].diagonal().min() >= 0.0
    g2 = synthetic.SyntheticGDF([1, 1, 2], 200, 1000, seed=gdf.seed, scale=2.0 * gdf.scale)
    e2 = et.get_emb_eri(g2.cell, g2, C_ao_lo=C, basis=basis)
    assert np.abs(e2 - 4.0 * e).max() < 1e-10 * max(1.0, np.abs(e2).max())
    assert st["launches"] > 0 and st["h2d_bytes"] == 0


@pytest.mark.parametrize("nsplit", [2, 3])
def test_aux_split_items(dev, nsplit):
    """(kL, aux-range) work items are independent and additive: splitting changes nothing beyond rounding"""
    gdf, C, basis = problem([1, 2, 2], 9, 29, 8)
    got, ref = both(gdf, C_ao_lo=C, basis=basis)
    from libdmet_preview_b200 import eri_transform as et
    for source in ("synth", "host"):
        sp = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, nsplit=nsplit, source=source)
        assert np.abs(sp - ref).max() < TOL


def test_gso_eri(dev):
    """GSO embedding ERI at a less trivial size against the oracle, incl. a basis with time-reversal structure
    for which the two schedules must agree (test_eri_transform_gso.py:155)."""
    from libdmet_preview_b200 import eri_transform as et
    from oracle import eri_transform as oe
    gdf, C, _ = problem([2, 1, 2], 7, 20, 4, spin=2)
    rng = np.random.default_rng(3)
    basis, _ = np.linalg.qr(rng.standard_normal((4 * 14, 9)))
    basis = basis.reshape(4, 14, 9)                 # real in R space -> time-reversal symmetric in k space
    got = et.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    ref = oe.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    assert got.shape == ref.shape == (1, 45, 45) and np.abs(got - ref).max() < TOL
    plain = et.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis=basis, t_reversal_symm=False)
    assert np.abs(plain - got).max() < TOL
    bk = np.einsum("Rim,Rk->kim", basis, __import__("oracle.fourier", fromlist=["x"]).get_phase_R2k_scaled(
        gdf.kmesh, gdf.kpts_scaled))
    via_k = et.get_emb_eri_gso(gdf.cell, gdf, C_ao_lo=C, basis_k=bk, nsplit=2, group=3)
    assert np.abs(via_k - ref).max() < TOL


def test_resident_gdf_cache(dev):
    """a GDF kept resident in HBM across calls (DMET iterations reuse the same tensor): first build fills the store,
    later builds read it in place; a budget smaller than the tensor falls back to streaming for the rest"""
    from libdmet_preview_b200 import eri_transform as et
    from oracle import eri_transform as oe
    gdf, C, basis = problem([1, 2, 2], 8, 22, 7)
    ref = oe.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)

    class HostOnly(object):          # hides the device generator: blocks come from host memory
        def __init__(self, g):
            self.g = g
            self.kpts_scaled, self.kmesh, self.nao, self.naux, self.cell = g.kpts_scaled, g.kmesh, g.nao, g.naux, g.cell
            self.loads = 0

        def load(self, ki, kj):
            self.loads += 1
            return self.g.load(ki, kj)

    for budget in (None, 5 * 22 * 8 * 8 * 16):          # everything resident / only 5 blocks fit
        host = HostOnly(gdf)
        res = et.ResidentGDF(host, max_bytes=budget)
        st1, st2 = {}, {}
        e1 = et.get_emb_eri(gdf.cell, res, C_ao_lo=C, basis=basis, stats=st1)
        n1 = host.loads
        e2 = et.get_emb_eri(gdf.cell, res, C_ao_lo=C, basis=basis * 1.0, stats=st2, nsplit=1)
        assert np.abs(e1 - ref).max() < TOL and np.abs(e2 - ref).max() < TOL
        if budget is None:
            assert host.loads == n1 and st2["h2d_bytes"] == 0          # second build touched no host block
        else:
            assert host.loads > n1 and 0 < st2["h2d_bytes"] < st1["h2d_bytes"] + 1
        res.release()
    syn = et.ResidentGDF(gdf)
    assert np.abs(et.get_emb_eri(gdf.cell, syn, C_ao_lo=C, basis=basis) - ref).max() < TOL
    # pairs that share a block (PooledGDF.block_key) share its resident copy: 3 slots serve the whole schedule
    from libdmet_preview_b200 import synthetic
    pooled = synthetic.PooledGDF(gdf, 3)
    refp = oe.get_emb_eri(gdf.cell, pooled, C_ao_lo=C, basis=basis)
    rp = et.ResidentGDF(pooled)
    for _ in range(2):
        assert np.abs(et.get_emb_eri(gdf.cell, rp, C_ao_lo=C, basis=basis) - refp).max() < TOL
    assert rp.bytes_cached == 3 * 22 * 8 * 8 * 16


def test_imaginary_part_diagnostic(dev):
    """without time reversal the reference forms the complex Lambda^dagger Lambda, logs max|imag| and warns above 1e-6
    (eri_transform.py:390-396).  Physical inputs give ~0; coefficients without time-reversal structure do not."""
    import warnings
    from libdmet_preview_b200 import eri_transform as et
    from oracle import eri_transform as oe
    gdf, C, basis = problem([1, 2, 2], 7, 15, 6, spin=2)          # odd naux exercises the padded panel columns
    st = {}
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        got = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, t_reversal_symm=False, stats=st)
    info = {}
    ref = oe.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_lo=C, basis=basis, t_reversal_symm=False, info=info)
    assert np.abs(got - ref).max() < TOL and st["eri_imag_norm"] < 1e-12 and info["eri_imag_norm"] < 1e-12
    rng = np.random.default_rng(0)
    C_bad = rng.standard_normal((2, 4, 7, 5)) + 1j * rng.standard_normal((2, 4, 7, 5))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = et.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_eo=C_bad, t_reversal_symm=False, stats=st)
    ref = oe.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_eo=C_bad, t_reversal_symm=False, info=info)
    assert np.abs(got - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
    assert info["eri_imag_norm"] > 1e-3
    assert abs(st["eri_imag_norm"] - info["eri_imag_norm"]) < 1e-10 * max(1.0, info["eri_imag_norm"])
    assert any("imaginary part" in str(x.message) for x in w)
    quiet = et.get_emb_eri_fast_gdf(gdf.cell, gdf, C_ao_eo=C_bad, t_reversal_symm=False, check_imag=False)
    assert np.array_equal(quiet, got)


@pytest.mark.parametrize("kmesh,nao,naux,neo", [([1, 1, 2], 120, 9, 210), ([1, 1, 1], 1, 1, 1), ([2, 1, 1], 3, 1, 2)])
def test_extreme_shapes(dev, kmesh, nao, naux, neo):
    """neo > 200 (two N tiles per GEMM), and degenerate one-orbital / one-auxiliary problems"""
    gdf, C, basis = problem(kmesh, nao, naux, neo)
    got, ref = both(gdf, C_ao_lo=C, basis=basis)
    assert got.shape == ref.shape and np.abs(got - ref).max() < TOL


def test_host_prefetch_thread(dev):
    """host blocks are loaded ahead on a background thread (the reference prefetches one chunk ahead,
    eri_transform.py:223); order and results are unchanged, provider errors surface in the caller"""
    import time
    from libdmet_preview_b200 import eri_transform as et
    gdf, C, basis = problem([1, 2, 2], 6, 12, 5)

    class Slow(object):
        def __init__(self, g, fail_at=None):
            self.g, self.calls, self.fail_at = g, [], fail_at
            self.kpts_scaled, self.kmesh, self.nao, self.naux, self.cell = g.kpts_scaled, g.kmesh, g.nao, g.naux, g.cell

        def load(self, ki, kj):
            self.calls.append((ki, kj))
            if self.fail_at is not None and len(self.calls) == self.fail_at:
                raise IOError("disk went away")
            time.sleep(0.002)
            return self.g.load(ki, kj)

    ref = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    slow = Slow(gdf)
    got = et.get_emb_eri(gdf.cell, slow, C_ao_lo=C, basis=basis)
    assert np.array_equal(got, ref) and len(slow.calls) == len(set(slow.calls)) == 10
    with pytest.raises(IOError):
        et.get_emb_eri(gdf.cell, Slow(gdf, fail_at=4), C_ao_lo=C, basis=basis)
    # the handle is usable again after the failed build
    assert np.array_equal(et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis), ref)


@pytest.mark.parametrize("nlo", [6, 4])
def test_gdf_in_lo_basis_gives_the_same_eri(dev, nlo, tmp_path):
    """transform_gdf_to_lo (eri_transform.py:1312-1407): oracle parity, the reference's own check that the
    time-reversal filling changes nothing (test_transform_gdf.py:104-108), and the defining property -- the
    embedding ERI from the LO-basis tensor with identity C equals the one from the AO tensor with C_ao_lo"""
    from libdmet_preview_b200 import eri_transform as et, synthetic
    from oracle import eri_transform as oe
    kmesh, nao, naux, neo = [2, 1, 3], 6, 13, 5
    gdf, _, _ = problem(kmesh, nao, naux, neo)
    C = synthetic.make_C_ao_lo(kmesh, nao, nlo, seed=5)
    basis = synthetic.make_emb_basis(kmesh, nlo, neo, seed=6)
    lo = et.transform_gdf_to_lo(gdf, C, fname=None)
    lo_plain = et.transform_gdf_to_lo(gdf, C, fname=None, t_reversal_symm=False)
    ref = oe.transform_gdf_to_lo(gdf, C)
    assert sorted(lo.j3c) == sorted(ref) == sorted(lo_plain.j3c)
    for k in ref:
        assert lo.j3c[k].shape == ref[k].shape and lo.j3c[k].dtype == ref[k].dtype
        assert np.abs(lo.j3c[k] - ref[k]).max() < TOL
        assert np.abs(lo.j3c[k] - lo_plain.j3c[k]).max() < TOL
    assert lo.nao == nlo and lo.cell.nao_nr() == nlo and gdf.cell.nao_nr() == nao
    eye = np.tile(np.eye(nlo, dtype=np.complex128), (len(gdf.kpts_scaled), 1, 1))
    a = et.get_emb_eri(lo.cell, lo, C_ao_lo=eye, basis=basis)
    b = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    assert np.abs(a - b).max() < TOL
    # file round trip in the reference's dataset naming
    f = str(tmp_path / "gdf_lo.npz")
    lo.save(f)
    z = np.load(f)
    assert z["j3c-kptij"].shape == (len(lo.kptij_idx), 2, 3)
    assert all(np.array_equal(z["j3c/%d/0" % k], v) for k, v in lo.j3c.items())
    # ... and as HDF5 (h5lite writer), served back through the file provider
    from libdmet_preview_b200.gdf_file import GDFFile
    h5 = str(tmp_path / "gdf_lo.h5")
    lo.save(h5)
    back = GDFFile(h5, cell=lo.cell, kpts=gdf.kpts)
    assert back.kptij_idx == lo.kptij_idx and np.array_equal(back.load(3, 1), lo.load(3, 1))
