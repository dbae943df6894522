"""CPU: GDF tensors on disk -- `gdf_file.GDFFile` serves PySCF-layout cderi files the way the reference reads them
(`get_naoaux` / `sr_loop` -> `_load3c`, eri_transform.py:159-227), checked against the oracle's restatement of that
access pattern, against the in-memory tensor the file was written from, and against golden results the reference's
own `get_emb_eri` / `transform_gdf_to_lo` produced from such a file (tests/golden/make_golden.py: gen_eri_file)."""
import os
import types

import numpy as np
import pytest

from libdmet_preview_b200 import synthetic, h5lite
from libdmet_preview_b200 import eri_transform as et
from libdmet_preview_b200.gdf_file import GDFFile, write_gdf_file
from oracle import eri_transform as oe
from helpers import write_case_file

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gdf_of(d):
    return synthetic.SyntheticGDF([int(x) for x in d["kmesh"]], int(d["nao"]), int(d["naux"]), seed=int(d["gdf_seed"]),
                                  scale=float(d["gdf_scale"]))


@pytest.mark.parametrize("version", ["v1", "v2"])
@pytest.mark.parametrize("nsegments,pack", [(1, True), (3, True), (2, False)])
def test_blocks_equal_the_tensor_written(tmp_path, version, nsegments, pack):
    gdf = synthetic.SyntheticGDF([2, 1, 3], 5, 11, seed=3)
    drop = {(1, 0): 9, (2, 2): 10}
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, version=version, nsegments=nsegments, pack_diagonal=pack,
                          naux_of=drop)
    with GDFFile(path, cell=gdf.cell) as f, h5lite.File(path) as raw:
        assert f.version == version and f.kmesh == [2, 1, 3] and f.nao == 5 and f.naux == 11 and f.aux_drop
        assert np.allclose(f.kpts_scaled, gdf.kpts_scaled) and np.allclose(f.kpts, gdf.kpts)
        assert f.kptij_idx == [(i, j) for i in range(6) for j in range(i + 1)]
        ora = oe.FileGDF(raw, gdf.cell, gdf.kpts)                 # the reference's access pattern, restated
        assert ora.naux == 11
        for i in range(6):
            for j in range(6):
                want = gdf.load(i, j).copy()
                rows = drop.get((i, j), drop.get((j, i), 11))
                want[rows:] = 0.0
                got = f.load(i, j)
                assert got.dtype == np.complex128 and got.flags.c_contiguous
                assert np.array_equal(got, want), (i, j)
                assert np.array_equal(ora.load(i, j), want), (i, j)
        # storage: Gamma block real, diagonal packed when asked, column segments
        g00 = raw["j3c/0"]
        assert len(g00) == nsegments and g00["0"].dtype == np.float64
        ncol = sum(g00[str(s)].shape[1] for s in range(nsegments))
        assert ncol == (15 if pack else 25)
        # a destination buffer (e.g. pinned memory) is filled in place
        buf = np.full((11, 5, 5), 7.0 + 7.0j)
        assert f.load(4, 1, out=buf) is buf and np.array_equal(buf, gdf.load(4, 1))
        assert f.load(0, 1, out=buf) is buf and np.array_equal(buf[:9], gdf.load(0, 1)[:9]) and not buf[9:].any()


def test_caller_kpoint_order_and_errors(tmp_path):
    gdf = synthetic.SyntheticGDF([1, 2, 2], 3, 4, seed=8)
    perm = [2, 0, 3, 1]
    for version in ("v1", "v2"):
        path = write_gdf_file(str(tmp_path / (version + ".h5")), gdf, version=version)
        f = GDFFile(path, cell=gdf.cell, kpts=gdf.kpts[perm])
        for a in range(4):
            for b in range(4):
                assert np.array_equal(f.load(a, b), gdf.load(perm[a], perm[b]))
        # lattice vectors instead of a cell: nao comes from the column count
        f2 = GDFFile(path, lattice_vectors=gdf.cell.lattice_vectors())
        assert f2.nao == 3 and f2.kmesh == [1, 2, 2]
        with pytest.raises(ValueError):
            GDFFile(path)
        with pytest.raises(ValueError):
            GDFFile(path, cell=gdf.cell, kpts=gdf.kpts + 0.01)
    # only some pairs stored: the others are served transposed, missing ones raise
    path = write_gdf_file(str(tmp_path / "few.h5"), gdf, pairs=[(0, 0), (1, 1), (2, 2), (3, 3), (0, 1), (3, 2)])
    f = GDFFile(path, cell=gdf.cell)
    assert np.array_equal(f.load(1, 0), gdf.load(1, 0)) and np.array_equal(f.load(2, 3), gdf.load(2, 3))
    with pytest.raises(KeyError):
        f.load(0, 2)
    # not a cderi file
    with h5lite.Writer(str(tmp_path / "other.h5")) as w:
        w["x"] = np.zeros(3)
    with pytest.raises(KeyError):
        GDFFile(str(tmp_path / "other.h5"), cell=gdf.cell)
    cell2d = synthetic.SyntheticCell(3, dimension=2)
    with pytest.raises(NotImplementedError):                      # eri_transform.py:226-227
        GDFFile(path, cell=cell2d)


def test_as_provider_accepts_a_gdf_like_object_with_a_cderi_path(tmp_path):
    gdf = synthetic.SyntheticGDF([1, 1, 3], 4, 6, seed=5)
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf)
    mydf = types.SimpleNamespace(_cderi=path, kpts=gdf.kpts, cell=gdf.cell, blockdim=120)
    prov = et.as_provider(gdf.cell, mydf)
    assert isinstance(prov, GDFFile) and prov.blockdim == 120 and prov.naux == 6
    assert np.array_equal(prov.load(2, 1), gdf.load(2, 1))
    assert et.as_provider(gdf.cell, gdf) is gdf
    with pytest.raises(ValueError):                               # eri_transform.py:89
        et.as_provider(gdf.cell, object())
    with pytest.raises(ValueError):
        et.as_provider(gdf.cell, types.SimpleNamespace(_cderi=str(tmp_path / "missing.h5"), kpts=gdf.kpts))


@pytest.mark.parametrize("name", ["eri_file_122", "eri_file_113"])
def test_oracle_from_file_vs_reference_python(tmp_path, name):
    """golden: the reference's own get_naoaux / get_emb_eri / transform_gdf_to_lo run over the same file"""
    d = np.load(os.path.join(G, name + ".npz"))
    gdf = _gdf_of(d)
    path = write_case_file(tmp_path / "cderi.h5", gdf, name)
    C, basis = d["C_ao_lo"], d["basis"]
    with GDFFile(path, cell=gdf.cell, kpts=gdf.kpts) as f, h5lite.File(path) as raw:
        assert f.naux == int(d["naoaux"]) == int(d["naux"])
        for prov in (f, oe.FileGDF(raw, gdf.cell, gdf.kpts)):
            for key, kw in (("s4_trs", {}), ("s4_plain", dict(t_reversal_symm=False)), ("s1_trs", dict(symmetry=1))):
                got = oe.get_emb_eri(gdf.cell, prov, C_ao_lo=C, basis=basis, **kw)
                assert got.shape == d[key].shape and np.abs(got - d[key]).max() < 1e-13, key
        lo = oe.transform_gdf_to_lo(f, d["C_lo"])
        assert np.allclose(d["lo_kptij"], [(gdf.kpts[i], gdf.kpts[j]) for i, j in f.kptij_idx])
        for k in range(len(d["lo_kptij"])):
            want = d["lo_%d" % k]
            assert lo[k].shape == want.shape and lo[k].dtype == want.dtype and np.abs(lo[k] - want).max() < 1e-13


def test_prefetcher_streams_file_blocks_through_a_buffer_ring(tmp_path):
    """run_items with a file provider: blocks are read on the background thread into a small ring of staging buffers
    (page-locked on a GPU box) that the consumer hands back after each host->device copy"""
    from libdmet_preview_b200.schedule import build_schedule, work_items
    gdf = synthetic.SyntheticGDF([1, 2, 2], 4, 10, seed=6)
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, nsegments=2)
    f = GDFFile(path, cell=gdf.cell)
    sch = build_schedule(f.kpts_scaled, True)

    class Recorder(object):
        def __init__(self):
            self.blocks, self.ptrs, self.kl = [], set(), 0

        def block_host(self, ki, kj, sym, L):
            self.blocks.append((ki, kj, np.array(L)))            # "copied to the device"
            self.ptrs.add(L.__array_interface__["data"][0])
            L[...] = np.nan                                         # a recycled buffer must be refilled completely

        def end_kl(self, weight):
            self.kl += 1

        def finish(self):
            self.done = True

    for nsplit in (1, 2):
        items = work_items(sch, f.naux, nsplit)
        for (l0, l1) in sorted({(a, b) for (_, a, b) in items}):
            sub = [it for it in items if (it[1], it[2]) == (l0, l1)]
            rec = Recorder()
            et.run_items(rec, f, sch, sub, source="auto", prefetch=2)
            want = [(ki, kj) for (u, _, _) in sub for (ki, kj, sym) in sch.units[u][2]]
            assert [(b[0], b[1]) for b in rec.blocks] == want and rec.kl == len(sub) and rec.done
            for ki, kj, L in rec.blocks:
                assert np.array_equal(L, gdf.load(ki, kj)[l0:l1])
            assert len(rec.ptrs) <= 4                                # prefetch depth 2 -> ring of 4 buffers
    assert len(f._staging[1]) == 4                                   # kept on the provider for the next call

    # a failing read surfaces in the consumer
    class Broken(GDFFile):
        def load(self, ki, kj, out=None):
            raise IOError("disk gone")
    b = Broken(path, cell=gdf.cell)
    with pytest.raises(IOError):
        et.run_items(Recorder(), b, sch, work_items(sch, b.naux, 1), source="host")


@pytest.mark.parametrize("version,nsegments,pack", [("v1", 1, True), ("v2", 3, True), ("v1", 2, False)])
def test_stored_entries_expand_to_the_blocks(tmp_path, version, nsegments, pack):
    """`load_stored` hands out the entries as they lie in the file (what crosses PCIe with DEVICE_UNPACK); their
    expansion -- host twin of ldm_unpack_stored -- equals `load`, for every pair, orientation and aux range"""
    from libdmet_preview_b200.gdf_file import STORED_SWAPPED, STORED_REAL
    gdf = synthetic.SyntheticGDF([2, 1, 3], 5, 11, seed=3)
    drop = {(1, 0): 9, (2, 2): 10, (4, 3): 3}
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, version=version, nsegments=nsegments, pack_diagonal=pack,
                          naux_of=drop)
    f = GDFFile(path, cell=gdf.cell)
    buf = np.empty((11, 5, 5), dtype=np.complex128)
    for (l0, l1) in ((0, 11), (0, 6), (6, 11), (4, 5)):
        for i in range(6):
            for j in range(6):
                buf[...] = np.nan
                e = f.load_stored(i, j, l0, l1, buf)
                assert e.data.size == 0 or np.shares_memory(e.data, buf)
                assert bool(e.flags & STORED_SWAPPED) == (i < j)
                assert bool(e.flags & STORED_REAL) == (i == 0 and j == 0)
                if i == j and pack:
                    assert e.data.shape[1] == 15
                else:
                    assert e.data.shape[1] == 25
                rows_stored = drop.get((i, j), drop.get((j, i), 11))
                assert e.data.shape[0] == max(0, min(l1, rows_stored) - min(l0, rows_stored))
                want = f.load(i, j)[l0:l1]
                assert np.array_equal(e.expand(l1 - l0, 5), want), (i, j, l0, l1)


def test_run_items_ships_stored_entries_when_the_build_can_unpack(tmp_path):
    from libdmet_preview_b200.schedule import build_schedule, work_items
    gdf = synthetic.SyntheticGDF([1, 2, 2], 4, 10, seed=6)
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, nsegments=2, naux_of={(3, 1): 7})
    f = GDFFile(path, cell=gdf.cell)
    sch = build_schedule(f.kpts_scaled, True)

    class Recorder(object):
        def __init__(self, naux):
            self.naux, self.blocks, self.bytes = naux, [], 0

        def block_stored(self, ki, kj, sym, entry):
            self.blocks.append((ki, kj, entry.expand(self.naux, 4)))
            self.bytes += entry.data.nbytes
            entry.data[...] = np.nan

        def block_host(self, *a):
            raise AssertionError("expanded block shipped although the build unpacks on the device")

        def end_kl(self, weight):
            pass

        def finish(self):
            pass

    for nsplit in (1, 2):
        items = work_items(sch, f.naux, nsplit)
        for (l0, l1) in sorted({(a, b) for (_, a, b) in items}):
            sub = [it for it in items if (it[1], it[2]) == (l0, l1)]
            rec = Recorder(l1 - l0)
            et.run_items(rec, f, sch, sub)
            want = [(ki, kj) for (u, _, _) in sub for (ki, kj, sym) in sch.units[u][2]]
            assert [(b[0], b[1]) for b in rec.blocks] == want
            for ki, kj, L in rec.blocks:
                full = gdf.load(ki, kj).copy()
                if (ki, kj) in ((3, 1), (1, 3)):
                    full[7:] = 0
                assert np.array_equal(L, full[l0:l1])
            assert rec.bytes < len(want) * (l1 - l0) * 16 * 16       # packed / real / dropped rows: fewer bytes


def _trs_reduced_pairs(gdf):
    """one member of every time-reversal class {(i, j), (j, i), (-i, -j), (-j, -i)} of k-point pairs"""
    nk = len(gdf.kpts_scaled)
    seen, keep = set(), []
    for i in range(nk):
        for j in range(nk):
            if (i, j) in seen:
                continue
            mi, mj = gdf.minus[i], gdf.minus[j]
            seen.update({(i, j), (j, i), (mi, mj), (mj, mi)})
            keep.append((i, j))
    return keep


@pytest.mark.parametrize("version", ["v1", "v2"])
def test_time_reversal_reduced_file(tmp_path, version):
    """a cderi file that keeps one member of each time-reversal class of pairs (what recent PySCF writes): a missing
    pair is served as the conjugate of (-k_i, -k_j), or the transpose of (-k_j, -k_i) -- on the host route (`load`) and
    as a stored entry + flags for the device unpacker (`load_stored` / `StoredEntry.expand`)"""
    gdf = synthetic.SyntheticGDF([2, 1, 3], 5, 9, seed=21)
    pairs = _trs_reduced_pairs(gdf)
    nk = len(gdf.kpts_scaled)
    assert len(pairs) < nk * (nk + 1) // 2
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, version=version, pairs=pairs, nsegments=2)
    f = GDFFile(path, cell=gdf.cell, kpts=gdf.kpts)
    buf = np.empty((gdf.naux, gdf.nao, gdf.nao), dtype=np.complex128)
    flags_seen = set()
    for i in range(nk):
        for j in range(nk):
            want = gdf.load(i, j)
            assert np.array_equal(f.load(i, j), want)
            e = f.load_stored(i, j, 0, gdf.naux, buf)
            flags_seen.add(e.flags & 5)
            assert np.array_equal(e.expand(gdf.naux, gdf.nao), want)
    assert flags_seen == {0, 1, 4, 5}
    # the whole pipeline's host logic over such a file (oracle arithmetic)
    C = synthetic.make_C_ao_lo(gdf.kmesh, gdf.nao, seed=3)
    basis = synthetic.make_emb_basis(gdf.kmesh, gdf.nao, 4, seed=4)
    assert np.abs(oe.get_emb_eri(gdf.cell, f, C_ao_lo=C, basis=basis) -
                  oe.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)).max() < 1e-13
