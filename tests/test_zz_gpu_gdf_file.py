"""GPU: the embedding ERI built from a cderi FILE (gdf_file.GDFFile -> host blocks -> ldm_eri_block_host) against the
oracle, against golden results of the reference's own Python over the same file, and the LO-basis tensor written to
and served from HDF5."""
import os
import types

import numpy as np
import pytest

from libdmet_preview_b200 import synthetic
from helpers import problem, write_case_file

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10      # BASELINE.json north_star: <= 1e-10 max-abs on embedding integrals


@pytest.mark.parametrize("name", ["eri_file_122", "eri_file_113"])
def test_get_emb_eri_from_file_vs_reference_python(dev, tmp_path, name):
    from libdmet_preview_b200 import eri_transform as et
    d = np.load(os.path.join(G, name + ".npz"))
    gdf = synthetic.SyntheticGDF([int(x) for x in d["kmesh"]], int(d["nao"]), int(d["naux"]), seed=int(d["gdf_seed"]),
                                 scale=float(d["gdf_scale"]))
    path = write_case_file(tmp_path / "cderi.h5", gdf, name)
    mydf = types.SimpleNamespace(_cderi=path, kpts=gdf.kpts, cell=gdf.cell)       # what a PySCF GDF exposes
    for key, kw in (("s4_trs", {}), ("s4_plain", dict(t_reversal_symm=False)), ("s1_trs", dict(symmetry=1))):
        got = et.get_emb_eri(gdf.cell, mydf, C_ao_lo=d["C_ao_lo"], basis=d["basis"], **kw)
        assert got.shape == d[key].shape and np.abs(got - d[key]).max() < TOL, key
    # LO-basis tensor: written as HDF5 by the product, compared dataset by dataset with the reference's file
    lo_path = str(tmp_path / "lo.h5")
    mydf_lo = et.transform_gdf_to_lo(mydf, d["C_lo"], fname=lo_path)
    assert mydf_lo._cderi == lo_path and mydf_lo.cell.nao_nr() == d["C_lo"].shape[-1]
    from libdmet_preview_b200 import h5lite
    with h5lite.File(lo_path) as f:
        assert np.allclose(f["j3c-kptij"][...], d["lo_kptij"])
        for k in range(len(d["lo_kptij"])):
            x = f["j3c/%d/0" % k][...]
            want = d["lo_%d" % k]
            assert x.shape == want.shape and x.dtype == want.dtype and np.abs(x - want).max() < TOL, k


@pytest.mark.parametrize("version,nsegments,pack", [("v1", 2, True), ("v2", 1, False)])
def test_file_resident_and_lo_round_trip(dev, tmp_path, version, nsegments, pack):
    """file provider == in-memory provider (bitwise: same blocks, same kernels); ResidentGDF over a file;
    ERI from the LO-basis file with C = identity equals the ERI from the AO tensor (the reference's own check,
    test_eri_transform_gdf.py)"""
    from libdmet_preview_b200 import eri_transform as et
    from libdmet_preview_b200.gdf_file import GDFFile, write_gdf_file
    from oracle import eri_transform as oe
    kmesh, nao, naux, neo = [2, 1, 2], 6, 13, 7
    gdf, C, basis = problem(kmesh, nao, naux, neo)
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, version=version, nsegments=nsegments, pack_diagonal=pack)
    f = GDFFile(path, cell=gdf.cell, kpts=gdf.kpts)
    ref = oe.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    mem = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, source="host")
    got = et.get_emb_eri(gdf.cell, f, C_ao_lo=C, basis=basis)
    assert np.abs(got - ref).max() < TOL and np.array_equal(got, mem)
    res = et.ResidentGDF(f)
    for _ in range(2):                                            # second call reads the device store only
        assert np.abs(et.get_emb_eri(gdf.cell, res, C_ao_lo=C, basis=basis) - ref).max() < TOL
    res.release()
    lo_path = str(tmp_path / "lo.h5")
    lo = et.transform_gdf_to_lo(f, C, fname=lo_path)              # a provider came in -> a provider comes back
    assert isinstance(lo, et.LoGDF)
    flo = GDFFile(lo_path, cell=lo.cell, kpts=gdf.kpts)
    eye = np.asarray([np.eye(nao, dtype=np.complex128)] * len(gdf.kpts))
    via_lo = et.get_emb_eri(lo.cell, flo, C_ao_lo=eye, basis=basis)
    assert np.abs(via_lo - ref).max() < TOL


@pytest.mark.parametrize("spin", [1, 2])
def test_outcore_eri_file(dev, tmp_path, spin):
    """incore=False: the s4 ERI lands in dataset "ccdd" of `fout`, spin blocks ordered aa, bb, ab
    (eri_transform.py:308, 311-320, 486-521), and the open file is returned"""
    from libdmet_preview_b200 import eri_transform as et
    gdf, C, basis = problem([1, 2, 2], 5, 11, 6, spin=spin)
    fout = str(tmp_path / "H2.h5")
    mem = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    f = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, incore=False, fout=fout)
    assert f.keys() == ["ccdd"] and os.path.exists(fout)
    got = f["ccdd"][...]
    assert got.dtype == np.float64 and got.shape == mem.shape
    assert np.array_equal(got, mem if spin == 1 else mem[[0, 2, 1]])
    f.close()


@pytest.mark.parametrize("nao,naux,rows", [(5, 7, 7), (33, 6, 4), (70, 3, 0), (64, 5, 5)])
def test_unpack_stored_kernel(dev, nao, naux, rows):
    """ldm_unpack_stored against its host twin (gdf_file.StoredEntry.expand): full / Hermitian-packed, complex /
    real, stored / swapped pair, fewer stored rows than naux -- bit for bit (pure data movement + sign flips)"""
    import torch
    from libdmet_preview_b200.gdf_file import StoredEntry, STORED_SWAPPED, STORED_CONJ
    rng = np.random.default_rng(nao)
    for ncols in (nao * nao, nao * (nao + 1) // 2):
        for real in (False, True):
            a = rng.standard_normal((rows, ncols))
            if not real:
                a = a + 1j * rng.standard_normal((rows, ncols))
            for flags in (0, STORED_SWAPPED, STORED_CONJ, STORED_SWAPPED | STORED_CONJ):
                want = StoredEntry(a, flags | (2 if real else 0)).expand(naux, nao)
                out = dev.empty((naux, nao, nao), torch.complex128)
                out.fill_(float("nan"))
                got = dev.unpack_stored(dev.to_device(a), naux, nao, flags, out=out)
                assert np.array_equal(got.cpu().numpy(), want), (ncols, real, flags)


def test_host_and_device_unpack_agree(dev, tmp_path, monkeypatch):
    """DEVICE_UNPACK on (stored entries shipped, ldm_eri_block_stored) and off (blocks assembled with numpy,
    ldm_eri_block_host) give the same bits; the device path moves fewer bytes over PCIe"""
    from libdmet_preview_b200 import eri_transform as et
    from libdmet_preview_b200.gdf_file import GDFFile, write_gdf_file
    gdf, C, basis = problem([1, 2, 2], 9, 14, 8)
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, nsegments=2, naux_of={(2, 1): 11})
    f = GDFFile(path, cell=gdf.cell, kpts=gdf.kpts)
    st_dev, st_host = {}, {}
    assert et.DEVICE_UNPACK
    e_dev = et.get_emb_eri(gdf.cell, f, C_ao_lo=C, basis=basis, stats=st_dev)
    e_split = et.get_emb_eri(gdf.cell, f, C_ao_lo=C, basis=basis, nsplit=2)
    monkeypatch.setattr(et, "DEVICE_UNPACK", False)
    e_host = et.get_emb_eri(gdf.cell, f, C_ao_lo=C, basis=basis, stats=st_host)
    assert np.array_equal(e_dev, e_host) and np.abs(e_split - e_host).max() < TOL
    assert 0 < st_dev["h2d_bytes"] < st_host["h2d_bytes"]


def test_time_reversal_reduced_file_on_the_device(dev, tmp_path):
    """a cderi file holding one member of each time-reversal class of pairs: missing pairs are served from
    (-k_i, -k_j) / (-k_j, -k_i) with LDM_STORED_CONJ, unpacked on the device -- same ERI as from memory"""
    from libdmet_preview_b200 import eri_transform as et
    from libdmet_preview_b200.gdf_file import GDFFile, write_gdf_file
    gdf, C, basis = problem([2, 1, 3], 7, 12, 6)
    nk = len(gdf.kpts_scaled)
    seen, pairs = set(), []
    for i in range(nk):
        for j in range(nk):
            if (i, j) not in seen:
                mi, mj = gdf.minus[i], gdf.minus[j]
                seen.update({(i, j), (j, i), (mi, mj), (mj, mi)})
                pairs.append((i, j))
    path = write_gdf_file(str(tmp_path / "cderi.h5"), gdf, pairs=pairs)
    f = GDFFile(path, cell=gdf.cell, kpts=gdf.kpts)
    ref = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, source="host")
    for trs in (True, False):
        got = et.get_emb_eri(gdf.cell, f, C_ao_lo=C, basis=basis, t_reversal_symm=trs)
        assert np.abs(got - ref).max() < TOL
    res = et.ResidentGDF(f)
    assert np.abs(et.get_emb_eri(gdf.cell, res, C_ao_lo=C, basis=basis) - ref).max() < TOL
    f.close()


def test_rho_glob_vs_reference_python(dev):
    """global density matrix by democratic partitioning (slater_helper.py:183-283): one batched device product
    against golden results of the reference's own loop; k-space form through the device R2k"""
    from libdmet_preview_b200 import slater, lattice as plat
    from oracle import fourier as of
    d = np.load(os.path.join(G, "rho_glob.npz"))

    def lat(kmesh, nlo, imp):
        L = plat.Lattice(synthetic.SyntheticCell(nlo), kmesh)
        L.set_val_virt_core([int(x) for x in imp], [], [])
        return L

    for tag in ("r", "u"):
        km = [int(x) for x in d["kmesh_" + tag]]
        basis, rho, want = d["basis_" + tag], d["rho_" + tag], d["glob_" + tag]
        L = lat(km, basis.shape[2], d["imp_" + tag])
        got = slater.get_rho_glob_R(basis, L, rho)
        assert got.shape == want.shape and got.dtype == np.float64
        assert np.abs(got - want).max() < TOL
        assert np.abs(slater.get_rho_glob_k(basis, L, rho) - of.R2k(want, km)).max() < TOL
    lats = [lat([1, 1, 3], 5, [0, 1]), lat([1, 1, 3], 5, [2, 3, 4])]
    got = slater.get_rho_glob_R([d["basis_f0"], d["basis_f1"]], lats, [d["rho_f0"], d["rho_f1"]])
    assert np.abs(got - d["glob_f"]).max() < TOL


@pytest.mark.parametrize("name", ["gso_embham_113", "gso_embham_221"])
def test_gso_energy_side_vs_reference_python(dev, name):
    """GSO get_H_dmet / transformResults (spinless.py:754-848, 948-1035) on the Hamiltonian the device built, against
    golden results of the reference's own code"""
    from helpers import GSOLattice
    from libdmet_preview_b200 import spinless, fourier
    d = np.load(os.path.join(G, name + ".npz"))
    gdf = synthetic.SyntheticGDF([int(x) for x in d["kmesh"]], int(d["nao"]), int(d["naux"]), seed=int(d["gdf_seed"]),
                                 scale=float(d["gdf_scale"]))
    Lat = GSOLattice(gdf, d["C_ao_lo"], fourier, eri_symmetry=int(d["sym"]))
    mu = float(d["mu"])
    Ham, _ = spinless.get_emb_Ham(Lat, d["basis"], None, mu)
    H2_before = Ham.H2["ccdd"].copy()
    Hd = spinless.get_H_dmet(d["basis"], Lat, Ham, last_dmu=0.1, mu=mu)
    assert Hd.H1["cd"].shape == d["Hd_H1"].shape and np.abs(Hd.H1["cd"] - d["Hd_H1"]).max() < TOL
    assert Hd.H2["ccdd"].shape == d["Hd_H2"].shape and np.abs(Hd.H2["ccdd"] - d["Hd_H2"]).max() < TOL
    assert Hd.H0 == float(d["Hd_H0"]) and np.array_equal(Ham.H2["ccdd"], H2_before)
    Hd1 = spinless.get_H_dmet(d["basis"], Lat, Ham, last_dmu=0.1, mu=mu, compact=False)
    assert Hd1.H2["ccdd"].shape == d["Hd_H2_s1"].shape and np.abs(Hd1.H2["ccdd"] - d["Hd_H2_s1"]).max() < TOL
    GRhoImp, Efrag, nelec = spinless.transformResults(d["GRhoEmb"], -2.75, Lat, d["basis"], Ham, None, mu,
                                                      last_dmu=0.1)
    assert np.abs(GRhoImp - d["GRhoImp"]).max() < TOL and abs(Efrag - float(d["Efrag"])) < 1e-8
    assert abs(nelec - float(d["nelec"])) < TOL


@pytest.mark.parametrize("tag", ["r", "u"])
def test_update_Ham_vs_reference_python(dev, tag):
    """Lattice.update_Ham (lattice.py:565-589): DMET density matrix -> k space -> AO basis -> every LO-basis
    quantity rebuilt, all on the device, against golden results of the reference's own Lattice"""
    import warnings
    from libdmet_preview_b200 import lattice as plat
    d = np.load(os.path.join(G, "lattice_misc.npz"))
    km, nao = [int(x) for x in d["kmesh_" + tag]], int(d["nao_" + tag])
    L = plat.Lattice(synthetic.SyntheticCell(nao), km)
    L.set_Ham(None, None, d["C_" + tag], ovlp=d["ovlp_" + tag], hcore=d["hcore_" + tag], rdm1=d["rdm1_" + tag],
              vhf=d["vhf_" + tag], H0=1.5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")              # the perturbed density matrix is not exactly real in R space
        L.update_Ham(d["new_R_" + tag], vhf=d["new_vhf_" + tag])
    for k in ("rdm1_ao_k", "rdm1_lo_k", "rdm1_lo_R", "fock_lo_k", "fock_hf_lo_k", "vhf_lo_R", "veff_lo_k",
              "hcore_lo_k"):
        got, want = np.asarray(getattr(L, k)), d["%s_%s" % (k, tag)]
        assert got.shape == want.shape and np.abs(got - want).max() < TOL, k
    assert L.H0 == 1.5 and L.has_Ham
