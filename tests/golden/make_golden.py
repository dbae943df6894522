"""Generate the golden fixtures of tests/golden/*.npz by running the reference's OWN Python, imported unmodified from
/root/reference, over the stub environment of ref_stubs.py (PySCF / h5py are not installed; see that file for what is
stubbed).  Run once in the build container:

    python tests/golden/make_golden.py

Functions executed from the reference:
    libdmet.basis_transform.eri_transform.get_emb_eri / get_unit_eri          (eri_transform.py:44-112, 235-399)
    libdmet.system.fourier.R2k / k2R                                           (fourier.py:129-177)
    libdmet.basis_transform.make_basis.transform_h1_to_lo / multiply_basis     (make_basis.py:524-558, 923-962)
    libdmet.system.lattice.Lattice(...).set_Ham(...)                           (lattice.py:31-56, 416-515, 591-673)
    libdmet.routine.slater.get_emb_basis / get_emb_Ham                         (slater.py:98-220, 320-370, ...)
    libdmet.basis_transform.eri_transform.transform_gdf_to_lo                  (eri_transform.py:1312-1427)
Inputs are the seeded synthetic problems of tests/helpers.py; small inputs are stored next to the outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_stubs  # noqa: E402

ref_stubs.install()

from helpers import problem, mean_field  # noqa: E402
from libdmet_preview_b200 import synthetic  # noqa: E402
from pyscf.pbc import df as stub_df  # noqa: E402
import libdmet.basis_transform.eri_transform as ref_eri  # noqa: E402
import libdmet.basis_transform.make_basis as ref_mb  # noqa: E402
import libdmet.system.fourier as ref_fourier  # noqa: E402
import libdmet.system.lattice as ref_lattice  # noqa: E402
import libdmet.routine.slater as ref_slater  # noqa: E402
from libdmet.utils import logger as ref_log  # noqa: E402

ref_log.verbose = "RESULT"


class GoldenCell(synthetic.SyntheticCell):
    """adds the members the reference's Lattice constructor touches (lattice.py:33-56)"""
    pbc_intor = True

    def __init__(self, nao):
        super().__init__(nao)
        self._atom = [("H", (0.0, 0.0, 0.0))]

    def super_cell(self, kmesh):
        big = GoldenCell(self._nao * int(np.prod(kmesh)))
        return big

    def copy(self):
        import copy
        return copy.copy(self)


def ref_gdf(gdf, key):
    mydf = stub_df.GDF(gdf.cell, gdf.kpts)
    mydf._cderi = key
    mydf.cell = gdf.cell
    ref_stubs.GDF_REGISTRY[key] = gdf
    return mydf


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %s" % (name + ".npz", {k: np.asarray(v).shape for k, v in arrays.items()}))


def gen_eri():
    cases = {
        "eri_r_113": dict(kmesh=[1, 1, 3], nao=4, naux=10, neo=6, spin=1),
        "eri_r_222": dict(kmesh=[2, 2, 2], nao=6, naux=14, neo=7, spin=1),
        "eri_r_331": dict(kmesh=[3, 3, 1], nao=5, naux=9, neo=5, spin=1),
        "eri_u_122": dict(kmesh=[1, 2, 2], nao=5, naux=11, neo=6, spin=2),
    }
    for name, c in cases.items():
        gdf, C, basis = problem(c["kmesh"], c["nao"], c["naux"], c["neo"], spin=c["spin"])
        gdf.cell = GoldenCell(c["nao"])
        mydf = ref_gdf(gdf, name)
        out = dict(kmesh=np.array(c["kmesh"]), nao=c["nao"], naux=c["naux"], neo=c["neo"], spin=c["spin"],
                   gdf_seed=gdf.seed, gdf_scale=gdf.scale, C_ao_lo=C, basis=basis)
        out["s4_trs"] = ref_eri.get_emb_eri(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=4)
        out["s4_plain"] = ref_eri.get_emb_eri(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=4,
                                              t_reversal_symm=False)
        out["s1_trs"] = ref_eri.get_emb_eri(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=1)
        out["s4_chunk16"] = ref_eri.get_emb_eri(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=4, max_memory=0.02)
        if c["spin"] == 1:
            out["s8_trs"] = ref_eri.get_emb_eri(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=8)
        out["unit_s4"] = ref_eri.get_unit_eri(gdf.cell, mydf, C_ao_lo=C, symmetry=4)
        save(name, **out)


def gen_fourier_basis():
    rng = np.random.default_rng(42)
    kmesh = [2, 3, 2]
    nk = 12
    A_R = rng.standard_normal((nk, 4, 5))
    B_k = rng.standard_normal((2, nk, 3, 3)) + 1j * rng.standard_normal((2, nk, 3, 3))
    C = synthetic.make_C_ao_lo(kmesh, 6, 4, seed=9, spin=2)
    h = rng.standard_normal((nk, 6, 6)) + 1j * rng.standard_normal((nk, 6, 6))
    b = rng.standard_normal((nk, 4, 3))
    ref_log.verbose = "FATAL"     # k2R of a non-physical array warns about its imaginary part
    save("fourier_basis", kmesh=np.array(kmesh), A_R=A_R, B_k=B_k, C=C, h=h, b=b,
         R2k_A=ref_fourier.R2k(A_R, kmesh), k2R_B=ref_fourier.k2R(B_k, kmesh),
         roundtrip=ref_fourier.k2R(ref_fourier.R2k(A_R, kmesh), kmesh),
         h1_lo=ref_mb.transform_h1_to_lo(h, C), h1_lo_r=ref_mb.transform_h1_to_lo(h, C[0]),
         mult=ref_mb.multiply_basis(C, b))
    ref_log.verbose = "RESULT"


def make_ref_lattice(kmesh, nao, naux, nval, spin, sym, key):
    gdf, C, _ = problem(kmesh, nao, naux, 2, spin=spin)
    cell = GoldenCell(nao)
    gdf.cell = cell
    hcore, ovlp, vhf, rdm1 = mean_field(kmesh, nao, nval // 2 + 1, spin=spin)
    Lat = ref_lattice.Lattice(cell, kmesh)
    Lat.set_val_virt_core(nval, nao - nval, 0)
    mydf = ref_gdf(gdf, key)
    Lat.set_Ham(object(), mydf, C, eri_symmetry=sym, ovlp=ovlp, hcore=hcore, rdm1=rdm1, vhf=vhf, veff=vhf,
                vj=np.zeros_like(vhf), vk=np.zeros_like(vhf), H0=1.5)
    return Lat, gdf, C, (hcore, ovlp, vhf, rdm1)


def gen_embham():
    for name, (kmesh, nao, naux, nval, spin, sym) in {
            "embham_r_s4": ([1, 1, 3], 5, 11, 3, 1, 4), "embham_r_s1": ([1, 2, 2], 4, 9, 2, 1, 1),
            "embham_u_s4": ([1, 1, 3], 5, 11, 3, 2, 4)}.items():
        Lat, gdf, C, (hcore, ovlp, vhf, rdm1) = make_ref_lattice(kmesh, nao, naux, nval, spin, sym, name)
        rho = Lat.rdm1_lo_R * (0.5 if spin == 1 else 1.0)
        basis = ref_slater.get_emb_basis(Lat, rho)
        Ham, _ = ref_slater.get_emb_Ham(Lat, basis, None)
        # energy side: scaled DMET Hamiltonian and result transformation (slater.py:1780-1840, 1957-2032)
        Hd = ref_slater.get_H_dmet(basis, Lat, Ham, 0.0)
        Hd1 = ref_slater.get_H_dmet(basis, Lat, Ham, 0.0, compact=False) if sym == 1 else None
        rng = np.random.default_rng(5)
        nb = basis.shape[-1]
        rho_emb = rng.standard_normal((spin, nb, nb))
        rho_emb = rho_emb + rho_emb.transpose(0, 2, 1)
        rhoImp, Efrag, nelec = ref_slater.transformResults(rho_emb, -3.25, basis, Ham, lattice=Lat, last_dmu=0.1)
        extra = dict(Hd_H1=Hd.H1["cd"], Hd_H2=Hd.H2["ccdd"], Hd_H0=Hd.H0, rho_emb=rho_emb, rhoImp=rhoImp,
                     Efrag=Efrag, nelec=nelec)
        if Hd1 is not None:
            extra["Hd_H2_s1"] = Hd1.H2["ccdd"]
        save(name, **extra, kmesh=np.array(kmesh), nao=nao, naux=naux, nval=nval, spin=spin, sym=sym, gdf_seed=gdf.seed,
             gdf_scale=gdf.scale, C_ao_lo=C, hcore=hcore, ovlp=ovlp, vhf=vhf, rdm1=rdm1,
             hcore_lo_k=Lat.hcore_lo_k, rdm1_lo_k=Lat.rdm1_lo_k, rdm1_lo_R=Lat.rdm1_lo_R, vhf_lo_k=Lat.vhf_lo_k,
             rho=rho, basis=basis, H1=Ham.H1["cd"], H2=Ham.H2["ccdd"], ovlp_emb=Ham.ovlp, JK_core=Lat.JK_core,
             H0=Ham.H0, norb=Ham.norb)


def gen_lattice_misc():
    """Lattice.update_Ham (lattice.py:565-589) with a supplied HF potential, Lattice.transpose and expand_orb
    (353-397) through the reference's own Lattice"""
    out = {}
    for tag, (kmesh, nao, spin) in {"r": ([1, 2, 2], 4, 1), "u": ([2, 1, 3], 3, 2)}.items():
        Lat, gdf, C, (hcore, ovlp, vhf, rdm1) = make_ref_lattice(kmesh, nao, 7, 2, spin, 4, "misc_" + tag)
        rng = np.random.default_rng(17)
        new_R = Lat.rdm1_lo_R + 0.05 * rng.standard_normal(Lat.rdm1_lo_R.shape)       # "the DMET density matrix"
        new_vhf = vhf * 1.1
        out["kmesh_" + tag], out["nao_" + tag], out["spin_" + tag] = np.array(kmesh), nao, spin
        out["C_" + tag], out["hcore_" + tag], out["ovlp_" + tag] = C, hcore, ovlp
        out["vhf_" + tag], out["rdm1_" + tag] = vhf, rdm1
        out["new_R_" + tag], out["new_vhf_" + tag] = new_R, new_vhf
        out["transpose_" + tag] = Lat.transpose(new_R)
        out["expand_orb_" + tag] = Lat.expand_orb(new_R)
        ref_log.verbose = "FATAL"          # k2R of the perturbed matrices warns about imaginary parts
        Lat.update_Ham(new_R, vhf=new_vhf)
        ref_log.verbose = "RESULT"
        for k in ("rdm1_ao_k", "rdm1_lo_k", "rdm1_lo_R", "fock_lo_k", "fock_hf_lo_k", "vhf_lo_R", "veff_lo_k",
                  "hcore_lo_k"):
            out["%s_%s" % (k, tag)] = np.asarray(getattr(Lat, k))
    save("lattice_misc", **out)


def gen_emb_basis_hchain():
    """the reference's own fixture libdmet/routine/test/rdm1_lo (H chain 321g, 1x1x3, nval = nvirt = 2;
    test_slater.py:20-54) through the reference's get_emb_basis"""
    rdm1_lo = np.load(os.path.join(ref_stubs.REF_ROOT, "libdmet", "routine", "test", "rdm1_lo"))
    cell = GoldenCell(4)
    Lat = ref_lattice.Lattice(cell, [1, 1, 3])
    Lat.set_val_virt_core(2, 2, 0)
    basis = ref_slater.get_emb_basis(Lat, rdm1_lo)
    Lat.val_idx, Lat.virt_idx = list(range(4)), []
    basis_trunc = ref_slater.get_emb_basis(Lat, rdm1_lo, nbath=2, valence_bath=False)
    save("emb_basis_hchain", rdm1_lo=rdm1_lo, basis=basis, basis_trunc=basis_trunc)


def gen_emb_basis_eig():
    """eigenvalue construction of the bath (slater.py:224-318) on the mean-field density matrices of the embHam
    cases and on the reference's H-chain fixture"""
    out = {}
    for tag, (kmesh, nao, naux, nval, spin) in {"r": ([1, 1, 3], 5, 11, 3, 1), "u": ([1, 2, 2], 4, 9, 2, 2)}.items():
        Lat, gdf, C, _ = make_ref_lattice(kmesh, nao, naux, nval, spin, 4, "eig_" + tag)
        rho = Lat.rdm1_lo_R * (0.5 if spin == 1 else 1.0)
        out["rho_" + tag] = rho
        out["kmesh_" + tag], out["nao_" + tag], out["nval_" + tag] = np.array(kmesh), nao, nval
        out["basis_" + tag] = ref_slater.get_emb_basis(Lat, rho, kind='eig')
        out["basis_full_" + tag] = ref_slater.get_emb_basis(Lat, rho, kind='eig', valence_bath=False)
    rdm1_lo = np.load(os.path.join(ref_stubs.REF_ROOT, "libdmet", "routine", "test", "rdm1_lo"))
    Lat = ref_lattice.Lattice(GoldenCell(4), [1, 1, 3])
    Lat.set_val_virt_core(2, 2, 0)
    out["basis_hchain"] = ref_slater.get_emb_basis(Lat, rdm1_lo, kind='eig')
    save("emb_basis_eig", **out)


def gen_gso_basis():
    """GSO bath construction through the reference's libdmet.routine.spinless.get_emb_basis (spinless.py:34-272) on
    the generalised mean-field density matrix of the GSO test lattice"""
    import libdmet.routine.spinless as ref_spinless
    import libdmet.system.fourier as ref_f
    from helpers import GSOLattice
    out = {}
    for tag, (kmesh, nao, nval) in {"a": ([1, 1, 3], 3, 2), "b": ([2, 2, 1], 4, 4)}.items():
        gdf, C, _ = problem(kmesh, nao, 6, 2, spin=2)
        gdf.cell = GoldenCell(nao)
        G = GSOLattice(gdf, C, ref_f)
        GRho = ref_f.k2R(G.rdm1_lo_k, kmesh)
        Lat = ref_lattice.Lattice(gdf.cell, kmesh)
        Lat.set_val_virt_core(nval, nao - nval, 0)
        out["kmesh_" + tag], out["nao_" + tag], out["nval_" + tag], out["GRho_" + tag] = np.array(kmesh), nao, nval, GRho
        for kind in ("svd", "eig"):
            out["%s_%s" % (kind, tag)] = ref_spinless.get_emb_basis(Lat, GRho, kind=kind)
            out["%s_full_%s" % (kind, tag)] = ref_spinless.get_emb_basis(Lat, GRho, kind=kind, valence_bath=False)
    save("gso_basis", **out)


def gen_rho_glob():
    """global density matrix by democratic partitioning through the reference's slater_helper.get_rho_glob_R
    (slater_helper.py:183-270): one fragment (restricted and unrestricted) and two fragments sharing cell 0"""
    import libdmet.routine.slater_helper as ref_helper
    out = {}
    rng = np.random.default_rng(123)
    for tag, (kmesh, nlo, neo, spin, imp) in {"r": ([1, 2, 2], 5, 7, 1, [0, 1, 2]), "u": ([2, 1, 3], 4, 6, 2, [0, 1, 2, 3])}.items():
        Lat = ref_lattice.Lattice(GoldenCell(nlo), kmesh)
        Lat.set_val_virt_core(len(imp), 0, 0)
        Lat.val_idx, Lat.virt_idx = list(imp), []
        nk = int(np.prod(kmesh))
        basis = rng.standard_normal((spin, nk, nlo, neo))
        rho = rng.standard_normal((spin, neo, neo))
        rho = rho + rho.transpose(0, 2, 1)
        assert list(Lat.imp_idx) == list(imp)
        out["kmesh_" + tag], out["imp_" + tag] = np.array(kmesh), np.array(imp)
        out["basis_" + tag], out["rho_" + tag] = basis, rho
        out["glob_" + tag] = ref_helper.get_rho_glob_R(basis, Lat, rho)
    # two fragments: orbitals {0, 1} and {2, 3, 4} of a 5-orbital cell, different numbers of embedding orbitals,
    # a non-symmetric "density matrix" for the second one (the reference does not symmetrise)
    kmesh, nlo = [1, 1, 3], 5
    lats, bases, rhos = [], [], []
    for imp, neo in (([0, 1], 4), ([2, 3, 4], 6)):
        Lat = ref_lattice.Lattice(GoldenCell(nlo), kmesh)
        Lat.set_val_virt_core(len(imp), 0, 0)
        Lat.val_idx, Lat.virt_idx = list(imp), []
        lats.append(Lat)
        bases.append(rng.standard_normal((1, 3, nlo, neo)))
        rhos.append(rng.standard_normal((1, neo, neo)))
    out["basis_f0"], out["basis_f1"], out["rho_f0"], out["rho_f1"] = bases[0], bases[1], rhos[0], rhos[1]
    out["glob_f"] = ref_helper.get_rho_glob_R(bases, lats, rhos)
    save("rho_glob", **out)


def gen_gso():
    """GSO embedding ERI (eri_transform.py:1104-1284); the reference itself imports
    libdmet.routine.spinless.separate_basis inside the function"""
    for name, (kmesh, nao, naux, nemb, cspin) in {"gso_113": ([1, 1, 3], 4, 9, 6, 1), "gso_122": ([1, 2, 2], 3, 8, 5, 2)}.items():
        gdf, C, _ = problem(kmesh, nao, naux, 2, spin=cspin)
        gdf.cell = GoldenCell(nao)
        mydf = ref_gdf(gdf, name)
        nk = int(np.prod(kmesh))
        rng = np.random.default_rng(77)
        basis, _ = np.linalg.qr(rng.standard_normal((nk * 2 * nao, nemb)))
        basis = basis.reshape(nk, 2 * nao, nemb)
        out = dict(kmesh=np.array(kmesh), nao=nao, naux=naux, nemb=nemb, gdf_seed=gdf.seed, gdf_scale=gdf.scale,
                   C_ao_lo=C, basis=basis)
        out["s4_trs"] = ref_eri.get_emb_eri_gso(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=4)
        out["s4_plain"] = ref_eri.get_emb_eri_gso(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=4,
                                                  t_reversal_symm=False)
        out["s1_trs"] = ref_eri.get_emb_eri_gso(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=1)
        out["unit_s4"] = ref_eri.get_emb_eri_gso(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=4, unit_eri=True)
        save(name, **out)


def gen_gso_embham():
    """GSO embedding Hamiltonian through the reference's libdmet.routine.spinless.get_emb_Ham (spinless.py:433-726)"""
    import libdmet.routine.spinless as ref_spinless
    import libdmet.system.fourier as ref_f
    from helpers import GSOLattice, gso_basis

    class Vcor(object):
        def islocal(self):
            return True

    for name, (kmesh, nao, naux, nemb, sym) in {"gso_embham_113": ([1, 1, 3], 3, 8, 5, 4),
                                                 "gso_embham_221": ([2, 2, 1], 3, 7, 6, 1)}.items():
        gdf, C, _ = problem(kmesh, nao, naux, 2, spin=2)
        gdf.cell = GoldenCell(nao)
        mydf = ref_gdf(gdf, name)
        Lat = GSOLattice(gdf, C, ref_f, eri_symmetry=sym)
        Lat.df = mydf
        basis = gso_basis(kmesh, nao, nemb)
        Ham, _ = ref_spinless.get_emb_Ham(Lat, basis, Vcor(), 0.3)
        # energy side (spinless.py:754-848, 948-1035)
        # (sic) with a 4-fold H2 the reference scales ImpHam.H2["ccdd"] IN PLACE here -- `ao2mo.restore(4, ..)` hands
        # back its input when nothing is to be converted (l.1026-1027) -- so every call gets a fresh copy
        H2_clean = np.array(Ham.H2["ccdd"])
        Hd = ref_spinless.get_H_dmet(basis, Lat, Ham, last_dmu=0.1, mu=0.3)
        Ham.H2["ccdd"] = H2_clean.copy()
        Hd1 = ref_spinless.get_H_dmet(basis, Lat, Ham, last_dmu=0.1, mu=0.3, compact=False)
        Ham.H2["ccdd"] = H2_clean.copy()
        rng = np.random.default_rng(11)
        GRhoEmb = rng.standard_normal((nemb, nemb))
        GRhoEmb = GRhoEmb + GRhoEmb.T
        GRhoImp, Efrag, nelec = ref_spinless.transformResults(GRhoEmb, -2.75, Lat, basis, Ham, None, 0.3,
                                                              last_dmu=0.1)
        save(name, kmesh=np.array(kmesh), nao=nao, naux=naux, nemb=nemb, sym=sym, gdf_seed=gdf.seed,
             gdf_scale=gdf.scale, C_ao_lo=C, basis=basis, mu=0.3, H1=Ham.H1["cd"], H2=Ham.H2["ccdd"],
             ovlp=Ham.ovlp, H0=Ham.H0, JK_core=Lat.JK_core, Hd_H1=Hd.H1["cd"], Hd_H2=Hd.H2["ccdd"], Hd_H0=Hd.H0,
             Hd_H2_s1=Hd1.H2["ccdd"], GRhoEmb=GRhoEmb, GRhoImp=GRhoImp, Efrag=Efrag, nelec=nelec)


def gen_gdf_lo():
    """GDF tensor rotated to the LO basis through the reference's transform_gdf_to_lo (eri_transform.py:1312-1427);
    the h5py stub collects what the reference writes to its output file"""
    for name, (kmesh, nao, nlo, naux) in {"gdf_lo_113": ([1, 1, 3], 4, 4, 6), "gdf_lo_221": ([2, 2, 1], 5, 3, 7)}.items():
        gdf, _, _ = problem(kmesh, nao, naux, 2)
        gdf.cell = GoldenCell(nao)
        mydf = ref_gdf(gdf, name)
        C = synthetic.make_C_ao_lo(kmesh, nao, nlo, seed=31)
        out = dict(kmesh=np.array(kmesh), nao=nao, nlo=nlo, naux=naux, gdf_seed=gdf.seed, gdf_scale=gdf.scale, C_ao_lo=C)
        for tag, trs in (("trs", True), ("plain", False)):
            ref_eri.transform_gdf_to_lo(mydf, C, fname=name + tag, t_reversal_symm=trs)
            written = ref_stubs.H5_WRITTEN[name + tag]
            npair = len(written["j3c-kptij"])
            assert sorted(k for k in written if k != "j3c-kptij") == sorted("j3c/%d/0" % k for k in range(npair))
            for k in range(npair):
                out["%s_%d" % (tag, k)] = written["j3c/%d/0" % k]
        out["kptij"] = written["j3c-kptij"]
        save(name, **out)


FILE_CASES = {
    # name: kmesh, nao, naux, neo, spin, cderi layout options (tests/helpers.py: write_case_file)
    "eri_file_122": dict(kmesh=[1, 2, 2], nao=4, naux=9, neo=5, spin=1, nsegments=2, pack_diagonal=True,
                         drop={(2, 1): 8, (3, 3): 7}),
    "eri_file_113": dict(kmesh=[1, 1, 3], nao=5, naux=8, neo=6, spin=2, nsegments=1, pack_diagonal=True, drop={}),
}


def gen_eri_file():
    """the reference's get_naoaux / sr_loop / get_emb_eri / transform_gdf_to_lo over a REAL cderi file (PySCF's v1
    layout: `j3c-kptij`, `j3c/<pair>/<segment>`, Hermitian-packed diagonal pairs, real Gamma block, fewer auxiliary
    rows for some pairs) written by libdmet_preview_b200.gdf_file.write_gdf_file and read through h5lite, which
    stands in for h5py"""
    import tempfile
    from libdmet_preview_b200 import h5lite
    from libdmet_preview_b200.gdf_file import write_gdf_file
    for name, c in FILE_CASES.items():
        gdf, C, basis = problem(c["kmesh"], c["nao"], c["naux"], c["neo"], spin=c["spin"])
        gdf.cell = GoldenCell(c["nao"])
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, name + "_cderi.h5")
            write_gdf_file(path, gdf, version="v1", nsegments=c["nsegments"], pack_diagonal=c["pack_diagonal"],
                           naux_of=c["drop"])
            ref_stubs.FILE_NAO[path] = c["nao"]
            mydf = stub_df.GDF(gdf.cell, gdf.kpts)
            mydf._cderi = path
            out = dict(kmesh=np.array(c["kmesh"]), nao=c["nao"], naux=c["naux"], neo=c["neo"], spin=c["spin"],
                       gdf_seed=gdf.seed, gdf_scale=gdf.scale, C_ao_lo=C, basis=basis)
            ref_log.verbose = "FATAL"            # "aux basis drop may happened" (eri_transform.py:191-192)
            out["naoaux"] = ref_eri.get_naoaux(mydf)
            ref_log.verbose = "RESULT"
            out["s4_trs"] = ref_eri.get_emb_eri(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=4)
            out["s4_plain"] = ref_eri.get_emb_eri(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=4,
                                                  t_reversal_symm=False)
            out["s1_trs"] = ref_eri.get_emb_eri(gdf.cell, mydf, C_ao_lo=C, basis=basis, symmetry=1)
            # LO-basis tensor written by the reference into a real file, read back dataset by dataset
            lo_path = os.path.join(tmp, name + "_lo.h5")
            Clo = synthetic.make_C_ao_lo(c["kmesh"], c["nao"], c["nao"] - 1, seed=41)
            mydf_lo = ref_eri.transform_gdf_to_lo(mydf, Clo, fname=lo_path)
            assert mydf_lo._cderi == lo_path
            out["C_lo"] = Clo
            with h5lite.File(lo_path) as f:
                out["lo_kptij"] = f["j3c-kptij"][...]
                for k in range(len(out["lo_kptij"])):
                    assert f["j3c/%d" % k].keys() == ["0"]
                    out["lo_%d" % k] = f["j3c/%d/0" % k][...]
        save(name, **out)


if __name__ == "__main__":
    gen_lattice_misc()
    gen_gso_basis()
    gen_rho_glob()
    gen_emb_basis_eig()
    gen_eri_file()
    gen_gdf_lo()
    gen_gso_embham()
    gen_gso()
    gen_eri()
    gen_fourier_basis()
    gen_embham()
    gen_emb_basis_hchain()
