"""HF-in-DMET on a seeded synthetic cell: what libdmet/test/test_mfd.py does with a PySCF KRHF object, without PySCF.

The k-point mean field has to be SELF-CONSISTENT with the very GDF tensor the embedding ERI is built from, otherwise
neither identity the reference asserts holds (folded density matrix = fixed point of the impurity HF,
test_mfd.py:138; fragment energy = k-point HF energy per cell, test_mfd.py:153).  `lattice_scf` therefore runs a
Hartree-Fock on the whole lattice with the two-electron integrals of the supercell obtained from the ORACLE's
`get_emb_eri` with the identity basis (its "k2gamma" mode, eri_transform.py:272-282) -- translation symmetry makes
that the k-point HF with exxdiv=None that the reference test feeds to `Lattice.set_Ham`.

`dmet_cycle` is the loop body: lattice HF (`mfd.HF`) -> bath (`get_emb_basis`) -> embedding Hamiltonian (`embHam`) ->
impurity HF (oracle/scf.py) -> `get_H_dmet` -> fragment energy.  It takes the module set to use, so that the same
code runs the product (CUDA) path and the oracle path.
"""
import numpy as np

from libdmet_preview_b200 import synthetic
from oracle import eri_transform as o_eri, fourier as o_fourier, slater as o_slater, mfd as o_mfd, scf as o_scf
from oracle import pyscf_lib as olib


def gapped_hcore(kmesh, nao, nocc, seed=0, gap=3.0, hop=0.35):
    """Hermitian hcore[k] with time-reversal structure and a gap at the Fermi level of BOTH spin channels: with
    nocc = (n_a, n_b) occupied bands per k-point, n_b levels sit near -gap, n_a - n_b near 0 and the rest near +gap,
    coupled by a random Hermitian matrix -- a (magnetic) band insulator, so SCF iterations converge to machine
    precision and filling the lowest levels per k-point equals filling them over the whole zone"""
    na, nb = (max(nocc), min(nocc)) if np.ndim(nocc) else (nocc, nocc)
    h = synthetic.make_hermitian_k(kmesh, nao, seed=900 + seed, scale=hop)
    level = np.diag([-gap] * nb + [0.0] * (na - nb) + [gap] * (nao - na)).astype(np.complex128)
    return h + level[None]


def lattice_scf(gdf, hcore_k, nocc, tol=1e-13, max_iter=200):
    """k-point HF (orthonormal AOs) on the synthetic GDF tensor.  nocc: occupied bands per k-point -- an int
    (restricted) or a pair (unrestricted).  Returns dict(vhf, rdm1 (PySCF convention: spin-traced for RHF, per spin
    for UHF), fock, e_cell, spin)."""
    kmesh, nao = gdf.kmesh, gdf.nao
    nk = len(gdf.kpts_scaled)
    lat = o_fourier.StripeLattice(kmesh, nao)
    eri = o_eri.get_emb_eri(gdf.cell, gdf)                       # (1, npair_s, npair_s), orbital index (R, p)
    spin = 2 if np.ndim(nocc) else 1
    occ = list(nocc) if spin == 2 else [int(nocc)]

    def veff_k(rho_k):
        """rho_k (spin, nk, nao, nao) per spin channel -> veff (spin, nk, nao, nao)"""
        rho_R = np.asarray([o_fourier.FFTtoT(rho_k[s], kmesh) for s in range(spin)])
        dm = lat.expand(rho_R, dense=True)
        v = o_slater._get_veff(dm * (2.0 if spin == 1 else 1.0), eri)
        v = np.asarray(v).reshape(spin, nk * nao, nk * nao)
        return np.asarray([o_fourier.FFTtoK(lat.extract_stripe(v[s]), kmesh) for s in range(spin)])

    def density(fock):
        """lowest occ[s] * nk levels of the whole zone per spin channel (what mfd.HF's assignocc fills)"""
        rho = np.zeros((spin, nk, nao, nao), dtype=np.complex128)
        for s in range(spin):
            ec = [np.linalg.eigh(fock[s, k]) for k in range(nk)]
            levels = np.sort(np.concatenate([e for e, _ in ec]))
            nel = occ[s] * nk
            if nel < len(levels) and levels[nel] - levels[nel - 1] < 1e-6:
                raise RuntimeError("synthetic mean field is not gapped: choose other parameters")
            mu = 0.5 * (levels[nel - 1] + levels[min(nel, len(levels) - 1)])
            for k in range(nk):
                e, c = ec[k]
                co = c[:, e < mu]
                rho[s, k] = co.dot(co.conj().T)
        return rho

    rho = density(np.asarray([hcore_k] * spin))
    e_old, focks, errs = None, [], []
    per = 2.0 if spin == 1 else 1.0
    for it in range(max_iter):
        v = veff_k(rho)
        fock = hcore_k[None] + v
        e = per * sum(np.einsum("kpq,kqp->", hcore_k + 0.5 * v[s], rho[s]).real for s in range(spin)) / nk
        comm = np.einsum("skpq,skqr->skpr", fock, rho) - np.einsum("skpq,skqr->skpr", rho, fock)
        err = np.abs(comm).max()
        if e_old is not None and abs(e - e_old) < tol and err < 1e-12:
            break
        e_old = e
        focks.append(fock)
        errs.append(comm.ravel())
        focks, errs = focks[-10:], errs[-10:]
        if len(errs) > 1:                                    # DIIS on the commutator [F, rho]
            m = len(errs)
            B = -np.ones((m + 1, m + 1))
            B[m, m] = 0.0
            for i in range(m):
                for j in range(m):
                    B[i, j] = np.vdot(errs[i], errs[j]).real
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            c = np.linalg.lstsq(B, rhs, rcond=1e-15)[0][:m]
            fock = sum(ci * Fi for ci, Fi in zip(c, focks))
        rho = density(fock)
    else:
        raise RuntimeError("lattice SCF did not converge (err %.2e)" % err)
    v = veff_k(rho)
    per = 2.0 if spin == 1 else 1.0
    e = per * sum(np.einsum("kpq,kqp->", hcore_k + 0.5 * v[s], rho[s]).real for s in range(spin)) / nk
    if spin == 1:
        return dict(vhf=v[0], rdm1=rho[0] * 2.0, fock=hcore_k + v[0], e_cell=float(e), spin=1, iterations=it)
    return dict(vhf=v, rdm1=rho, fock=hcore_k[None] + v, e_cell=float(e), spin=2, iterations=it)


class OracleMods(object):
    """the oracle's routines under the names `dmet_cycle` uses"""
    get_emb_basis = staticmethod(o_slater.get_emb_basis)
    embHam = staticmethod(o_slater.embHam)
    get_H_dmet = staticmethod(o_slater.get_H_dmet)

    @staticmethod
    def foldRho_k(rho_k, basis_k):
        return o_slater.transform_h1(rho_k, basis_k)


class ProductMods(object):
    """the CUDA path (libdmet_preview_b200.slater)"""

    def __init__(self):
        from libdmet_preview_b200 import slater
        self.get_emb_basis, self.embHam, self.get_H_dmet = slater.get_emb_basis, slater.embHam, slater.get_H_dmet
        self.foldRho_k = slater.transform_h1


def fragment_energy(Hd, rdm1_emb):
    """E = H0 + sum h1 . gamma + 1/2 sum (pq|rs) Gamma_pqrs with the Hartree-Fock two-body density matrix of
    `rdm1_emb` (spin, n, n; per spin channel), evaluated through J and K of the SCALED integrals -- what
    test_mfd.py:143-151 computes with explicit rdm2 einsums"""
    H1, H2 = np.asarray(Hd.H1["cd"]), np.asarray(Hd.H2["ccdd"])
    spin = rdm1_emb.shape[0]
    if spin == 1:
        dm = rdm1_emb * 2.0
        veff = o_slater._get_veff(dm, H2)
        return float(Hd.H0 + np.sum(H1[0] * dm[0]) + 0.5 * np.sum(veff[0] * dm[0]))
    veff = o_slater._get_veff(rdm1_emb, H2)
    return float(Hd.H0 + sum(np.sum((H1[s] + 0.5 * veff[s]) * rdm1_emb[s]) for s in range(2)))


def dmet_cycle(Lat, mods, filling, restricted, nelec_emb):
    """One HF-in-DMET pass over a lattice that already carries its Hamiltonian.  Returns a dict with the quantities
    libdmet/test/test_mfd.py asserts on."""
    out = {}
    rhoT, mu, E = o_mfd.HF(Lat, None, filling, restricted, mu0=0.0, beta=np.inf)
    out["E_lattice_HF"] = E
    out["rhoT"] = rhoT
    ref = np.asarray(Lat.rdm1_lo_R) * (0.5 if restricted else 1.0)
    out["rdm_diff"] = float(np.abs(rhoT - ref).max())                       # test_mfd.py:107-113
    basis = mods.get_emb_basis(Lat, rhoT)
    basis_k = Lat.R2k_basis(basis)
    ImpHam, _ = mods.embHam(Lat, basis, None)
    rdm1_fold = np.asarray(mods.foldRho_k(np.asarray(Lat.rdm1_lo_k), basis_k)).real
    rdm1_fold = rdm1_fold * (0.5 if restricted else 1.0)
    E_imp, rdm1_emb = o_scf.hf_energy(ImpHam.H0, ImpHam.H1["cd"], ImpHam.H2["ccdd"], nelec_emb, tol=1e-14,
                                      dm0=rdm1_fold)
    out["fixed_point_diff"] = float(np.abs(rdm1_emb - rdm1_fold).max())     # test_mfd.py:131-138
    Hd = mods.get_H_dmet(basis, Lat, ImpHam, 0.0, compact=True)
    out["E_frag"] = fragment_energy(Hd, rdm1_emb)                            # test_mfd.py:143-153
    out["E_imp"] = E_imp
    out["basis"], out["ImpHam"], out["rdm1_emb"] = basis, ImpHam, rdm1_emb
    return out
