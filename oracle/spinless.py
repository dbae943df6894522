"""Oracle: generalised-spin-orbital embedding Hamiltonian, ab-initio interacting-bath Hartree-Fock branch.

Behaviour restated (own numpy code) from libdmet/routine/spinless.py:433-462 (get_emb_Ham), 464-558 (two-body part:
get_emb_eri_gso), 560-726 (one-body part) and libdmet/routine/spinless_helper.py:31-46, 349-438.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np

from . import eri_transform
from .slater import Integral, transform_trans_inv_k as fold_generalised, _get_jk


def separate_basis(basis):
    half = basis.shape[1] // 2
    return basis[:, :half], basis[:, half:]


def transform_trans_inv_k(basis_ka, basis_kb, H_k):
    """Re sum_k [Ba^+ H_aa Ba + Bb^+ H_bb Bb (+ Ba^+ H_ab Bb + h.c.)] / nkpts   (spinless_helper.py:349-381)"""
    H_k = np.asarray(H_k)
    assert H_k.ndim == 4 and H_k.shape[0] in (2, 3)
    tot = np.einsum("kpm,kpq,kqn->mn", basis_ka.conj(), H_k[0], basis_ka)
    tot = tot + np.einsum("kpm,kpq,kqn->mn", basis_kb.conj(), H_k[1], basis_kb)
    if H_k.shape[0] == 3:
        ab = np.einsum("kpm,kpq,kqn->mn", basis_ka.conj(), H_k[2], basis_kb)
        tot = tot + ab + ab.conj().T
    return tot.real / float(len(basis_ka))


def transform_local(basis_Ra, basis_Rb, H):
    """spinless_helper.py:383-409"""
    H = np.asarray(H)
    res = np.einsum("Rpm,pq,Rqn->mn", basis_Ra.conj(), H[0], basis_Ra)
    res = res + np.einsum("Rpm,pq,Rqn->mn", basis_Rb.conj(), H[1], basis_Rb)
    if H.shape[0] == 3:
        ab = np.einsum("Rpm,pq,Rqn->mn", basis_Ra.conj(), H[2], basis_Rb)
        res = res + ab + ab.conj().T
    return res


def transform_imp(basis_Ra, basis_Rb, H):
    """spinless_helper.py:411-438"""
    return transform_local(basis_Ra[:1], basis_Rb[:1], H)


def get_emb_Ham(lattice, basis, vcor, mu, local=True, int_bath=True, add_vcor=False, **kwargs):
    """spinless.py:433-726, interacting bath + Hartree-Fock"""
    assert int_bath and not lattice.is_model
    basis = np.asarray(basis)
    nao, nb = lattice.nscsites, basis.shape[-1]
    H2 = kwargs.get("H2_given", None)
    if H2 is None:
        H2 = eri_transform.get_emb_eri_gso(lattice.cell, lattice.df, C_ao_lo=lattice.C_ao_lo, basis=basis,
                                           kscaled_center=kwargs.get("kscaled_center", None),
                                           symmetry=lattice.eri_symmetry,
                                           t_reversal_symm=kwargs.get("t_reversal_symm", True))
    basis_k = lattice.R2k_basis(basis)
    Ra, Rb = separate_basis(basis)
    ka, kb = separate_basis(basis_k)
    hcore_k = kwargs.get("hcore_custom", None)
    hcore_emb = transform_trans_inv_k(ka, kb, lattice.getH1(kspace=True) if hcore_k is None else hcore_k)
    ovlp_emb = transform_trans_inv_k(ka, kb, lattice.get_ovlp(kspace=True))
    rdm1_emb = fold_generalised(basis_k, lattice.rdm1_lo_k)
    H1 = transform_trans_inv_k(ka, kb, lattice.fock_hf_lo_k)
    if kwargs.get("hcore_add", None) is not None:
        H1 = H1 + transform_imp(Ra, Rb, kwargs["hcore_add"])
    vj, vk = _get_jk(rdm1_emb, H2)
    H1 = H1 - (vj[0] - vk[0])
    lattice.JK_core = H1 - hcore_emb
    mu_mat = np.zeros((2, nao, nao))
    np.fill_diagonal(mu_mat[0], -mu)
    np.fill_diagonal(mu_mat[1], mu)
    H1 = H1 + transform_local(Ra, Rb, mu_mat)
    if add_vcor:
        H1 = H1 + transform_local(Ra, Rb, vcor.get())
        if not kwargs.get("fitting", False):
            H1 = H1 - transform_imp(Ra, Rb, vcor.get())
        if lattice.get_JK_imp() is not None:
            H1 = H1 - transform_imp(Ra, Rb, lattice.get_JK_imp())
    H0 = lattice.getH0() + kwargs.get("H0_add", 0.0)
    return Integral(nb, True, False, H0, {"cd": np.asarray(H1)[None]}, {"ccdd": H2}, ovlp=ovlp_emb), None


embHam = get_emb_Ham


# ---------------------------------------------------------------------------------------------------------
# GSO bath construction (spinless.py:34-272)
# ---------------------------------------------------------------------------------------------------------
def get_emb_basis(lattice, GRho, kind='svd', valence_bath=True, tol_bath=1e-9, nbath=None):
    """embedding basis (ncells, 2 nlo, nimp + nbath) from the generalised density matrix (ncells, 2 nlo, 2 nlo):
    bath = left singular vectors of the environment x impurity block ('svd', l.58-162) or the fractionally occupied
    eigenvectors of the environment block ('eig', l.167-272); impurity rows of the bath zeroed and the bath Loewdin
    re-orthonormalised; bath columns ordered by decreasing weight on the alpha rows."""
    import scipy.linalg as la
    from .slater import vec_lowdin
    ncells, nlo = lattice.ncells, lattice.nscsites
    nso = 2 * nlo
    two_spins = lambda orbs: np.concatenate([np.asarray(orbs, dtype=int), np.asarray(orbs, dtype=int) + nlo])  # noqa
    imp_idx = two_spins(lattice.imp_idx)
    generators = two_spins(lattice.val_idx) if valence_bath else imp_idx
    every = np.arange(ncells * nso)                      # spin orbital R * nso + s * nlo + i, in increasing order
    env_idx = every[~np.isin(every, generators)]
    virt_mask = np.isin(env_idx, imp_idx)                # impurity (virtual) spin orbitals left in the environment
    alpha_mask = (env_idx % nso) < nlo                   # s == 0 half of every cell
    rdm1 = np.asarray(GRho).real
    if kind == 'svd':
        u, sigma, _ = la.svd(rdm1.reshape(ncells * nso, nso)[env_idx][:, generators], full_matrices=False)
        nbath = int((sigma >= tol_bath).sum()) if nbath is None else nbath
        B = u[:, :nbath]
    else:
        ew, ev = la.eigh(lattice.expand(rdm1)[env_idx][:, env_idx])
        B = np.asarray([ev[:, i] for i, e in enumerate(ew) if abs(e) > tol_bath and abs(1 - e) > tol_bath]).T
        nbath = B.shape[-1]
    assert nbath % 2 == 0
    B[virt_mask] = 0.0
    B = vec_lowdin(B)
    w = np.einsum("ai,ai->i", B[alpha_mask], B[alpha_mask])
    order = np.argsort(w, kind='mergesort')[::-1]
    nimp = len(imp_idx)
    basis = np.zeros((ncells * nso, nimp + nbath))
    basis[imp_idx, :nimp] = np.eye(nimp)
    basis[env_idx, nimp:] = B[:, order]
    return basis.reshape(ncells, nso, nimp + nbath)


# ---------------------------------------------------------------------------------------------------------
# energy side of the GSO iteration (spinless.py:754-848, 948-1035)
# ---------------------------------------------------------------------------------------------------------
def idx_ao2so(idx, nao):
    """spinless_helper.py:247-259"""
    return [i for i in idx], [i + nao for i in idx]


def transformResults(GRhoEmb, E, lattice, basis, ImpHam, H1e, mu, **kwargs):
    """(GRhoImp, Efrag, nelec) of one GSO impurity problem (spinless.py:754-848, fit_ghf=False)"""
    from .slater import get_H1_scaled
    ncells, nso, nbasis = basis.shape
    nao = nso // 2
    ia, ib = idx_ao2so(lattice.imp_idx, nao)
    GRhoEmb = np.asarray(GRhoEmb)
    if GRhoEmb.ndim == 3:
        GRhoEmb = GRhoEmb[0] if GRhoEmb.shape[0] == 1 else GRhoEmb.sum(axis=0)
    GRhoImp = basis[0].dot(GRhoEmb).dot(basis[0].conj().T)
    nelec = GRhoImp[ia, ia].sum() - GRhoImp[ib, ib].sum() + len(ib)
    if E is None:
        return GRhoImp, None, nelec
    Ra, Rb = separate_basis(basis)
    H1 = np.asarray(ImpHam.H1["cd"][0])
    E2 = E - np.sum(H1 * GRhoEmb.T) - ImpHam.H0                       # two-body part of the solver energy
    where = kwargs.get("dmu_idx", None)
    where = list(lattice.imp_idx) if where is None else list(where)
    ea, eb = idx_ao2so(kwargs.get("imp_idx", np.arange(lattice.nimp)), lattice.nimp)
    # chemical potentials go back in: last_dmu on the chosen cell-0 orbitals only, mu on every orbital of every cell;
    # opposite signs for the two spin flavours (the reference reuses one buffer and overwrites its whole diagonal)
    local = np.zeros((2, nao, nao))
    local[0][where, where] = kwargs["last_dmu"]
    local[1][where, where] = -kwargs["last_dmu"]
    everywhere = np.asarray([np.eye(nao) * mu, np.eye(nao) * (-mu)])
    H1_eff = H1 + transform_imp(Ra, Rb, local) + transform_local(Ra, Rb, everywhere)
    if lattice.JK_core is not None:
        H1_eff = H1_eff - 0.5 * lattice.JK_core
    H1_eff = get_H1_scaled(np.array(H1_eff)[None], list(ea) + list(eb))[0]
    return GRhoImp, np.sum(H1_eff * GRhoEmb.T) + E2 + ImpHam.H0, nelec


def get_H_dmet(basis, lattice, ImpHam, last_dmu=None, mu=None, imp_idx=None, compact=True, **kwargs):
    """scaled GSO DMET Hamiltonian (spinless.py:948-1035; default branch: E1 from the lattice, JK_core of the last
    embHam, no vcor / GV terms)"""
    from .slater import get_H1_scaled, get_H2_scaled
    from . import pyscf_lib as lib
    nbasis = basis.shape[-1]
    basis_k = lattice.R2k_basis(basis)
    ka, kb = separate_basis(basis_k)
    ea, eb = idx_ao2so(np.arange(lattice.nimp) if imp_idx is None else imp_idx, lattice.nimp)
    imp = list(ea) + list(eb)
    H1 = transform_trans_inv_k(ka, kb, lattice.getH1(kspace=True))
    H1 = H1 + 0.5 * (lattice.JK_core if lattice.JK_core is not None else 0.0)
    H1 = get_H1_scaled(np.asarray(H1)[None], imp)
    H2 = lib.restore(4, np.asarray(ImpHam.H2["ccdd"][0]), nbasis)
    H2 = get_H2_scaled(np.array(H2)[None], imp)
    if not compact:
        H2 = lib.restore(1, H2[0], nbasis)[None]
    return Integral(nbasis, True, False, lattice.getH0(), {"cd": H1}, {"ccdd": H2})
