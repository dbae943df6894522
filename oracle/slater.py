"""Oracle: embedding basis and embedding Hamiltonian (Slater-determinant DMET).

Restates libdmet/routine/slater.py:98-220 (get_emb_basis, SVD), 320-370 (get_emb_Ham), 438-476 (ab-initio branch of
__embHam2e), 478-523 (get_veff, HF branch), 525-547 + 559-560 + 590-605 + 639-643 (__embHam1e, interacting bath,
HF), 690-712; libdmet/routine/slater_helper.py:37-50,73-80,102-103,494-517; libdmet/solver/scf.py:255-352;
libdmet/lo/lowdin.py:83-136; libdmet/system/integral.py:61-105,883-928.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`lattice` is duck-typed: .ncells .nscsites .imp_idx .val_idx .expand() .R2k_basis() .df .cell .C_ao_lo
.eri_symmetry .is_model .hcore_lo_k .vhf_lo_k .ovlp_lo_k .fock_lo_k .rdm1_lo_k .rdm1_lo_R .getH0() .JK_core
"""
import numpy as np
import scipy.linalg as la

from . import pyscf_lib as lib
from .fourier import max_abs, IMAG_DISCARD_TOL
from .make_basis import mdot, add_spin_dim
from . import eri_transform


# ---------------------------------------------------------------------------------------------------------
# libdmet/lo/lowdin.py:83-136
# ---------------------------------------------------------------------------------------------------------
def _lowdin(s, tol=1e-14):
    e, v = la.eigh(s)
    idx = e > tol
    return np.dot(v[:, idx] / np.sqrt(e[idx]), v[:, idx].conj().T)


def _vec_lowdin(c, s=1, f=None):
    if f is None:
        return np.dot(c, _lowdin(mdot(c.conj().T, s, c)))
    return np.dot(c * f, _lowdin(mdot(c.conj().T, s, c)))


def vec_lowdin(C, S, f=None):
    """lowdin.py:103-136 (only the S.ndim == 2 branches are reachable from get_emb_basis)."""
    S = np.asarray(S)
    assert S.ndim == 2
    if C.ndim == 2:
        return _vec_lowdin(C, S, f)
    C_orth = np.zeros_like(C)
    for s in range(C.shape[0]):
        C_orth[s] = _vec_lowdin(C[s], S, f if f is None else f[s])
    return C_orth


def check_span_same_space(a, b, ovlp=None, tol=1e-8):
    """Whether the columns of a and b span the same space (what the reference's own test compares baths
    with, libdmet/routine/test/test_slater.py:46-54): the projector of one must reproduce the other."""
    a = np.asarray(a)
    b = np.asarray(b)
    if ovlp is None:
        ovlp = np.eye(a.shape[0])
    if a.shape != b.shape:
        return False
    sa = mdot(a.conj().T, ovlp, a)
    pa = mdot(a, la.inv(sa), a.conj().T, ovlp)
    return max_abs(pa.dot(b) - b) < tol


# ---------------------------------------------------------------------------------------------------------
# get_emb_basis
# ---------------------------------------------------------------------------------------------------------
def get_emb_basis(lattice, rho=None, local=True, kind='svd', **kwargs):
    """slater.py:98-115."""
    if rho is None:
        rho = lattice.rdm1_lo_R
    assert local, "oracle restates the local branch only"
    if kind == 'svd':
        return _get_emb_basis_svd(lattice, np.asarray(rho).real, **kwargs)
    raise ValueError("get_emb_basis: Unknown kind %s" % kind)


def _get_emb_basis_svd(lattice, rdm1, **kwargs):
    """slater.py:117-220."""
    imp_idx = kwargs.get("imp_idx", lattice.imp_idx)
    val_idx = kwargs.get("val_idx", lattice.val_idx)
    valence_bath = kwargs.get("valence_bath", True)
    orth = kwargs.get("orth", True)
    tol_bath = kwargs.get("tol_bath", 1e-9)
    nbath = kwargs.get("nbath", None)

    ncells = lattice.ncells
    nlo = lattice.nscsites
    imp_idx_bath = val_idx if valence_bath else imp_idx
    env_idx = []
    virt_mask = []
    for i in range(ncells * nlo):
        if i not in imp_idx_bath:
            env_idx.append(i)
            virt_mask.append(i in imp_idx)
    nimp = len(imp_idx)

    rdm1 = np.asarray(rdm1)
    if rdm1.ndim == 3:
        rdm1 = rdm1[np.newaxis]
    assert rdm1.shape[-3:] == (ncells, nlo, nlo)
    spin = rdm1.shape[0]

    if np.max(imp_idx_bath) >= nlo - 1:
        rdm1_env_imp = lattice.expand(rdm1)[:, env_idx][:, :, imp_idx_bath]
        nbath_final = len(imp_idx_bath)
    else:
        rdm1_env_imp = rdm1.reshape(spin, ncells * nlo, nlo)[:, env_idx][:, :, imp_idx_bath]
        nbath_final = nlo
    basis = np.zeros((spin, ncells * nlo, nimp * 2))

    for s in range(spin):
        u, sigma, vt = la.svd(rdm1_env_imp[s], full_matrices=False)
        if nbath is None:
            nbath_s = (sigma >= tol_bath).sum()
        else:
            nbath_s = nbath
        B = u[:, :nbath_s]
        if nbath_s > 0:
            if orth:
                B[virt_mask] = 0.0
                B = vec_lowdin(B, np.eye(B.shape[0]))
        basis[s, imp_idx, :nimp] = np.eye(nimp)
        basis[s, env_idx, nimp:nimp + nbath_s] = B
        nbath_final = min(nbath_final, nbath_s)

    basis = basis[:, :, :nimp + nbath_final].reshape(spin, ncells, nlo, nimp + nbath_final)
    return basis


embBasis = get_emb_basis


# ---------------------------------------------------------------------------------------------------------
# libdmet/system/integral.py
# ---------------------------------------------------------------------------------------------------------
class Integral(object):
    """integral.py:61-105."""

    def __init__(self, norb, restricted, bogoliubov, H0, H1, H2, ovlp=None):
        self.norb = norb
        self.restricted = restricted
        self.bogoliubov = bogoliubov
        self.H0 = H0
        if isinstance(H1, np.ndarray):
            H1 = {"cd": H1}
        if isinstance(H2, np.ndarray):
            H2 = {"ccdd": H2}
        for key in H1:
            assert H1[key] is None or (H1[key].ndim == 3 and H1[key].shape[-1] == self.norb)
        self.H1 = H1
        for key in H2:
            if H2[key] is not None:
                assert H2[key].ndim in (5, 3, 2)
        self.H2 = H2
        self.ovlp = np.eye(self.norb) if ovlp is None else ovlp


def get_eri_format(eri, nao):
    """integral.py:883-928."""
    eri = np.asarray(eri)
    nao_pair = nao * (nao + 1) // 2
    s1_size = nao ** 4
    s4_size = nao_pair * nao_pair
    s8_size = nao_pair * (nao_pair + 1) // 2
    if eri.ndim == 5:
        eri_format, spin_dim = 's1', eri.size // s1_size
    elif eri.ndim == 4 and eri.size == s1_size:
        eri_format, spin_dim = 's1', 0
    elif eri.ndim == 3:
        eri_format, spin_dim = 's4', eri.size // s4_size
    elif eri.ndim == 2 and eri.size == s4_size:
        eri_format, spin_dim = 's4', 0
    elif eri.ndim == 2 and eri.size == s8_size:
        eri_format, spin_dim = 's8', 1
    elif eri.ndim == 1 and eri.size == s8_size:
        eri_format, spin_dim = 's8', 0
    else:
        raise ValueError("Unknown ERI shape %s, nao %s" % (str(eri.shape), nao))
    assert spin_dim in [0, 1, 3]
    return eri_format, spin_dim


# ---------------------------------------------------------------------------------------------------------
# libdmet/solver/scf.py:255-352
# ---------------------------------------------------------------------------------------------------------
def _get_jk(dm, eri, with_j=True, with_k=True):
    dm = np.asarray(dm, dtype=np.double)
    if dm.ndim == 2:
        dm = dm[np.newaxis]
    spin = dm.shape[0]
    nao = dm.shape[-1]
    eri = np.asarray(eri, dtype=np.double)
    eri_format, spin_dim = get_eri_format(eri, nao)
    if spin_dim == 0:
        eri = eri[None]
        spin_dim = 1
    if spin == 1 or spin_dim == 1:
        if eri_format == 's1':
            eri = lib.restore(8, eri[0], nao)
        else:
            eri = eri[0]
        vj, vk = lib.dot_eri_dm(eri, dm, hermi=1, with_j=with_j, with_k=with_k)
    elif spin_dim == 3:
        assert dm.shape[0] == 2
        eri_aa = lib.restore(4, eri[0], nao)
        vj00, vk00 = lib.dot_eri_dm(eri_aa, dm[0], hermi=1, with_j=with_j, with_k=with_k)
        eri_bb = lib.restore(4, eri[1], nao)
        vj11, vk11 = lib.dot_eri_dm(eri_bb, dm[1], hermi=1, with_j=with_j, with_k=with_k)
        eri_ab = lib.restore(4, eri[2], nao)
        vj01 = lib.dot_eri_dm(eri_ab, dm[1], hermi=1, with_j=with_j, with_k=False)[0]
        vj10 = lib.dot_eri_dm(eri_ab.T, dm[0], hermi=1, with_j=with_j, with_k=False)[0]
        vj = np.asarray(((vj00, vj11), (vj01, vj10)))
        vk = np.asarray((vk00, vk11))
    else:
        raise ValueError
    return vj, vk


def _get_veff(dm, eri):
    dm = np.asarray(dm, dtype=np.double)
    if dm.ndim == 2:
        dm = dm[np.newaxis]
    spin = dm.shape[0]
    vj, vk = _get_jk(dm, eri)
    if spin == 1:
        veff = vj - vk * 0.5
    else:
        veff = vj[0] + vj[1] - vk
    return veff


def get_veff(rdm1, eri, hyb=1.0):
    """slater.py:478-523, HF branch (hyb == 1.0, non-GHF)."""
    rdm1 = np.asarray(rdm1)
    if rdm1.ndim == 2:
        rdm1 = rdm1[None]
    assert hyb == 1.0
    return _get_veff(rdm1, eri)


# ---------------------------------------------------------------------------------------------------------
# libdmet/routine/slater_helper.py
# ---------------------------------------------------------------------------------------------------------
def transform_trans_inv_k(basis_k, H_k, warn=None):
    """slater_helper.py:37-50."""
    nkpts, nlo, nbasis = basis_k.shape
    res = np.zeros((nbasis, nbasis), dtype=np.complex128)
    for k in range(nkpts):
        res += mdot(basis_k[k].conj().T, H_k[k], basis_k[k])
    if max_abs(res.imag) > IMAG_DISCARD_TOL and warn is not None:
        warn.append(max_abs(res.imag))
    return res.real / float(nkpts)


def transform_local(basis, lattice, H):
    """slater_helper.py:73-80."""
    res = np.zeros((basis.shape[-1],) * 2)
    for i in range(lattice.ncells):
        res += mdot(basis[i].T, H, basis[i])
    return res


def transform_imp(basis, lattice, H):
    """slater_helper.py:102-103."""
    return mdot(basis[0].T, H, basis[0])


def init_H2(norb, symmetry, spin_dim):
    npair = norb * (norb + 1) // 2
    if symmetry == 1:
        return np.zeros((spin_dim, norb, norb, norb, norb))
    if symmetry == 4:
        return np.zeros((spin_dim, npair, npair))
    return np.zeros((spin_dim, npair * (npair + 1) // 2))


def unit2emb(H2_unit, neo):
    """slater_helper.py:494-517 (ndarray branch)."""
    spin_pair = H2_unit.shape[0]
    if H2_unit.ndim == 5:
        H2_emb = init_H2(neo, 1, spin_pair)
    elif H2_unit.ndim == 3:
        H2_emb = init_H2(neo, 4, spin_pair)
    elif H2_unit.ndim == 2:
        H2_emb = init_H2(neo, 8, spin_pair)
    else:
        raise ValueError
    fill_idx = tuple(map(slice, H2_unit.shape))
    H2_emb[fill_idx] = H2_unit
    return H2_emb


def transform_h1(H1_k, basis_k):
    """slater.py:690-697."""
    spin = basis_k.shape[0]
    nbasis = basis_k.shape[-1]
    H1_k = add_spin_dim(H1_k, spin, non_spin_dim=3)
    H1 = np.empty((spin, nbasis, nbasis))
    for s in range(spin):
        H1[s] = transform_trans_inv_k(basis_k[s], H1_k[s])
    return H1


foldRho_k = transform_h1   # slater.py:712


# ---------------------------------------------------------------------------------------------------------
# get_emb_Ham
# ---------------------------------------------------------------------------------------------------------
def _embHam2e(lattice, basis, vcor, local, int_bath=True, last_aabb=True, **kwargs):
    """slater.py:372-476, ab-initio branch (438-472)."""
    nbasis = basis.shape[-1]
    eri_symmetry = lattice.eri_symmetry
    max_memory = kwargs.get("max_memory", None)
    assert not lattice.is_model
    cell = lattice.cell
    mydf = lattice.df
    C_ao_lo = lattice.C_ao_lo
    kscaled_center = kwargs.get("kscaled_center", None)
    t_reversal_symm = kwargs.get("t_reversal_symm", True)
    if int_bath:
        H2 = eri_transform.get_emb_eri(cell, mydf, C_ao_lo=C_ao_lo, basis=basis, kscaled_center=kscaled_center,
                                       symmetry=eri_symmetry, max_memory=max_memory,
                                       t_reversal_symm=t_reversal_symm)
        if last_aabb and isinstance(H2, np.ndarray) and H2.shape[0] == 3:
            H2 = H2[[0, 2, 1]]
    else:
        H2 = eri_transform.get_unit_eri(cell, mydf, C_ao_lo=C_ao_lo, kscaled_center=kscaled_center,
                                        symmetry=eri_symmetry, max_memory=max_memory,
                                        t_reversal_symm=t_reversal_symm)
        if last_aabb and isinstance(H2, np.ndarray) and H2.shape[0] == 3:
            H2 = H2[[0, 2, 1]]
        H2 = unit2emb(H2, nbasis)
    return H2


def _embHam1e(lattice, basis, vcor, H2_emb, int_bath=True, add_vcor=False, **kwargs):
    """slater.py:525-688, interacting-bath Hartree-Fock branch (590-605, 639-643)."""
    assert int_bath, "oracle restates the interacting-bath branch only"
    spin = basis.shape[0]
    basis_k = lattice.R2k_basis(basis)
    hcore_k = lattice.hcore_lo_k
    ovlp_k = lattice.ovlp_lo_k
    hcore_emb = transform_h1(hcore_k, basis_k)
    ovlp_emb = transform_h1(ovlp_k, basis_k)
    if ovlp_emb.ndim == 3 and ovlp_emb.shape[0] == 1:
        ovlp_emb = ovlp_emb[0]
    rdm1_emb = foldRho_k(lattice.rdm1_lo_k, basis_k)
    fock_k = lattice.hcore_lo_k + lattice.vhf_lo_k
    H1 = transform_h1(fock_k, basis_k)
    JK_emb = get_veff(rdm1_emb, H2_emb)
    H1 -= JK_emb
    lattice.JK_core = H1 - hcore_emb
    if add_vcor:
        for s in range(spin):
            H1[s] += transform_local(basis[s], lattice, vcor.get()[s])
            if not kwargs.get("fitting", False):
                H1[s] -= transform_imp(basis[s], lattice, vcor.get()[s])
    return H1, ovlp_emb


def get_emb_Ham(lattice, basis, vcor, local=True, **kwargs):
    """slater.py:320-370."""
    basis = np.asarray(basis)
    spin = basis.shape[0]
    nbasis = basis.shape[-1]
    H2_given = kwargs.get("H2_given", None)
    if H2_given is None:
        H2 = _embHam2e(lattice, basis, vcor, local, **kwargs)
    else:
        H2 = H2_given
    H1, ovlp_emb = _embHam1e(lattice, basis, vcor, H2, **kwargs)
    H0 = lattice.getH0()
    if isinstance(H2, np.ndarray):
        H2 = {"ccdd": H2}
    ImpHam = Integral(nbasis, spin == 1, False, H0, {"cd": H1}, H2, ovlp=ovlp_emb)
    return ImpHam, None


embHam = get_emb_Ham


# ---------------------------------------------------------------------------------------------------------
# energy side (slater.py:1716-1840, 1957-2032)
# ---------------------------------------------------------------------------------------------------------
def get_H1_scaled(H1, imp_idx, env_idx=None):
    """slater.py:1716-1732."""
    assert H1.ndim == 3
    nbasis = H1.shape[-1]
    if env_idx is None:
        env_idx = np.asarray([idx for idx in range(nbasis) if idx not in imp_idx], dtype=int)
    imp_env = np.ix_(imp_idx, env_idx)
    env_imp = np.ix_(env_idx, imp_idx)
    env_env = np.ix_(env_idx, env_idx)
    for s in range(H1.shape[0]):
        H1[s][imp_env] *= 0.5
        H1[s][env_imp] *= 0.5
        H1[s][env_env] = 0.0
    return H1


def get_H2_scaled(H2, imp_idx, env_idx=None):
    """slater.py:1734-1778."""
    if H2.ndim == 3:
        nbasis_pair = H2.shape[-1]
        nbasis = int(np.sqrt(nbasis_pair * 2))
        tril_idx = np.tril_indices(nbasis)
        mask = np.isin(tril_idx, imp_idx)
        zero = np.logical_not(np.logical_or(*mask))
        half = np.logical_xor(*mask)
        one = np.logical_and(*mask)
        mask_list = (zero, half, one)
        for s in range(H2.shape[0]):
            for i, mask_i in enumerate(mask_list):
                for j, mask_j in enumerate(mask_list):
                    if i + j == 4:
                        continue
                    elif i + j == 0:
                        H2[s][np.ix_(mask_i, mask_j)] = 0.0
                    else:
                        H2[s][np.ix_(mask_i, mask_j)] *= ((i + j) * 0.25)
    elif H2.ndim == 5:
        nbasis = H2.shape[-1]
        if env_idx is None:
            env_idx = np.asarray([idx for idx in range(nbasis) if idx not in imp_idx], dtype=int)
        mask_list = (env_idx, imp_idx)
        for s in range(H2.shape[0]):
            for i, mi in enumerate(mask_list):
                for j, mj in enumerate(mask_list):
                    for k, mk in enumerate(mask_list):
                        for l, ml in enumerate(mask_list):
                            H2[s][np.ix_(mi, mj, mk, ml)] *= (i + j + k + l) * 0.25
    else:
        raise ValueError("Unknown H2 shape to scale: %s" % (str(H2.shape)))
    return H2


def get_H_dmet(basis, lattice, ImpHam, last_dmu, imp_idx=None, compact=True, **kwargs):
    """slater.py:1957-2032, default branch (E1, veff not given)."""
    spin = basis.shape[0]
    nbasis = basis.shape[-1]
    if imp_idx is None:
        imp_idx = list(range(len(lattice.imp_idx)))
    imp_idx = np.asarray(imp_idx)
    env_idx = np.asarray([idx for idx in range(nbasis) if idx not in imp_idx], dtype=int)
    basis_k = lattice.R2k_basis(basis)
    H1_scaled = transform_h1(lattice.hcore_lo_k, basis_k)
    JK_core = lattice.JK_core if lattice.JK_core is not None else [0.0 for s in range(spin)]
    for s in range(spin):
        H1_scaled[s] += 0.5 * JK_core[s]
    H1_scaled = get_H1_scaled(H1_scaled, imp_idx, env_idx)
    H0 = lattice.getH0()
    npair = nbasis * (nbasis + 1) // 2
    H2_scaled = np.empty((spin * (spin + 1) // 2, npair, npair))
    for s in range(spin * (spin + 1) // 2):
        H2_scaled[s] = lib.restore(4, ImpHam.H2["ccdd"][s], nbasis)
    H2_scaled = get_H2_scaled(H2_scaled, imp_idx, env_idx)
    if not compact:
        H2_scaled = np.stack([lib.restore(1, H2_scaled[s], nbasis) for s in range(H2_scaled.shape[0])])
    return Integral(nbasis, spin == 1, False, H0, {"cd": H1_scaled}, {"ccdd": H2_scaled})


def transformResults(rhoEmb, E, basis, ImpHam, H1e=None, **kwargs):
    """slater.py:1780-1840."""
    spin = rhoEmb.shape[0]
    nscsites = basis.shape[2]
    nbasis = basis.shape[-1]
    if "lattice" in kwargs:
        imp_idx = np.asarray(kwargs.get("imp_idx", range(len(kwargs["lattice"].imp_idx))))
    else:
        imp_idx = np.asarray(kwargs.get("imp_idx", np.arange(nscsites)))
    nelec = 0.0
    for s in range(spin):
        nelec += np.sum(rhoEmb[s, imp_idx, imp_idx])
    nelec *= (2.0 / spin)
    rhoImp = rhoEmb[np.ix_(range(spin), imp_idx, imp_idx)]
    if E is not None:
        lattice = kwargs["lattice"]
        last_dmu = kwargs["last_dmu"]
        imp_idx = np.asarray(kwargs.get("imp_idx", list(range(len(lattice.imp_idx)))))
        dmu_idx = kwargs.get("dmu_idx", None)
        if dmu_idx is None:
            dmu_idx = list(range(nscsites))
        env_idx = np.asarray([idx for idx in range(nbasis) if idx not in imp_idx], dtype=int)
        E2 = E - np.einsum('spq,sqp', ImpHam.H1["cd"], rhoEmb) * (2.0 / spin) - ImpHam.H0
        H1_scaled = np.array(ImpHam.H1["cd"], copy=True)
        dmu_mat = np.zeros((nscsites, nscsites))
        dmu_mat[dmu_idx, dmu_idx] = -last_dmu
        for s in range(spin):
            H1_scaled[s] -= transform_imp(basis[s], lattice, dmu_mat)
            if lattice.JK_core is not None:
                H1_scaled[s] -= 0.5 * lattice.JK_core[s]
        H1_scaled = get_H1_scaled(H1_scaled, imp_idx, env_idx)
        E1 = np.einsum('spq,sqp', H1_scaled, rhoEmb) * (2.0 / spin)
        Efrag = E1 + E2 + lattice.getH0()
    else:
        Efrag = None
    return rhoImp, Efrag, nelec
