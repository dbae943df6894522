"""Oracle: embedding basis, embedding Hamiltonian and the energy-side helpers of Slater-determinant DMET.

Behaviour restated (own numpy code) from
  libdmet/routine/slater.py          98-220 get_emb_basis (SVD bath), 320-370 get_emb_Ham, 438-476 ab-initio branch of
                                     __embHam2e, 478-523 get_veff (HF), 525-605 + 639-643 __embHam1e (interacting bath,
                                     HF), 690-712 transform_h1 / foldRho_k, 1716-1778 get_H1_scaled / get_H2_scaled,
                                     1780-1840 transformResults, 1957-2032 get_H_dmet (default branch)
  libdmet/routine/slater_helper.py   37-50 transform_trans_inv_k, 73-80, 102-103, 494-517 unit2emb
  libdmet/solver/scf.py              255-352 _get_jk / _get_veff
  libdmet/lo/lowdin.py               83-136 Loewdin orthogonalisation
  libdmet/system/integral.py         61-105 Integral, 883-928 get_eri_format
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`lattice` is duck-typed: .ncells .nscsites .imp_idx .val_idx .expand() .R2k_basis() .df .cell .C_ao_lo
.eri_symmetry .is_model .hcore_lo_k .vhf_lo_k .ovlp_lo_k .rdm1_lo_k .rdm1_lo_R .getH0() .JK_core
"""
import numpy as np
import scipy.linalg as la

from . import pyscf_lib as lib
from .fourier import max_abs, IMAG_DISCARD_TOL
from .make_basis import add_spin_dim
from . import eri_transform


# ---------------------------------------------------------------------------------------------------------
# small linear-algebra helpers
# ---------------------------------------------------------------------------------------------------------
def _inv_sqrt(s, tol=1e-14):
    """S^(-1/2) on the span of eigenvalues > tol (lowdin.py:83-92)"""
    w, v = la.eigh(s)
    keep = w > tol
    return (v[:, keep] / np.sqrt(w[keep])) @ v[:, keep].conj().T


def vec_lowdin(C, S=None):
    """symmetric orthonormalisation C (C^dagger S C)^(-1/2) of a set of vectors (lowdin.py:94-136, fixed metric)"""
    C = np.asarray(C)
    if C.ndim == 3:
        return np.stack([vec_lowdin(c, S) for c in C])
    metric = C.conj().T @ C if S is None else C.conj().T @ S @ C
    return C @ _inv_sqrt(metric)


def check_span_same_space(a, b, ovlp=None, tol=1e-8):
    """do the columns of a and b span the same space?  (the comparison the reference's own bath test makes,
    libdmet/routine/test/test_slater.py:46-54: orbitals are only defined up to rotations among themselves)"""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    S = np.eye(a.shape[0]) if ovlp is None else ovlp
    proj = a @ la.inv(a.conj().T @ S @ a) @ a.conj().T @ S
    return max_abs(proj @ b - b) < tol


# ---------------------------------------------------------------------------------------------------------
# get_emb_basis
# ---------------------------------------------------------------------------------------------------------
def get_emb_basis(lattice, rho=None, local=True, kind='svd', **kwargs):
    """dispatcher (slater.py:98-115); the local SVD and eigenvalue constructions are restated"""
    rho = lattice.rdm1_lo_R if rho is None else rho
    assert local, "oracle restates the local branch only"
    if kind == 'eig':
        return _get_emb_basis_eig(lattice, np.asarray(rho).real, **kwargs)
    if kind != 'svd':
        raise ValueError("get_emb_basis: Unknown kind %s" % kind)
    return _get_emb_basis_svd(lattice, np.asarray(rho).real, **kwargs)


def _get_emb_basis_eig(lattice, rdm1, imp_idx=None, val_idx=None, valence_bath=True, orth=True, tol_bath=1e-9,
                       **unused):
    """Bath orbitals = eigenvectors of the environment-environment block of the density matrix with occupation
    away from 0 and 1 (slater.py:224-318)."""
    imp_idx = list(lattice.imp_idx if imp_idx is None else imp_idx)
    val_idx = list(lattice.val_idx if val_idx is None else val_idx)
    ncells, nlo = lattice.ncells, lattice.nscsites
    ntot = ncells * nlo
    generators = val_idx if valence_bath else imp_idx
    env = np.setdiff1d(np.arange(ntot), generators)
    env_is_imp = np.isin(env, imp_idx)
    nimp = len(imp_idx)
    rdm1 = np.asarray(rdm1)
    rdm1 = rdm1[None] if rdm1.ndim == 3 else rdm1
    spin = rdm1.shape[0]
    block = lattice.expand(rdm1)[:, env][:, :, env]
    bath = []
    for s in range(spin):
        ew, ev = la.eigh(block[s])
        bath.append(np.asarray([ev[:, i] for i, e in enumerate(ew)
                                if abs(e) > tol_bath and abs(1 - e) > tol_bath]).T)
    bath = np.asarray(bath)                                     # (spin, nenv, nbath): equal counts per spin
    nbath = bath.shape[-1]
    out = np.zeros((spin, ntot, nimp + nbath))
    for s in range(spin):
        B = bath[s]
        if nbath > 0 and orth:
            B[env_is_imp] = 0.0
            B = vec_lowdin(B)
        out[s, imp_idx, :nimp] = np.eye(nimp)
        out[s, env, nimp:] = B
    return out.reshape(spin, ncells, nlo, nimp + nbath)


def _get_emb_basis_svd(lattice, rdm1, imp_idx=None, val_idx=None, valence_bath=True, orth=True, tol_bath=1e-9,
                       nbath=None, **unused):
    """Bath orbitals = left singular vectors of the environment x impurity block of the density matrix
    (slater.py:117-220).  With a valence bath only the valence orbitals generate bath states; the virtual impurity
    rows of the bath are then zeroed and the bath re-orthonormalised.  Output (spin, ncells, nlo, nimp + nbath) with
    the identity on the impurity rows."""
    imp_idx = list(lattice.imp_idx if imp_idx is None else imp_idx)
    val_idx = list(lattice.val_idx if val_idx is None else val_idx)
    ncells, nlo = lattice.ncells, lattice.nscsites
    ntot = ncells * nlo
    generators = val_idx if valence_bath else imp_idx
    env = np.setdiff1d(np.arange(ntot), generators)            # sorted, like the reference's scan
    env_is_imp = np.isin(env, imp_idx)
    nimp = len(imp_idx)

    rdm1 = np.asarray(rdm1)
    rdm1 = rdm1[None] if rdm1.ndim == 3 else rdm1
    assert rdm1.shape[1:] == (ncells, nlo, nlo)
    spin = rdm1.shape[0]
    # (sic) the reference switches to the expanded matrix already when the LAST orbital of cell 0 generates bath
    # states (`>= nlo - 1`, slater.py:167); it only changes the cap on the number of bath orbitals
    if max(generators) >= nlo - 1:
        coupling = lattice.expand(rdm1)[:, env][:, :, generators]
        cap = len(generators)
    else:
        coupling = rdm1.reshape(spin, ntot, nlo)[:, env][:, :, generators]
        cap = nlo
    out = np.zeros((spin, ntot, 2 * nimp))
    for s in range(spin):
        u, sigma, _ = la.svd(coupling[s], full_matrices=False)
        keep = int(np.count_nonzero(sigma >= tol_bath)) if nbath is None else nbath
        bath = u[:, :keep]
        if keep > 0 and orth:
            bath[env_is_imp] = 0.0
            bath = vec_lowdin(bath)
        out[s, imp_idx, :nimp] = np.eye(nimp)
        out[s, env, nimp:nimp + keep] = bath
        cap = min(cap, keep)
    return out[:, :, :nimp + cap].reshape(spin, ncells, nlo, nimp + cap)


embBasis = get_emb_basis


# ---------------------------------------------------------------------------------------------------------
# Integral container and ERI layout detection
# ---------------------------------------------------------------------------------------------------------
class Integral(object):
    """what embHam returns: norb, restricted, bogoliubov, H0, H1 {"cd"}, H2 {"ccdd"}, ovlp (integral.py:61-105)"""

    def __init__(self, norb, restricted, bogoliubov, H0, H1, H2, ovlp=None):
        self.norb, self.restricted, self.bogoliubov, self.H0 = norb, restricted, bogoliubov, H0
        self.H1 = {"cd": H1} if isinstance(H1, np.ndarray) else H1
        self.H2 = {"ccdd": H2} if isinstance(H2, np.ndarray) else H2
        for v in self.H1.values():
            assert v is None or (v.ndim == 3 and v.shape[-1] == norb)
        for v in self.H2.values():
            assert v is None or v.ndim in (5, 3, 2)
        self.ovlp = np.eye(norb) if ovlp is None else ovlp


def get_eri_format(eri, nao):
    """('s1' | 's4' | 's8', spin_dim in {0, 1, 3}) from the array rank and size (integral.py:883-928)"""
    eri = np.asarray(eri)
    npair = nao * (nao + 1) // 2
    size = {'s1': nao ** 4, 's4': npair * npair, 's8': npair * (npair + 1) // 2}
    if eri.ndim == 5:
        fmt, spin_dim = 's1', eri.size // size['s1']
    elif eri.ndim == 4 and eri.size == size['s1']:
        fmt, spin_dim = 's1', 0
    elif eri.ndim == 3:
        fmt, spin_dim = 's4', eri.size // size['s4']
    elif eri.ndim == 2 and eri.size == size['s4']:
        fmt, spin_dim = 's4', 0
    elif eri.ndim == 2 and eri.size == size['s8']:
        fmt, spin_dim = 's8', 1
    elif eri.ndim == 1 and eri.size == size['s8']:
        fmt, spin_dim = 's8', 0
    else:
        raise ValueError("Unknown ERI shape %s, nao %s" % (str(eri.shape), nao))
    assert spin_dim in (0, 1, 3)
    return fmt, spin_dim


# ---------------------------------------------------------------------------------------------------------
# J/K and the HF effective potential in the embedding space (solver/scf.py:255-352)
# ---------------------------------------------------------------------------------------------------------
def _get_jk(dm, eri, with_j=True, with_k=True):
    """PySCF convention J_ij = (ij|kl) D_kl, K_jk = (ij|kl) D_il.  One ERI block: J, K per density matrix.  Three blocks
    (aa, bb, ab as embHam orders them): vj = ((J_a[D_a], J_b[D_b]), (J_a[D_b], J_b[D_a])), vk = (K_a, K_b)."""
    dm = np.asarray(dm, dtype=float)
    dm = dm[None] if dm.ndim == 2 else dm
    n = dm.shape[-1]
    eri = np.asarray(eri, dtype=float)
    fmt, spin_dim = get_eri_format(eri, n)
    if spin_dim == 0:
        eri, spin_dim = eri[None], 1
    if dm.shape[0] == 1 or spin_dim == 1:
        block = lib.restore(8, eri[0], n) if fmt == 's1' else eri[0]
        return lib.dot_eri_dm(block, dm, hermi=1, with_j=with_j, with_k=with_k)
    if spin_dim != 3:
        raise ValueError
    assert dm.shape[0] == 2
    aa, bb, ab = (lib.restore(4, eri[i], n) for i in range(3))
    vj_aa, vk_aa = lib.dot_eri_dm(aa, dm[0], hermi=1, with_j=with_j, with_k=with_k)
    vj_bb, vk_bb = lib.dot_eri_dm(bb, dm[1], hermi=1, with_j=with_j, with_k=with_k)
    vj_ab = lib.dot_eri_dm(ab, dm[1], hermi=1, with_j=with_j, with_k=False)[0]       # alpha feels beta density
    vj_ba = lib.dot_eri_dm(ab.T, dm[0], hermi=1, with_j=with_j, with_k=False)[0]     # beta feels alpha density
    return np.asarray(((vj_aa, vj_bb), (vj_ab, vj_ba))), np.asarray((vk_aa, vk_bb))


def _get_veff(dm, eri):
    """restricted (spin-traced dm): J - K/2; unrestricted: J[a] + J[b] - K per spin (scf.py:334-352)"""
    dm = np.asarray(dm, dtype=float)
    dm = dm[None] if dm.ndim == 2 else dm
    vj, vk = _get_jk(dm, eri)
    return vj - 0.5 * vk if dm.shape[0] == 1 else vj[0] + vj[1] - vk


def get_veff(rdm1, eri, hyb=1.0):
    """slater.py:478-523, HF branch"""
    assert hyb == 1.0
    rdm1 = np.asarray(rdm1)
    return _get_veff(rdm1[None] if rdm1.ndim == 2 else rdm1, eri)


# ---------------------------------------------------------------------------------------------------------
# one-body transforms into the embedding space
# ---------------------------------------------------------------------------------------------------------
def transform_trans_inv_k(basis_k, H_k, warn=None):
    """Re[ sum_k B_k^dagger H_k B_k ] / nkpts; the imaginary-part warning (> 1e-7) goes to `warn`
    (slater_helper.py:37-50)"""
    total = np.einsum("kpm,kpq,kqn->mn", np.conj(basis_k), H_k, basis_k)
    if warn is not None and max_abs(total.imag) > IMAG_DISCARD_TOL:
        warn.append(max_abs(total.imag))
    return total.real / float(len(basis_k))


def transform_local(basis, lattice, H):
    """sum over cells of B_R^T H B_R (slater_helper.py:73-80)"""
    return np.einsum("Rpm,pq,Rqn->mn", basis, H, basis)


def transform_imp(basis, lattice, H):
    """impurity-cell term B_0^T H B_0 (slater_helper.py:102-103)"""
    return basis[0].T @ H @ basis[0]


def unit2emb(H2_unit, neo):
    """zero-pad the impurity-cell ERI to neo embedding orbitals, same symmetry layout (slater_helper.py:494-517)"""
    H2_unit = np.asarray(H2_unit)
    npair = neo * (neo + 1) // 2
    tail = {5: (neo,) * 4, 3: (npair, npair), 2: (npair * (npair + 1) // 2,)}.get(H2_unit.ndim)
    if tail is None:
        raise ValueError
    out = np.zeros((H2_unit.shape[0],) + tail)
    out[tuple(slice(0, n) for n in H2_unit.shape)] = H2_unit
    return out


def transform_h1(H1_k, basis_k):
    """per-spin transform_trans_inv_k with spin broadcasting of H1_k (slater.py:690-697)"""
    H1_k = add_spin_dim(H1_k, basis_k.shape[0], non_spin_dim=3)
    return np.stack([transform_trans_inv_k(basis_k[s], H1_k[s]) for s in range(basis_k.shape[0])])


foldRho_k = transform_h1   # slater.py:712


# ---------------------------------------------------------------------------------------------------------
# get_emb_Ham
# ---------------------------------------------------------------------------------------------------------
def _embHam2e(lattice, basis, vcor, local, int_bath=True, last_aabb=True, **kwargs):
    """two-body part, ab-initio branch (slater.py:438-476): interacting bath -> get_emb_eri, otherwise the unit ERI
    zero-padded; unrestricted blocks reordered aa, ab, bb -> aa, bb, ab"""
    assert not lattice.is_model
    common = dict(C_ao_lo=lattice.C_ao_lo, kscaled_center=kwargs.get("kscaled_center", None),
                  symmetry=lattice.eri_symmetry, max_memory=kwargs.get("max_memory", None),
                  t_reversal_symm=kwargs.get("t_reversal_symm", True))
    if int_bath:
        H2 = eri_transform.get_emb_eri(lattice.cell, lattice.df, basis=basis, **common)
    else:
        H2 = eri_transform.get_unit_eri(lattice.cell, lattice.df, **common)
    if last_aabb and H2.shape[0] == 3:
        H2 = H2[[0, 2, 1]]
    return H2 if int_bath else unit2emb(H2, basis.shape[-1])


def _embHam1e(lattice, basis, vcor, H2_emb, int_bath=True, add_vcor=False, **kwargs):
    """one-body part, interacting bath + Hartree-Fock (slater.py:525-547, 559-560, 590-605, 639-643):
    H1 = T[hcore + vhf] - veff_emb[folded density], JK_core = H1 - T[hcore] stored on the lattice"""
    assert int_bath, "oracle restates the interacting-bath branch only"
    basis_k = lattice.R2k_basis(basis)
    hcore_emb = transform_h1(lattice.hcore_lo_k, basis_k)
    ovlp_emb = transform_h1(lattice.ovlp_lo_k, basis_k)
    if ovlp_emb.shape[0] == 1:
        ovlp_emb = ovlp_emb[0]
    rdm1_emb = foldRho_k(lattice.rdm1_lo_k, basis_k)
    H1 = transform_h1(lattice.hcore_lo_k + lattice.vhf_lo_k, basis_k) - get_veff(rdm1_emb, H2_emb)
    lattice.JK_core = H1 - hcore_emb
    if add_vcor:
        for s in range(basis.shape[0]):
            v = vcor.get()[s]
            H1[s] += transform_local(basis[s], lattice, v)
            if not kwargs.get("fitting", False):
                H1[s] -= transform_imp(basis[s], lattice, v)
    return H1, ovlp_emb


def get_emb_Ham(lattice, basis, vcor, local=True, **kwargs):
    """embedding Hamiltonian as an Integral, plus the deprecated None (slater.py:320-370)"""
    basis = np.asarray(basis)
    H2 = kwargs.get("H2_given", None)
    if H2 is None:
        H2 = _embHam2e(lattice, basis, vcor, local, **kwargs)
    H1, ovlp_emb = _embHam1e(lattice, basis, vcor, H2, **kwargs)
    nbasis = basis.shape[-1]
    return Integral(nbasis, basis.shape[0] == 1, False, lattice.getH0(), {"cd": H1}, {"ccdd": H2},
                    ovlp=ovlp_emb), None


embHam = get_emb_Ham


# ---------------------------------------------------------------------------------------------------------
# energy side
# ---------------------------------------------------------------------------------------------------------
def _complement(nbasis, imp_idx):
    return np.setdiff1d(np.arange(nbasis), np.asarray(imp_idx, dtype=int))


def get_H1_scaled(H1, imp_idx, env_idx=None):
    """in place: weight = (number of impurity indices) / 2 (slater.py:1716-1732)"""
    assert H1.ndim == 3
    member = np.zeros(H1.shape[-1])
    member[np.asarray(imp_idx, dtype=int)] = 1.0
    H1 *= 0.5 * (member[:, None] + member[None, :])
    return H1


def get_H2_scaled(H2, imp_idx, env_idx=None):
    """in place: weight = (number of impurity indices among the four) / 4, s4 (3-d) or s1 (5-d) layout
    (slater.py:1734-1778)"""
    if H2.ndim == 3:
        n = int(np.sqrt(H2.shape[-1] * 2))
        member = np.zeros(n)
        member[np.asarray(imp_idx, dtype=int)] = 1.0
        r, c = np.tril_indices(n)
        wpair = member[r] + member[c]
        H2 *= 0.25 * (wpair[:, None] + wpair[None, :])
    elif H2.ndim == 5:
        member = np.zeros(H2.shape[-1])
        member[np.asarray(imp_idx, dtype=int)] = 1.0
        m = member
        H2 *= 0.25 * (m[:, None, None, None] + m[None, :, None, None] + m[None, None, :, None] + m[None, None, None, :])
    else:
        raise ValueError("Unknown H2 shape to scale: %s" % (str(H2.shape)))
    return H2


def get_H_dmet(basis, lattice, ImpHam, last_dmu, imp_idx=None, compact=True, **kwargs):
    """DMET Hamiltonian weighted by impurity-index counts, default branch (slater.py:1957-2032):
    H1 = scaled( T[hcore] + JK_core / 2 ), H2 = scaled( s4 ERI ), H0 = lattice H0"""
    spin, nbasis = basis.shape[0], basis.shape[-1]
    imp_idx = np.arange(len(lattice.imp_idx)) if imp_idx is None else np.asarray(imp_idx)
    H1 = transform_h1(lattice.hcore_lo_k, lattice.R2k_basis(basis))
    if lattice.JK_core is not None:
        H1 += 0.5 * np.asarray(lattice.JK_core)
    H1 = get_H1_scaled(H1, imp_idx)
    H2 = np.stack([lib.restore(4, blk, nbasis) for blk in ImpHam.H2["ccdd"]]).copy()
    H2 = get_H2_scaled(H2, imp_idx)
    if not compact:
        H2 = np.stack([lib.restore(1, blk, nbasis) for blk in H2])
    return Integral(nbasis, spin == 1, False, lattice.getH0(), {"cd": H1}, {"ccdd": H2})


def transformResults(rhoEmb, E, basis, ImpHam, H1e=None, **kwargs):
    """impurity density, fragment energy and electron number from the solver's density matrix
    (slater.py:1780-1840).  E2 = E - <H1> - H0 is the two-body part of the solver energy; the one-body part is
    re-evaluated with the chemical-potential shift and half of JK_core removed and the impurity weights applied."""
    spin, nscsites, nbasis = rhoEmb.shape[0], basis.shape[2], basis.shape[-1]
    lattice = kwargs.get("lattice", None)
    default_imp = np.arange(len(lattice.imp_idx)) if lattice is not None else np.arange(nscsites)
    imp_idx = np.asarray(kwargs.get("imp_idx", default_imp))
    per_spin = 2.0 / spin
    nelec = per_spin * sum(np.trace(rhoEmb[s][np.ix_(imp_idx, imp_idx)]) for s in range(spin))
    rhoImp = rhoEmb[:, imp_idx][:, :, imp_idx]
    if E is None:
        return rhoImp, None, nelec
    dmu_idx = kwargs.get("dmu_idx", None)
    dmu_idx = list(range(nscsites)) if dmu_idx is None else dmu_idx
    H1 = ImpHam.H1["cd"]
    E2 = E - per_spin * np.einsum("spq,sqp", H1, rhoEmb) - ImpHam.H0
    shift = np.zeros((nscsites, nscsites))
    shift[dmu_idx, dmu_idx] = -kwargs["last_dmu"]
    H1s = np.array(H1, copy=True)
    for s in range(spin):
        H1s[s] -= transform_imp(basis[s], lattice, shift)
        if lattice.JK_core is not None:
            H1s[s] -= 0.5 * lattice.JK_core[s]
    H1s = get_H1_scaled(H1s, imp_idx)
    E1 = per_spin * np.einsum("spq,sqp", H1s, rhoEmb)
    return rhoImp, E1 + E2 + lattice.getH0(), nelec


# ---------------------------------------------------------------------------------------------------------
# global density matrix by democratic partitioning (slater_helper.py:158-283)
# ---------------------------------------------------------------------------------------------------------
def get_emb_basis_other_cell(lattice, basis, R):
    """embedding basis of the impurity problem sitting in cell R: cell I of it is cell I - R of the original
    (slater_helper.py:158-181)"""
    basis = np.asarray(basis)
    order = [lattice.subtract(I, R) for I in range(basis.shape[-3])]
    return basis[..., order, :, :]


def get_rho_glob_R(basis, lattice, rho_emb, compact=True):
    """stripe-shaped global density matrix (spin, ncells, nlo, nlo): for every cell R the full product of the
    translated embedding basis with the embedded density matrix, impurity-environment blocks halved,
    environment-environment blocks dropped, summed over R and over fragments (slater_helper.py:183-270, compact
    branch; lists of bases / lattices / density matrices = several fragments)"""
    assert compact
    if isinstance(lattice, (list, tuple)):
        frags = list(zip(basis, lattice, rho_emb))
    else:
        frags = [(basis, lattice, rho_emb)]
    total = 0.0
    for basis_I, lat, rho_I in frags:
        basis_I = np.asarray(basis_I)
        basis_I = basis_I[None] if basis_I.ndim == 3 else basis_I
        spin, ncells, nlo, _ = basis_I.shape
        rho_I = np.asarray(rho_I)
        rho_I = rho_I[None] if rho_I.ndim == 2 else rho_I
        if rho_I.shape[0] < spin:
            rho_I = np.asarray([rho_I[0]] * spin)
        rho_R = np.zeros((spin, ncells * nlo, nlo))
        for R in range(ncells):
            other = get_emb_basis_other_cell(lat, basis_I, R)
            imp = np.asarray(lat.imp_idx) + R * nlo
            env = np.setdiff1d(np.arange(ncells * nlo), imp)
            in_cell0 = np.isin(np.arange(nlo), imp)
            imp0, env0 = np.flatnonzero(in_cell0), np.flatnonzero(~in_cell0)
            for s in range(spin):
                C_R = other[s].reshape(ncells * nlo, -1)
                block = C_R.dot(rho_I[s]).dot(C_R[:nlo].conj().T)
                block[np.ix_(imp, env0)] *= 0.5
                block[np.ix_(env, imp0)] *= 0.5
                block[np.ix_(env, env0)] = 0.0
                rho_R[s] += block
        total = total + rho_R.reshape(spin, ncells, nlo, nlo)
    return total
