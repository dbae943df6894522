"""Generalised-spin-orbital (GSO / BCS-type) embedding Hamiltonian -- drop-in for the ab-initio, interacting-bath,
Hartree-Fock branch of `libdmet.routine.spinless.get_emb_Ham` (`embHam`), spinless.py:433-462, 464-558 (two-body part
via `get_emb_eri_gso`) and 560-726 (one-body part), with the helpers of spinless_helper.py:31-46 (`separate_basis`),
349-438 (`transform_trans_inv_k`, `transform_local`, `transform_imp`), and for the GSO bath construction
`get_emb_basis` (spinless.py:34-272, kinds 'svd' and 'eig'; host LAPACK like the reference).

A GSO basis has 2*nao rows per cell (alpha rows, then beta rows).  Lattice quantities come as (3, nkpts, nao, nao)
stacks (aa, bb, ab); the density matrix is one (nkpts, 2*nao, 2*nao) generalised matrix.  The two-body integrals stay
on the GPU between the ERI build and the J - K contraction; all contractions run in libldm_b200.so.
"""
import warnings

import numpy as np
import torch

from ._lib import UnsupportedBranch
from .device import get_device
from .eri_transform import get_emb_eri_gso, separate_basis
from .fourier import IMAG_DISCARD_TOL
from .integral import Integral
from . import slater
from .make_basis import sandwich


def _zdev(a):
    return slater._zdev(a)


# ---------------------------------------------------------------------------------------------------------
# GSO bath construction (spinless.py:34-272): host gather + LAPACK, as the reference (and as slater.get_emb_basis)
# ---------------------------------------------------------------------------------------------------------
def _gso_index_sets(lattice, valence_bath):
    """spin-orbital index sets of the generalised problem: orbital i of cell R has the alpha row R*nso + i and the
    beta row R*nso + nlo + i.  Returns (imp, env, env_is_imp, env_is_alpha)."""
    ncells, nlo = int(lattice.ncells), int(lattice.nscsites)
    nso = 2 * nlo
    both = lambda idx: [int(i) for i in idx] + [int(i) + nlo for i in idx]       # noqa: E731
    imp = both(lattice.imp_idx)
    gen = both(lattice.val_idx) if valence_bath else imp
    is_gen = np.zeros(ncells * nso, dtype=bool)
    is_gen[gen] = True
    is_imp = np.zeros(ncells * nso, dtype=bool)
    is_imp[imp] = True
    env = np.flatnonzero(~is_gen)
    return imp, gen, env, is_imp[env], (env % nso) < nlo


def _gso_finish(lattice, bath, imp, env, env_is_imp, env_is_alpha, orth, kwargs):
    """shared tail of both constructions: virtual rows zeroed + Loewdin, columns ordered by decreasing alpha
    (particle) weight, identity on the impurity rows (spinless.py:124-162 / 240-271)"""
    if kwargs.get("localize_bath") is not None:
        raise UnsupportedBranch("bath localisation is only defined for model Hamiltonians in the reference")
    if not orth:
        raise NotImplementedError                                   # l.129-131
    ncells, nso = int(lattice.ncells), 2 * int(lattice.nscsites)
    nbath = bath.shape[1]
    if nbath % 2 != 0:
        raise ValueError("nbath (%s) should be even in GSO." % nbath)            # l.114
    if nbath > 0:
        bath[env_is_imp] = 0.0
        bath = slater.vec_lowdin(bath)
    weight = np.einsum("ai,ai->i", bath[env_is_alpha], bath[env_is_alpha])
    order = np.argsort(weight, kind="mergesort")[::-1]
    nimp = len(imp)
    basis = np.zeros((ncells * nso, nimp + nbath))
    basis[imp, :nimp] = np.eye(nimp)
    basis[env, nimp:] = bath[:, order]
    return basis.reshape(ncells, nso, nimp + nbath)


def get_emb_basis(lattice, GRho, local=True, kind='svd', **kwargs):
    """spinless.py:34-53: embedding basis (ncells, 2*nlo, nimp + nbath) of the generalised density matrix
    GRho (ncells, 2*nlo, 2*nlo); kind 'svd' (58-162) or 'eig' (167-272)."""
    if not local:
        raise NotImplementedError
    if kwargs.get("bath_opt", False):
        raise UnsupportedBranch("bath optimisation (get_emb_basis_opt) is outside the path")
    rdm1 = np.asarray(GRho.cpu() if isinstance(GRho, torch.Tensor) else GRho).real
    ncells, nso = int(lattice.ncells), 2 * int(lattice.nscsites)
    assert rdm1.shape == (ncells, nso, nso)
    valence_bath = kwargs.get("valence_bath", True)
    orth = kwargs.get("orth", True)
    tol_bath = kwargs.get("tol_bath", 1e-9)
    imp, gen, env, env_is_imp, env_is_alpha = _gso_index_sets(lattice, valence_bath)
    import scipy.linalg as la
    if kind == 'svd':
        coupling = rdm1.reshape(ncells * nso, nso)[env][:, gen]
        u, sigma, _ = la.svd(coupling, full_matrices=False)
        nbath = kwargs.get("nbath", None)
        nbath = int(np.count_nonzero(sigma >= tol_bath)) if nbath is None else int(nbath)
        if np.any(np.abs(sigma[:nbath]) < tol_bath):
            warnings.warn("Zero singular value exists, \nthis may cause numerical instability.")
        bath = u[:, :nbath]
    elif kind == 'eig':
        occ, vec = la.eigh(lattice.expand(rdm1)[env][:, env])
        bath = vec[:, (np.abs(occ) > tol_bath) & (np.abs(1.0 - occ) > tol_bath)]
    elif kind == 'ph':
        raise UnsupportedBranch("particle-hole bath (model Hamiltonians) is outside the ab-initio path")
    else:
        raise ValueError("get_emb_basis: Unknown kind %s" % kind)
    return _gso_finish(lattice, np.array(bath), imp, env, env_is_imp, env_is_alpha, orth, kwargs)


embBasis = get_emb_basis


def transform_trans_inv_k_dev(basis_ka, basis_kb, H_k):
    """Re sum_k [ Ba^dagger H_aa Ba + Bb^dagger H_bb Bb + (Ba^dagger H_ab Bb + h.c.) ] / nkpts  -> (nbasis, nbasis)
    float64 device tensor (spinless_helper.py:349-381).  H_k: (2 or 3, nkpts, nao, nao)."""
    dev = get_device()
    H = _zdev(H_k)
    assert H.dim() == 4 and H.shape[0] in (2, 3)
    Ba, Bb = _zdev(basis_ka), _zdev(basis_kb)
    nk, nao, nb = Ba.shape
    BaT = dev.ztranspose(Ba)                     # (nk, nbasis, nao): k-contiguous rows
    BbT = dev.ztranspose(Bb)
    coef = torch.cat([BaT, BbT]).contiguous()    # slices 0..nk-1 = alpha, nk..2nk-1 = beta
    H3 = H.reshape(-1, nao, nao)
    k = np.arange(nk)
    # aa and bb: batch entry b uses coefficient slice b (left, conjugated, and right) and H slice b
    terms = sandwich(coef, True, coef, False, H3, h_index=np.concatenate([k, nk + k]),
                     a_index=np.concatenate([k, nk + k]))
    res, imag = dev.ksum_real(terms.contiguous(), scale=1.0)
    worst = imag
    if H.shape[0] == 3:
        # ab: Ba^dagger H_ab Bb; its Hermitian conjugate is added on the (real part of the) k-sum
        VT = dev.empty((nk, nb, nao), torch.complex128)
        segs = np.zeros((nk, 4), dtype=np.int32)
        segs[:, 0] = nk + k                       # right factor Bb
        segs[:, 1] = 2 * nk + k                   # H_ab
        dev.zgemm_tn(coef, H3, segs, VT, c_off=np.arange(nk, dtype=np.int64) * nb * nao, s_outer=nao, nbatch=nk,
                     nseg=1)
        out = dev.empty((nk, nb, nb), torch.complex128)
        segs2 = np.zeros((nk, 4), dtype=np.int32)
        segs2[:, 0] = k                           # left factor Ba, conjugated
        segs2[:, 1] = k
        segs2[:, 2] = 1
        dev.zgemm_tn(coef, VT, segs2, out, c_off=np.arange(nk, dtype=np.int64) * nb * nb, s_outer=nb, nbatch=nk,
                     nseg=1)
        ab, imag_ab = dev.ksum_real(out, scale=1.0)
        # Re(T + T^dagger) = Re T + (Re T)^T ; Im(T + T^dagger) is antisymmetric and bounded by 2 max|Im T|
        res = res + ab + ab.t()
        worst = worst + 2.0 * imag_ab
    if worst > IMAG_DISCARD_TOL:
        warnings.warn("transform_trans_inv_k: has imag part %s" % worst)
    return res / float(nk)


def transform_trans_inv_k(basis_ka, basis_kb, H_k):
    """spinless_helper.py:349-381, numpy in / numpy out"""
    return transform_trans_inv_k_dev(basis_ka, basis_kb, H_k).cpu().numpy()


def transform_local(basis_Ra, basis_Rb, H):
    """sum over cells of Ba^T H_aa Ba + Bb^T H_bb Bb (+ Ba^T H_ab Bb + h.c.)   (spinless_helper.py:383-409)"""
    H = np.asarray(H)
    assert H.shape[0] in (2, 3)
    res = np.einsum("Rpm,pq,Rqn->mn", basis_Ra.conj(), H[0], basis_Ra) + \
        np.einsum("Rpm,pq,Rqn->mn", basis_Rb.conj(), H[1], basis_Rb)
    if H.shape[0] == 3:
        ab = np.einsum("Rpm,pq,Rqn->mn", basis_Ra.conj(), H[2], basis_Rb)
        res = res + ab + ab.conj().T
    return res


def transform_imp(basis_Ra, basis_Rb, H):
    """the cell-0 term of transform_local (spinless_helper.py:411-438)"""
    return transform_local(basis_Ra[:1], basis_Rb[:1], H)


def foldRho_k(GRho_k, basis_k):
    """generalised density matrix folded to the embedding space (spinless.py:727-737)"""
    return slater.transform_trans_inv_k(basis_k, GRho_k)


def _embHam2e(lattice, basis, vcor, local, int_bath=True, last_aabb=True, **kwargs):
    """spinless.py:464-558, ab-initio interacting bath: one GSO ERI block, built in s4 on the device"""
    if getattr(lattice, "is_model", False):
        raise UnsupportedBranch("model Hamiltonians are outside the ab-initio hot path")
    if not int_bath:
        raise UnsupportedBranch("the reference's non-interacting-bath GSO branch feeds a (1, npair, npair) unit ERI "
                                  "to a 3-block unit2emb (spinless_helper.py:288-313) and cannot run; not mirrored")
    nb = basis.shape[-1]
    eri4 = get_emb_eri_gso(lattice.cell, lattice.df, C_ao_lo=lattice.C_ao_lo, basis=basis,
                           kscaled_center=kwargs.get("kscaled_center", None), symmetry=4,
                           t_reversal_symm=kwargs.get("t_reversal_symm", True), return_device=True,
                           **{k: kwargs[k] for k in ("source", "group", "kl_group", "stats") if k in kwargs})
    dev = get_device()
    sym = lattice.eri_symmetry
    if sym == 4:
        H2 = dev.to_host(eri4)
    elif sym == 1:
        H2 = dev.to_host(dev.restore_s1(eri4[0], nb))[None]
    elif sym == 8:
        H2 = dev.to_host(dev.restore_s8(eri4[0], nb))[None]
    else:
        raise ValueError("unknown eri_symmetry %s" % sym)
    return H2, [eri4[0]]


def _embHam1e(lattice, basis, vcor, mu, H2_emb, eri4_blocks, int_bath=True, add_vcor=False, **kwargs):
    """spinless.py:560-726, interacting bath, Hartree-Fock: H1 = T[fock_hf] - (J - K)[folded density] - mu N
    (+ optional local terms); side effect lattice.JK_core.  Returns (H1 (1, nbasis, nbasis), ovlp_emb)."""
    if not int_bath or kwargs.get("dft", False):
        raise UnsupportedBranch("only the interacting-bath Hartree-Fock branch is mirrored")
    if vcor is not None and hasattr(vcor, "islocal") and not vcor.islocal():
        raise Exception("nonlocal correlation potential cannot be treated in this routine")
    basis = np.asarray(basis)
    nao = int(lattice.nscsites)
    nb = basis.shape[-1]
    basis_k = lattice.R2k_basis(basis)
    basis_Ra, basis_Rb = separate_basis(basis)
    basis_ka, basis_kb = separate_basis(basis_k)
    hcore_k = kwargs.get("hcore_custom", None)
    hcore_k = lattice.getH1(kspace=True) if hcore_k is None else hcore_k
    hcore_emb = transform_trans_inv_k_dev(basis_ka, basis_kb, hcore_k)
    ovlp_emb = transform_trans_inv_k_dev(basis_ka, basis_kb, lattice.get_ovlp(kspace=True)).cpu().numpy()
    bk = slater._BasisK(np.asarray(basis_k)[None])
    rdm1_emb = slater.transform_trans_inv_k_dev(bk, 0, _zdev(lattice.rdm1_lo_k), 0)     # foldRho_k
    H1 = transform_trans_inv_k_dev(basis_ka, basis_kb, lattice.fock_hf_lo_k)
    hcore_add = kwargs.get("hcore_add", None)
    if hcore_add is not None:
        H1 = H1 + get_device().to_device(transform_imp(basis_Ra, basis_Rb, hcore_add).real, torch.float64)
    if eri4_blocks is None:
        eri4_blocks = slater._s4_blocks_dev(H2_emb, nb)
    vj, vk = get_device().jk_s4(eri4_blocks[0], rdm1_emb.contiguous())                 # GHF: J - K (scf.py:732-740)
    H1 = H1 - (vj - vk)
    lattice.JK_core = (H1 - hcore_emb).cpu().numpy()
    H1 = H1.cpu().numpy()
    mu_mat = np.zeros((2, nao, nao))
    np.fill_diagonal(mu_mat[0], -mu)
    np.fill_diagonal(mu_mat[1], mu)
    H1 += transform_local(basis_Ra, basis_Rb, mu_mat).real
    if add_vcor:
        H1 += transform_local(basis_Ra, basis_Rb, vcor.get()).real
        if not kwargs.get("fitting", False):
            H1 -= transform_imp(basis_Ra, basis_Rb, vcor.get()).real
        JK_imp = lattice.get_JK_imp()
        if JK_imp is not None:
            H1 -= transform_imp(basis_Ra, basis_Rb, JK_imp).real
    return H1[np.newaxis], ovlp_emb


def get_emb_Ham(lattice, basis, vcor, mu, local=True, **kwargs):
    """spinless.py:433-462.  Returns (Integral, None) with H1 {"cd": (1, n, n)} and H2 {"ccdd": (1, ...)}."""
    basis = np.asarray(basis)
    nb = basis.shape[-1]
    H2_given = kwargs.get("H2_given", None)
    blocks = None
    # unsupported branches are refused before any ERI work (see slater._check_supported)
    if getattr(lattice, "is_model", False) and H2_given is None and kwargs.get("H2_fname", None) is None:
        raise UnsupportedBranch("model Hamiltonians are outside the ab-initio hot path")
    if not kwargs.get("int_bath", True) or kwargs.get("dft", False):
        raise UnsupportedBranch("only the interacting-bath Hartree-Fock branch is mirrored")
    if H2_given is None:
        if kwargs.get("H2_fname", None) is not None:
            H2 = slater.load_H2(kwargs["H2_fname"])                      # spinless.py:446-449
        else:
            H2, blocks = _embHam2e(lattice, basis, vcor, local, **kwargs)
    else:
        H2 = H2_given
    kw1 = {k: v for k, v in kwargs.items() if k != "last_aabb"}
    H1, ovlp = _embHam1e(lattice, basis, vcor, mu, H2, blocks, **kw1)
    H0 = lattice.getH0() + kwargs.get("H0_add", 0.0)
    return Integral(nb, True, False, H0, {"cd": H1}, {"ccdd": H2}, ovlp=ovlp), None


embHam = get_emb_Ham


# ---------------------------------------------------------------------------------------------------------
# energy side of the GSO iteration (spinless.py:754-848, 948-1035)
# ---------------------------------------------------------------------------------------------------------
def _so_idx(idx, n):
    """orbital indices -> (alpha, beta) spin-orbital indices of an n-orbital block (spinless_helper.py:247-259)"""
    return [int(i) for i in idx], [int(i) + n for i in idx]


def transformResults(GRhoEmb, E, lattice, basis, ImpHam, H1e, mu, fit_ghf=False, **kwargs):
    """spinless.py:754-848: impurity block of the generalised density matrix, fragment energy, electron number
    (small host matrices, as in slater.transformResults).  E2 = E - <H1> - H0 of the solver is kept; E1 is
    re-evaluated with the chemical potentials added back, half of JK_core removed and the impurity weights
    applied."""
    if fit_ghf:
        raise UnsupportedBranch("fit_ghf (several bases) belongs to the correlation-potential fitting")
    basis = np.asarray(basis)
    ncells, nso, nbasis = basis.shape
    nao = nso // 2
    imp_a, imp_b = _so_idx(lattice.imp_idx, nao)
    GRhoEmb = np.asarray(GRhoEmb)
    if GRhoEmb.ndim == 3:                        # density matrices of UHF-type solvers (l.785-791)
        if GRhoEmb.shape[0] == 1:
            GRhoEmb = GRhoEmb[0]
        elif GRhoEmb.shape[0] == 2:
            GRhoEmb = GRhoEmb.sum(axis=0)
        else:
            raise ValueError
    GRhoImp = basis[0].dot(GRhoEmb).dot(basis[0].conj().T)
    nelec = GRhoImp[imp_a, imp_a].sum() - GRhoImp[imp_b, imp_b].sum() + len(imp_b)
    if E is None:
        return GRhoImp, None, nelec
    last_dmu = kwargs["last_dmu"]
    basis_Ra, basis_Rb = separate_basis(basis)
    H1 = np.asarray(ImpHam.H1["cd"][0])
    E2 = E - np.sum(H1 * GRhoEmb.T) - ImpHam.H0
    dmu_idx = kwargs.get("dmu_idx", None)
    dmu_idx = list(lattice.imp_idx) if dmu_idx is None else list(dmu_idx)
    emb_a, emb_b = _so_idx(kwargs.get("imp_idx", np.arange(lattice.nimp)), lattice.nimp)   # in the embedding basis
    # the chemical potentials the solver saw are put back: last_dmu on the chosen orbitals of cell 0, the global mu
    # on all orbitals of all cells, with opposite signs for the alpha and beta flavours (l.820-833)
    on_imp = np.zeros((2, nao, nao))
    on_imp[0][dmu_idx, dmu_idx] = last_dmu
    on_imp[1][dmu_idx, dmu_idx] = -last_dmu
    everywhere = np.asarray([mu * np.eye(nao), -mu * np.eye(nao)])
    H1_scaled = H1 + transform_imp(basis_Ra, basis_Rb, on_imp) + transform_local(basis_Ra, basis_Rb, everywhere)
    if lattice.JK_core is not None:
        H1_scaled = H1_scaled - 0.5 * np.asarray(lattice.JK_core)
    H1_scaled = slater.get_H1_scaled(np.array(H1_scaled, dtype=np.float64)[None], emb_a + emb_b)[0]
    return GRhoImp, np.sum(H1_scaled * GRhoEmb.T) + E2 + ImpHam.H0, nelec


def get_H_dmet(basis, lattice, ImpHam, last_dmu=None, mu=None, imp_idx=None, dmu_idx=None, add_vcor_to_E=False,
               vcor=None, compact=True, rdm1_emb=None, veff=None, rebuild_veff=False, E1=None, GV0=None, GV1=None,
               **kwargs):
    """spinless.py:948-1035: the GSO DMET Hamiltonian scaled by the number of impurity indices.  One-body part on
    the device (`transform_trans_inv_k`), two-body weights by `ldm_scale_eri`; the branches that rebuild J/K from a
    global density matrix through the lattice mean-field object stay with the reference."""
    if veff is not None or rebuild_veff:
        raise UnsupportedBranch("rebuilding JK_core from the global density needs the lattice mean-field object")
    basis = np.asarray(basis)
    nbasis = basis.shape[-1]
    basis_Ra, basis_Rb = separate_basis(basis)
    basis_k = lattice.R2k_basis(basis)
    basis_ka, basis_kb = separate_basis(basis_k)
    emb_a, emb_b = _so_idx(np.arange(lattice.nimp) if imp_idx is None else imp_idx, lattice.nimp)
    imp = emb_a + emb_b
    dev = get_device()
    blocks = slater._s4_blocks_dev(ImpHam.H2["ccdd"], nbasis)                # restore 4-fold symmetry (l.1026)
    if E1 is None:
        H1_scaled = transform_trans_inv_k(basis_ka, basis_kb, lattice.getH1(kspace=True))
        if lattice.JK_core is not None:                                      # double counting cf. the HF energy
            H1_scaled = H1_scaled + 0.5 * np.asarray(lattice.JK_core)
        if add_vcor_to_E:
            H1_scaled = H1_scaled + transform_local(basis_Ra, basis_Rb, vcor.get() * 0.5).real
            H1_scaled = H1_scaled - transform_imp(basis_Ra, basis_Rb, vcor.get() * 0.5).real
        if GV1 is not None:
            H1_scaled = H1_scaled - slater.transform_trans_inv_k(basis_k, GV1)
        H0 = lattice.getH0()
    else:                                                                    # E1 given: -(J - K) of the GHF matrix
        vj, vk = dev.jk_s4(blocks[0], dev.to_device(np.ascontiguousarray(rdm1_emb, dtype=np.float64), torch.float64))
        H1_scaled = -(vj - vk).cpu().numpy()
        H0 = float(np.real(E1 + lattice.getH0()))
    H1_scaled = slater.get_H1_scaled(np.array(H1_scaled, dtype=np.float64)[None], imp)
    if GV0 is not None:
        H0 = H0 - GV0 * 0.5
    H2_dev = slater.get_H2_scaled(torch.stack([blocks[0].clone()]), imp)
    if compact:
        H2_scaled = H2_dev.cpu().numpy()
    else:                                                                    # restore_Ham(.., 1) (l.1032-1034)
        H2_scaled = dev.restore_s1(H2_dev[0].contiguous(), nbasis).cpu().numpy()[None]
    return Integral(nbasis, True, False, H0, {"cd": H1_scaled}, {"ccdd": H2_scaled})
