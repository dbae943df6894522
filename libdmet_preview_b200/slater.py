"""Embedding basis and embedding Hamiltonian -- drop-in for `libdmet.routine.slater.get_emb_basis` (`embBasis`) and
`get_emb_Ham` (`embHam`), ab-initio interacting-bath Hartree-Fock branch plus the non-interacting-bath ERI
(slater.py:98-220, 320-370, 438-476, 478-523, 525-605, 639-643, 690-712).

`embHam` keeps the two-electron integrals on the GPU between the ERI build and the J/K contraction of the one-body
part; every contraction runs in libldm_b200.so.  `get_emb_basis` gathers the environment-impurity block of the
density matrix and calls LAPACK's SVD on the host, like the reference (north_star keeps the dense eigen/singular
value solves on the reference's own route).
"""
import warnings

import numpy as np
import scipy.linalg as la
import torch

from ._lib import UnsupportedBranch
from .device import get_device
from . import eri_transform
from .eri_transform import get_emb_eri, get_unit_eri
from .fourier import IMAG_DISCARD_TOL
from .integral import Integral, get_eri_format
from .make_basis import add_spin_dim, sandwich, _finish


# ---------------------------------------------------------------------------------------------------------
# get_emb_basis
# ---------------------------------------------------------------------------------------------------------
def _lowdin(s, tol=1e-14):
    """libdmet/lo/lowdin.py:83-92."""
    e, v = la.eigh(s)
    idx = e > tol
    if not idx.all():
        warnings.warn("_vec_lowdin has almost zero eigenvalues:\n%s" % e[~idx])
    return np.dot(v[:, idx] / np.sqrt(e[idx]), v[:, idx].conj().T)


def vec_lowdin(c, s=None):
    """libdmet/lo/lowdin.py:94-101 for an identity metric: c (c^T c)^{-1/2}."""
    m = c.conj().T.dot(c) if s is None else c.conj().T.dot(s).dot(c)
    return np.dot(c, _lowdin(m))


def get_emb_basis(lattice, rho=None, local=True, kind='svd', **kwargs):
    """slater.py:98-115."""
    if rho is None:
        rho = lattice.rdm1_lo_R
    if not local:
        raise UnsupportedBranch("non-local (particle-hole symmetric model) bath is outside the ab-initio path")
    rho = rho.cpu().numpy() if isinstance(rho, torch.Tensor) else np.asarray(rho)
    if kind == 'svd':
        return _get_emb_basis_svd(lattice, rho.real, **kwargs)
    elif kind == 'eig':
        return _get_emb_basis_eig(lattice, rho.real, **kwargs)
    else:
        raise ValueError("get_emb_basis: Unknown kind %s" % kind)


embBasis = get_emb_basis


def _get_emb_basis_svd(lattice, rdm1, **kwargs):
    """slater.py:117-220.  The bath orbitals are the left singular vectors of the (environment x bath-generating
    impurity orbitals) block of the density matrix with singular value >= tol_bath (or the first `nbath`).  With a
    valence bath only the valence orbitals generate bath states; rows of the bath on the virtual impurity orbitals
    are then zeroed and the bath is Loewdin re-orthonormalised (`orth`).  Result: (spin, ncells, nlo, nimp + nbath),
    identity on the impurity rows, bath on the environment rows."""
    opt = dict(imp_idx=lattice.imp_idx, val_idx=lattice.val_idx, valence_bath=True, orth=True, tol_bath=1e-9,
               nbath=None, localize_bath=None)
    opt.update({k: v for k, v in kwargs.items() if k in opt})
    if opt["localize_bath"] is not None:
        raise UnsupportedBranch("bath localisation is only defined for model Hamiltonians in the reference")
    imp_idx, val_idx = list(opt["imp_idx"]), list(opt["val_idx"])
    ncells, nlo = int(lattice.ncells), int(lattice.nscsites)
    ntot = ncells * nlo
    gen_idx = val_idx if opt["valence_bath"] else imp_idx        # orbitals whose entanglement defines the bath
    is_gen = np.zeros(ntot, dtype=bool)
    is_gen[gen_idx] = True
    is_imp = np.zeros(ntot, dtype=bool)
    is_imp[imp_idx] = True
    env_idx = np.flatnonzero(~is_gen)
    nimp = len(imp_idx)

    rdm1 = np.asarray(rdm1)
    rdm1 = rdm1[None] if rdm1.ndim == 3 else rdm1
    assert rdm1.shape[-3:] == (ncells, nlo, nlo)
    spin = rdm1.shape[0]
    # the reference takes the expanded-matrix route as soon as the LAST orbital of cell 0 generates bath states
    # (">= nlo - 1", slater.py:167); the route only changes the cap on the number of bath orbitals
    if max(gen_idx) >= nlo - 1:
        coupling = lattice.expand(rdm1)[:, env_idx][:, :, gen_idx]
        nbath_cap = len(gen_idx)
    else:
        coupling = rdm1.reshape(spin, ntot, nlo)[:, env_idx][:, :, gen_idx]
        nbath_cap = nlo

    basis = np.zeros((spin, ntot, 2 * nimp))
    for s in range(spin):
        u, sigma, _ = la.svd(coupling[s], full_matrices=False)          # host LAPACK, as in the reference
        nb = int(np.count_nonzero(sigma >= opt["tol_bath"])) if opt["nbath"] is None else int(opt["nbath"])
        if np.any(np.abs(sigma[:nb]) < opt["tol_bath"]):
            warnings.warn("Zero singular value exists, \nthis may cause numerical instability.")
        bath = u[:, :nb]
        if nb > 0 and opt["orth"]:
            bath[is_imp[env_idx]] = 0.0
            bath = vec_lowdin(bath)
        basis[s, imp_idx, :nimp] = np.eye(nimp)
        basis[s, env_idx, nimp:nimp + nb] = bath
        nbath_cap = min(nbath_cap, nb)
    return basis[:, :, :nimp + nbath_cap].reshape(spin, ncells, nlo, nimp + nbath_cap)


def _get_emb_basis_eig(lattice, rdm1, **kwargs):
    """slater.py:224-318 (fractionally occupied case).  The bath orbitals are the eigenvectors of the
    environment-environment block of the density matrix whose occupation is neither 0 nor 1 (within `tol_bath`);
    everything else -- which orbitals generate the bath, zeroing of the virtual impurity rows + Loewdin step, layout
    of the result -- is as in the SVD construction.  Host LAPACK (`eigh`), like the reference."""
    opt = dict(imp_idx=lattice.imp_idx, val_idx=lattice.val_idx, valence_bath=True, orth=True, tol_bath=1e-9,
               localize_bath=None)
    opt.update({k: v for k, v in kwargs.items() if k in opt})
    if opt["localize_bath"] is not None:
        raise UnsupportedBranch("bath localisation is only defined for model Hamiltonians in the reference")
    imp_idx, val_idx = list(opt["imp_idx"]), list(opt["val_idx"])
    ncells, nlo = int(lattice.ncells), int(lattice.nscsites)
    ntot = ncells * nlo
    gen_idx = val_idx if opt["valence_bath"] else imp_idx
    is_gen = np.zeros(ntot, dtype=bool)
    is_gen[gen_idx] = True
    is_imp = np.zeros(ntot, dtype=bool)
    is_imp[imp_idx] = True
    env_idx = np.flatnonzero(~is_gen)
    nimp = len(imp_idx)

    rdm1 = np.asarray(rdm1)
    rdm1 = rdm1[None] if rdm1.ndim == 3 else rdm1
    assert rdm1.shape[-3:] == (ncells, nlo, nlo)
    spin = rdm1.shape[0]
    env_env = lattice.expand(rdm1)[:, env_idx][:, :, env_idx]
    baths = []
    for s in range(spin):
        occ, vec = la.eigh(env_env[s])
        fractional = (np.abs(occ) > opt["tol_bath"]) & (np.abs(1.0 - occ) > opt["tol_bath"])
        baths.append(vec[:, fractional])
    nb = baths[0].shape[1]
    if any(b.shape[1] != nb for b in baths):       # the reference stacks the spin channels into one array (l.286)
        raise ValueError("the spin channels have different numbers of bath orbitals: %s"
                         % [b.shape[1] for b in baths])
    basis = np.zeros((spin, ntot, nimp + nb))
    for s in range(spin):
        bath = baths[s]
        if nb > 0 and opt["orth"]:
            bath[is_imp[env_idx]] = 0.0
            bath = vec_lowdin(bath)
        basis[s, imp_idx, :nimp] = np.eye(nimp)
        basis[s, env_idx, nimp:] = bath
    return basis.reshape(spin, ncells, nlo, nimp + nb)


# ---------------------------------------------------------------------------------------------------------
# one-body pieces on the device
# ---------------------------------------------------------------------------------------------------------
def _zdev(a):
    dev = get_device()
    if isinstance(a, torch.Tensor):
        return (a if a.dtype == torch.complex128 else a.to(torch.complex128)).contiguous()
    return dev.to_device(np.asarray(a).astype(np.complex128, copy=False), torch.complex128)


class _BasisK(object):
    """basis_k (spin, nkpts, nlo, neo) on the device plus its k-contiguous transpose."""

    def __init__(self, basis_k):
        dev = get_device()
        self.bk = _zdev(basis_k)
        self.spin, self.nk, self.nlo, self.neo = self.bk.shape
        self.bkT = dev.ztranspose(self.bk.reshape(-1, self.nlo, self.neo))      # (spin*nk, neo, nlo)


def transform_trans_inv_k_dev(bk, s, H_k_dev, h_spin_index):
    """Re[ sum_k B_k^dagger H_k B_k ] / nkpts for spin s  (slater_helper.py:37-50) -> (neo, neo) device tensor."""
    dev = get_device()
    nk, nlo, neo = bk.nk, bk.nlo, bk.neo
    a_index = s * nk + np.arange(nk)
    h_index = h_spin_index * nk + np.arange(nk)
    # V^T[k][n][l'] = sum_l B_k^T[n][l] H_k[l'][l]  ;  R[k][m][n] = sum_l' conj(B_k^T[m][l']) V^T[k][n][l']
    VT = dev.empty((nk, neo, nlo), torch.complex128)
    segs = np.zeros((nk, 4), dtype=np.int32)
    segs[:, 0] = a_index
    segs[:, 1] = h_index
    dev.zgemm_tn(bk.bkT, H_k_dev, segs, VT, c_off=np.arange(nk, dtype=np.int64) * neo * nlo, s_outer=nlo,
                 nbatch=nk, nseg=1)
    R = dev.empty((nk, neo, neo), torch.complex128)
    segs2 = np.zeros((nk, 4), dtype=np.int32)
    segs2[:, 0] = a_index
    segs2[:, 1] = np.arange(nk)
    segs2[:, 2] = 1
    dev.zgemm_tn(bk.bkT, VT, segs2, R, c_off=np.arange(nk, dtype=np.int64) * neo * neo, s_outer=neo, nbatch=nk,
                 nseg=1)
    res, imag = dev.ksum_real(R, scale=1.0 / float(nk))
    if imag > IMAG_DISCARD_TOL:
        warnings.warn("transform_trans_inv_k: has imag part %s" % imag)
    return res


def transform_h1_dev(H1_k, bk):
    """slater.py:690-697 (`transform_h1` / `foldRho_k`) -> (spin, neo, neo) float64 device tensor."""
    H = _zdev(H1_k)
    if H.dim() == 3:
        H = H[None]
    hs = H.shape[0]
    H3 = H.reshape(-1, H.shape[-2], H.shape[-1])
    return torch.stack([transform_trans_inv_k_dev(bk, s, H3, min(s, hs - 1)) for s in range(bk.spin)])


def transform_h1(H1_k, basis_k):
    """slater.py:690-697, numpy in / numpy out."""
    return transform_h1_dev(H1_k, _BasisK(basis_k)).cpu().numpy()


foldRho_k = transform_h1


def transform_trans_inv_k(basis_k, H_k):
    """slater_helper.py:37-50, numpy in / numpy out."""
    bk = _BasisK(np.asarray(basis_k)[None])
    return transform_trans_inv_k_dev(bk, 0, _zdev(np.asarray(H_k)), 0).cpu().numpy()


def _s4_blocks_dev(H2, norb):
    """any accepted H2 layout -> list of (npair, npair) device tensors (s4)."""
    dev = get_device()
    if isinstance(H2, torch.Tensor) and H2.dim() == 3:
        return [H2[i] for i in range(H2.shape[0])]
    H2 = H2.cpu().numpy() if isinstance(H2, torch.Tensor) else np.asarray(H2)
    fmt, spin_dim = get_eri_format(H2, norb)
    if spin_dim == 0:
        H2 = H2[None]
    npair = norb * (norb + 1) // 2
    idx = np.tril_indices(norb)
    out = []
    for blk in H2:
        if fmt == 's4':
            e4 = blk.reshape(npair, npair)
        elif fmt == 's1':
            e4 = blk.reshape(norb, norb, norb, norb)[idx[0], idx[1]][:, idx[0], idx[1]]
        else:   # s8
            e4 = np.zeros((npair, npair))
            t = np.tril_indices(npair)
            e4[t] = blk.reshape(-1)
            e4[(t[1], t[0])] = blk.reshape(-1)
        out.append(dev.to_device(np.ascontiguousarray(e4), torch.float64))
    return out


def get_veff_dev(rdm1_emb, eri4_blocks):
    """HF effective potential from embedding ERI and density (slater.py:478-523 -> solver/scf.py:255-352).
    rdm1_emb: (spin, n, n) device; eri4_blocks: 1 block (restricted / UHF with one ERI) or 3 blocks aa, bb, ab
    (the order embHam uses, slater.py:461-462).  Returns (spin, n, n) device."""
    dev = get_device()
    spin = rdm1_emb.shape[0]
    dm = rdm1_emb.contiguous()
    # the density matrices are symmetric (the reference passes hermi=1 throughout) and the restricted / aa / bb
    # blocks are symmetric matrices: those calls read only the lower triangle of the block
    if spin == 1:
        vj, vk = dev.jk_s4(eri4_blocks[0], dm[0], symmetric=True)
        return (vj - vk * 0.5)[None]                                    # scf.py:347-348
    if len(eri4_blocks) == 1:                                            # UHF with a spin-free ERI (scf.py:303-309)
        vj0, vk0 = dev.jk_s4(eri4_blocks[0], dm[0], symmetric=True)
        vj1, vk1 = dev.jk_s4(eri4_blocks[0], dm[1], symmetric=True)
        return torch.stack([vj0 + vj1 - vk0, vj0 + vj1 - vk1])
    assert len(eri4_blocks) == 3 and spin == 2                          # UIHF (scf.py:310-331)
    vj00, vk00 = dev.jk_s4(eri4_blocks[0], dm[0], symmetric=True)
    vj11, vk11 = dev.jk_s4(eri4_blocks[1], dm[1], symmetric=True)
    vj01, _ = dev.jk_s4(eri4_blocks[2], dm[1], with_k=False)            # J on alpha from beta density
    eri_ba = eri4_blocks[2].t().contiguous()
    vj10, _ = dev.jk_s4(eri_ba, dm[0], with_k=False)                    # J on beta from alpha density
    return torch.stack([vj00 + vj01 - vk00, vj11 + vj10 - vk11])        # scf.py:350 with vj=((00,11),(01,10))


def get_veff(rdm1, eri, hyb=1.0):
    """slater.py:478-523 (HF branch), numpy in / numpy out."""
    if hyb != 1.0:
        raise UnsupportedBranch("DFT / hybrid branches are outside the hot path")
    rdm1 = np.asarray(rdm1, dtype=np.double)
    if rdm1.ndim == 2:
        rdm1 = rdm1[None]
    dev = get_device()
    n = rdm1.shape[-1]
    return get_veff_dev(dev.to_device(rdm1, torch.float64), _s4_blocks_dev(eri, n)).cpu().numpy()


def unit2emb(H2_unit, neo):
    """slater_helper.py:494-517 (ndarray branch): the impurity-cell ERI zero-padded to `neo` embedding orbitals in
    the same s1 / s4 / s8 layout."""
    H2_unit = np.asarray(H2_unit)
    npair = neo * (neo + 1) // 2
    tail = {5: (neo,) * 4, 3: (npair, npair), 2: (npair * (npair + 1) // 2,)}.get(H2_unit.ndim)
    if tail is None:
        raise ValueError
    H2_emb = np.zeros((H2_unit.shape[0],) + tail)
    H2_emb[tuple(slice(0, n) for n in H2_unit.shape)] = H2_unit
    return H2_emb


# ---------------------------------------------------------------------------------------------------------
# get_emb_Ham
# ---------------------------------------------------------------------------------------------------------
def _embHam2e(lattice, basis, vcor, local, int_bath=True, last_aabb=True, **kwargs):
    """slater.py:372-476, ab-initio branch (438-472).  Returns (H2 numpy in the requested symmetry,
    s4 device blocks for the J/K step)."""
    if getattr(lattice, "is_model", False):
        raise UnsupportedBranch("model Hamiltonians are outside the ab-initio hot path")
    nbasis = basis.shape[-1]
    eri_symmetry = lattice.eri_symmetry
    cell, mydf, C_ao_lo = lattice.cell, lattice.df, lattice.C_ao_lo
    common = dict(kscaled_center=kwargs.get("kscaled_center", None), max_memory=kwargs.get("max_memory", None),
                  swap_idx=kwargs.get("swap_idx", None), t_reversal_symm=kwargs.get("t_reversal_symm", True),
                  incore=kwargs.get("incore", True), fout=kwargs.get("fout", "H2.h5"),
                  use_mpi=kwargs.get("use_mpi", False))
    for k in ("source", "group", "kl_group", "stats"):
        if k in kwargs:
            common[k] = kwargs[k]
    if int_bath:
        # build once in s4 on the device, keep it there for J/K, re-lay out for the caller
        eri4 = get_emb_eri(cell, mydf, C_ao_lo=C_ao_lo, basis=basis, symmetry=4, return_device=True, **common)
        order = [0, 2, 1] if (last_aabb and eri4.shape[0] == 3) else list(range(eri4.shape[0]))   # l.461-462
        blocks = [eri4[i] for i in order]
        dev = get_device()
        if eri_symmetry == 4:
            H2 = dev.to_host(eri4 if order == list(range(eri4.shape[0])) else torch.stack(blocks))
        elif eri_symmetry == 1:
            H2 = np.stack([dev.to_host(dev.restore_s1(b, nbasis)) for b in blocks])
        elif eri_symmetry == 8:
            assert len(blocks) == 1
            H2 = dev.to_host(dev.restore_s8(blocks[0], nbasis))[None]
        else:
            raise ValueError("unknown eri_symmetry %s" % eri_symmetry)
        return H2, blocks
    H2 = get_unit_eri(cell, mydf, C_ao_lo=C_ao_lo, symmetry=eri_symmetry, **common)
    if last_aabb and H2.shape[0] == 3:
        H2 = H2[[0, 2, 1]]
    H2 = unit2emb(H2, nbasis)                                            # l.472
    return H2, None


def _embHam1e(lattice, basis, vcor, H2_emb, eri4_blocks, int_bath=True, add_vcor=False, **kwargs):
    """slater.py:525-688, interacting-bath HF branch (590-605, 639-643).  Side effect: lattice.JK_core."""
    if not int_bath:
        raise UnsupportedBranch("the non-interacting-bath one-body branch is outside the hot path "
                                  "(its ERI is available through get_unit_eri / unit2emb)")
    for flag in ("dft", "qsgw"):
        if kwargs.get(flag, False):
            raise UnsupportedBranch("%s branch is outside the hot path" % flag)
    spin = basis.shape[0]
    nbasis = basis.shape[-1]
    bk = _BasisK(lattice.R2k_basis(basis))                               # l.533
    hcore_emb = transform_h1_dev(lattice.hcore_lo_k, bk)                 # l.542
    ovlp_emb = transform_h1_dev(lattice.ovlp_lo_k, bk)                   # l.544
    rdm1_emb = transform_h1_dev(lattice.rdm1_lo_k, bk)                   # l.560 (foldRho_k)
    fock_k = _zdev(lattice.hcore_lo_k) + _zdev(lattice.vhf_lo_k)         # l.592
    H1 = transform_h1_dev(fock_k, bk)                                    # l.597
    if eri4_blocks is None:
        eri4_blocks = _s4_blocks_dev(H2_emb, nbasis)
    JK_emb = get_veff_dev(rdm1_emb, eri4_blocks)                         # l.600
    H1 = H1 - JK_emb                                                     # l.605
    JK_core = H1 - hcore_emb                                             # l.640
    H1 = H1.cpu().numpy()
    lattice.JK_core = JK_core.cpu().numpy()                              # l.643
    ovlp_emb = ovlp_emb.cpu().numpy()
    if ovlp_emb.ndim == 3 and ovlp_emb.shape[0] == 1:
        ovlp_emb = ovlp_emb[0]                                           # l.545-546
    if add_vcor:                                                         # l.676-687 (small host matrices)
        b = np.asarray(basis)
        for s in range(spin):
            v = np.asarray(vcor.get()[s])
            H1[s] += sum(b[s, i].T.dot(v).dot(b[s, i]) for i in range(b.shape[1]))
            if not kwargs.get("fitting", False):
                H1[s] -= b[s, 0].T.dot(v).dot(b[s, 0])
    return H1, ovlp_emb


def _check_supported(lattice, kwargs):
    """branches `get_emb_Ham` does not mirror are refused HERE, before any ERI work, so that `patch.install()` can
    hand the call to the reference without the embedding ERI having been built twice"""
    if getattr(lattice, "is_model", False) and kwargs.get("H2_given", None) is None \
            and kwargs.get("H2_fname", None) is None:
        raise UnsupportedBranch("model Hamiltonians are outside the ab-initio hot path")
    if not kwargs.get("int_bath", True):
        raise UnsupportedBranch("the non-interacting-bath one-body branch is outside the hot path "
                                "(its ERI is available through get_unit_eri / unit2emb)")
    for flag in ("dft", "qsgw"):
        if kwargs.get(flag, False):
            raise UnsupportedBranch("%s branch is outside the hot path" % flag)


def load_H2(fname):
    """the embedding ERI a previous run stored in an HDF5 file, dataset 'emb_eri' (slater.py:349-355)"""
    from . import h5lite
    with h5lite.File(fname) as f:
        return np.asarray(f["emb_eri"][...])


def get_emb_Ham(lattice, basis, vcor, local=True, **kwargs):
    """slater.py:320-370.  Returns (Integral, None); H2 blocks come in the order aa, bb, ab (l.461-462).
    `H2_given` / `H2_fname` supply the two-electron integrals instead of building them (l.346-358)."""
    basis = np.asarray(basis)
    spin = basis.shape[0]
    nbasis = basis.shape[-1]
    _check_supported(lattice, kwargs)
    H2_given = kwargs.get("H2_given", None)
    blocks = None
    if H2_given is None:
        if kwargs.get("H2_fname", None) is not None:
            H2 = load_H2(kwargs["H2_fname"])
        else:
            H2, blocks = _embHam2e(lattice, basis, vcor, local, **kwargs)
    else:
        H2 = H2_given
    kw1 = {k: v for k, v in kwargs.items() if k != "last_aabb"}
    H1, ovlp_emb = _embHam1e(lattice, basis, vcor, H2, blocks, **kw1)
    H0 = lattice.getH0()
    if isinstance(H2, np.ndarray):
        H2 = {"ccdd": H2}
    ImpHam = Integral(nbasis, spin == 1, False, H0, {"cd": H1}, H2, ovlp=ovlp_emb)
    return ImpHam, None


embHam = get_emb_Ham


# ---------------------------------------------------------------------------------------------------------
# energy side of the iteration: scaled DMET Hamiltonian and result transformation
# ---------------------------------------------------------------------------------------------------------
def _env_idx(nbasis, imp_idx, env_idx):
    if env_idx is None:
        env_idx = np.setdiff1d(np.arange(nbasis), np.asarray(imp_idx, dtype=int))
    return np.asarray(env_idx, dtype=int)


def get_H1_scaled(H1, imp_idx, env_idx=None):
    """slater.py:1716-1732 (in place): every element is weighted by (number of impurity indices) / 2 --
    impurity-environment blocks halved, environment-environment block zeroed."""
    assert H1.ndim == 3
    member = np.zeros(H1.shape[-1])
    member[np.asarray(imp_idx, dtype=int)] = 1.0
    H1 *= 0.5 * (member[:, None] + member[None, :])
    return H1


def get_H2_scaled(H2, imp_idx, env_idx=None):
    """slater.py:1734-1778: scale every integral by (number of impurity indices) / 4; s4 (3-d) or s1 (5-d) layout.
    numpy arrays are scaled in place (through the device), torch CUDA tensors stay on the device."""
    dev = get_device()
    on_dev = isinstance(H2, torch.Tensor)
    if H2.ndim == 3:
        npair = H2.shape[-1]
        nbasis = int(np.sqrt(npair * 2))
        tri = np.tril_indices(nbasis)
        member = np.zeros(nbasis, dtype=np.int32)
        member[np.asarray(imp_idx, dtype=int)] = 1
        w = dev.to_device((member[tri[0]] + member[tri[1]]).astype(np.int32), torch.int32)
        sym = 4
    elif H2.ndim == 5:
        nbasis = H2.shape[-1]
        member = np.zeros(nbasis, dtype=np.int32)
        member[np.asarray(imp_idx, dtype=int)] = 1
        w = dev.to_device(member, torch.int32)
        sym = 1
    else:
        raise ValueError("Unknown H2 shape to scale: %s" % (str(tuple(H2.shape))))
    d = H2 if on_dev else dev.to_device(np.ascontiguousarray(H2), torch.float64)
    for blk in range(d.shape[0]):
        dev.scale_eri(d[blk], nbasis, sym, w)
    if on_dev:
        return d
    H2[...] = d.cpu().numpy()
    return H2


def get_H_dmet(basis, lattice, ImpHam, last_dmu, imp_idx=None, dmu_idx=None, add_vcor_to_E=False, vcor=None,
               compact=True, rdm1_emb=None, veff=None, rebuild_veff=False, E1=None, **kwargs):
    """slater.py:1957-2032: the DMET Hamiltonian scaled by the number of impurity indices, whose expectation value
    is the fragment energy.  The branches that rebuild J/K from a global density matrix through the lattice
    mean-field object (`veff`, `rebuild_veff`) stay with the reference."""
    if veff is not None or rebuild_veff:
        raise UnsupportedBranch("rebuilding JK_core from the global density needs the lattice mean-field object")
    basis = np.asarray(basis)
    spin = basis.shape[0]
    nbasis = basis.shape[-1]
    if imp_idx is None:
        imp_idx = list(range(lattice.nimp))
    imp_idx = np.asarray(imp_idx)
    env_idx = _env_idx(nbasis, imp_idx, None)
    dev = get_device()
    if E1 is None:
        bk = _BasisK(lattice.R2k_basis(basis))
        H1_scaled = transform_h1_dev(lattice.hcore_lo_k, bk).cpu().numpy()
        JK_core = lattice.JK_core if lattice.JK_core is not None else [0.0 for s in range(spin)]
        for s in range(spin):
            H1_scaled[s] += 0.5 * JK_core[s]
            if add_vcor_to_E:
                v = np.asarray(vcor.get()[s]) * 0.5
                H1_scaled[s] += sum(basis[s, i].T.dot(v).dot(basis[s, i]) for i in range(basis.shape[1]))
                H1_scaled[s] -= basis[s, 0].T.dot(v).dot(basis[s, 0])
        H1_scaled = get_H1_scaled(H1_scaled, imp_idx, env_idx)
        H0 = lattice.getH0()
    else:
        H1_scaled = (-1.0 / spin) * get_veff(rdm1_emb, ImpHam.H2["ccdd"], hyb=1.0)
        H1_scaled = get_H1_scaled(H1_scaled, imp_idx, env_idx)
        H0 = (E1 + lattice.getH0()).real
    blocks = _s4_blocks_dev(ImpHam.H2["ccdd"], nbasis)                       # restore 4-fold symmetry (l.2019-2022)
    H2_dev = torch.stack([b.clone() for b in blocks])
    H2_dev = get_H2_scaled(H2_dev, imp_idx, env_idx)
    if compact:
        H2_scaled = H2_dev.cpu().numpy()
    else:                                                                    # restore_Ham(ImpHam_dmet, 1) (l.2029-2031)
        H2_scaled = np.stack([dev.restore_s1(H2_dev[i].contiguous(), nbasis).cpu().numpy()
                              for i in range(H2_dev.shape[0])])
    return Integral(nbasis, spin == 1, False, H0, {"cd": H1_scaled}, {"ccdd": H2_scaled})


def transformResults(rhoEmb, E, basis, ImpHam, H1e=None, **kwargs):
    """slater.py:1780-1840: impurity block of the solver's density matrix, fragment energy and electron number
    (small host matrices).  The two-body part of the solver energy, E2 = E - <H1> - H0, is kept; the one-body part is
    re-evaluated with the chemical-potential shift and half of JK_core removed and the impurity weights applied."""
    rhoEmb, basis = np.asarray(rhoEmb), np.asarray(basis)
    spin, nscsites, nbasis = rhoEmb.shape[0], basis.shape[2], basis.shape[-1]
    lattice = kwargs.get("lattice", None)
    default_imp = range(lattice.nimp) if lattice is not None else np.arange(nscsites)
    imp_idx = np.asarray(kwargs.get("imp_idx", default_imp))
    if np.any(imp_idx >= nscsites):
        warnings.warn("imp_idx is out of the first cell... imp_idx:\n%s" % imp_idx)
    per_spin = 2.0 / spin
    nelec = per_spin * float(sum(rhoEmb[s, imp_idx, imp_idx].sum() for s in range(spin)))
    rhoImp = rhoEmb[:, imp_idx][:, :, imp_idx]
    if E is None:
        return rhoImp, None, nelec
    dmu_idx = kwargs.get("dmu_idx", None)
    dmu_idx = list(range(nscsites)) if dmu_idx is None else dmu_idx
    H1 = ImpHam.H1["cd"]
    E2 = E - per_spin * np.einsum("spq,sqp", H1, rhoEmb) - ImpHam.H0
    shift = np.zeros((nscsites, nscsites))
    shift[dmu_idx, dmu_idx] = -kwargs["last_dmu"]
    H1_scaled = np.array(H1, copy=True)
    for s in range(spin):
        H1_scaled[s] -= basis[s, 0].T.dot(shift).dot(basis[s, 0])          # chemical potential acts on cell 0
        if lattice.JK_core is not None:
            H1_scaled[s] -= 0.5 * lattice.JK_core[s]
    H1_scaled = get_H1_scaled(H1_scaled, imp_idx)
    E1 = per_spin * np.einsum("spq,sqp", H1_scaled, rhoEmb)
    return rhoImp, E1 + E2 + lattice.getH0(), nelec


# ---------------------------------------------------------------------------------------------------------
# global density matrix by democratic partitioning (slater_helper.py:183-283)
# ---------------------------------------------------------------------------------------------------------
def get_rho_glob_R(basis, lattice, rho_emb, symmetric=True, compact=True, sign=None):
    """Global one-particle density matrix in the LO basis, stripe shape (spin, ncells, nlo, nlo), averaged
    democratically over the impurity problems of all cells (slater_helper.py:183-270): every cell R carries a
    translated copy of the embedding problem, its impurity-impurity block counts fully, impurity-environment blocks
    half, environment-environment blocks not at all.

    The reference builds, for each R, the full-lattice product C_R rho C_R[:nlo]^T and masks it.  With the impurity
    inside cell 0 the sum over R collapses to
        rho_R[I] = 1/2 [ (M C_0) rho C_{-I}^T  +  C_I rho (M C_0)^T ]            (M = diagonal impurity mask,
                                                                                  C_I = rows of cell I)
    which is ONE batched product on the device: operands [1/2 M C_0 | C_I] and [C_{-I} | 1/2 M C_0] around
    blockdiag(rho, rho); several fragments (lists of bases / lattices / density matrices) extend the contraction
    index.  The masked operands are laid out on the host (no arithmetic beyond the mask), the contraction runs in
    `zgemm_tn` (make_basis.sandwich)."""
    from collections.abc import Iterable
    if isinstance(lattice, Iterable):
        frags = list(zip(basis, lattice, rho_emb))
    else:
        frags = [(basis, lattice, rho_emb)]
    if sign is not None or not compact:
        raise UnsupportedBranch("full-shape / signed global density matrices (particle-hole fragments) are "
                                  "outside the ab-initio path")
    left, right, blocks = [], [], []
    spin = ncells = nlo = None
    for basis_f, lat_f, rho_f in frags:
        b = np.asarray(basis_f.cpu() if isinstance(basis_f, torch.Tensor) else basis_f, dtype=np.float64)
        b = b[None] if b.ndim == 3 else b
        if spin is None:
            spin, ncells, nlo = b.shape[:3]
        assert b.shape[:3] == (spin, ncells, nlo)
        rho = np.asarray(rho_f.cpu() if isinstance(rho_f, torch.Tensor) else rho_f, dtype=np.float64)
        rho = add_spin_dim(rho, spin, non_spin_dim=2)
        imp = np.asarray(lat_f.imp_idx, dtype=int)
        if imp.size and (imp.min() < 0 or imp.max() >= nlo):
            raise UnsupportedBranch("impurity orbitals outside the first cell")
        mask = np.zeros(nlo)
        mask[imp] = 0.5
        half = np.broadcast_to((mask[None, :, None] * b[:, 0])[:, None], b.shape)         # 1/2 M C_0 for every cell
        minus = [lat_f.subtract(0, i) for i in range(ncells)]
        left += [half, b]
        right += [b[:, minus], half]
        blocks += [rho, rho]
    A1 = np.ascontiguousarray(np.concatenate(left, axis=-1)).reshape(spin * ncells, nlo, -1)
    A2T = np.ascontiguousarray(np.concatenate(right, axis=-1)).reshape(spin * ncells, nlo, -1)
    ntot = A1.shape[-1]
    H = np.zeros((spin, ntot, ntot))
    off = 0
    for r in blocks:
        n = r.shape[-1]
        H[:, off:off + n, off:off + n] = r
        off += n
    out = sandwich(_zdev(A1), False, _zdev(A2T), False, _zdev(H), h_index=np.repeat(np.arange(spin), ncells),
                   a_index=np.arange(spin * ncells))
    return _finish(out, (spin, ncells, nlo, nlo), np.float64, False)


def get_rho_glob_k(basis, lattice, rho_emb, symmetric=True, compact=True, sign=None):
    """slater_helper.py:272-283: the same matrix in k space (R2k on the device)."""
    from collections.abc import Iterable
    rho_R = get_rho_glob_R(basis, lattice, rho_emb, symmetric=symmetric, compact=compact, sign=sign)
    lat0 = lattice[0] if isinstance(lattice, Iterable) else lattice
    return lat0.R2k(rho_R)
