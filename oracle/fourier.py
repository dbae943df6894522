"""Oracle: lattice Fourier transforms and k-point helpers.  Restates libdmet/system/fourier.py:39-177 and
libdmet/system/lattice.py:44-56,304-351.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import itertools as it
import numpy as np
from scipy import fft as scifft

from . import pyscf_lib as lib
from .pyscf_lib import KPT_DIFF_TOL

IMAG_DISCARD_TOL = 1e-7  # libdmet/settings.py:4


def max_abs(x):
    """libdmet/utils/misc.py:34-41."""
    x = np.asarray(x)
    if x.size == 0:
        return 0.0
    if np.iscomplexobj(x):
        return np.abs(x).max()
    return max(np.max(x), abs(np.min(x)))


def get_R_vec_rel(kmesh):
    """Integer cell positions, C order (fourier.py:39-44 before the multiplication by lattice vectors)."""
    return lib.cartesian_prod([np.arange(x) for x in kmesh])


def make_kpts_scaled(kmesh):
    """fourier.py:46-53."""
    ks_each_axis = [scifft.fftfreq(kmesh[d], 1.0) for d in range(len(kmesh))]
    return lib.cartesian_prod(ks_each_axis)


def round_to_FBZ(kpts, tol=1e-10, wrap_around=True):
    """fourier.py:55-65."""
    kpts_round = kpts - np.floor(kpts)
    if wrap_around:
        kpts_round[kpts_round > (0.5 - tol)] -= 1.0
    else:
        kpts_round[kpts_round > (1.0 - tol)] = 0.0
    return kpts_round


def kpt_member(kpt, kpts, tol=KPT_DIFF_TOL):
    """fourier.py:73-81."""
    kpts = np.reshape(kpts, (len(kpts), kpt.size))
    dk = kpts - kpt.ravel()
    dk = np.linalg.norm(dk - np.round(dk), axis=-1)
    return np.where(dk < tol)[0]


def get_phase_R2k_scaled(kmesh, kpts_scaled):
    """exp(-i R.k), shape (ncells, nkpts) (fourier.py:112-121) written with scaled k-points and integer cell
    vectors: R_abs . k_abs = 2 pi R_rel . k_scaled."""
    R_rel = get_R_vec_rel(kmesh)
    return np.exp(-2.0j * np.pi * np.einsum("Ru,ku->Rk", R_rel, np.asarray(kpts_scaled)))


def FFTtoK(A, kmesh):
    """fourier.py:160-166."""
    return scifft.fftn(A.reshape(tuple(kmesh) + A.shape[-2:]),
                       axes=range(len(kmesh)), workers=-1).reshape(A.shape)


def FFTtoT(B, kmesh, tol=IMAG_DISCARD_TOL, warn=None):
    """fourier.py:168-177 (the warning is returned through `warn` if a list is given)."""
    A = scifft.ifftn(B.reshape(tuple(kmesh) + B.shape[-2:]),
                     axes=range(len(kmesh)), workers=-1).reshape(B.shape)
    if max_abs(A.imag) > tol and warn is not None:
        warn.append(max_abs(A.imag))
    return A.real


def R2k(dm_R, kmesh):
    """fourier.py:129-142."""
    if dm_R.ndim == 3:
        dm_k = FFTtoK(dm_R, kmesh)
    elif dm_R.ndim == 4:
        dm_k = np.zeros_like(dm_R, dtype=np.complex128)
        for s in range(dm_R.shape[0]):
            dm_k[s] = FFTtoK(dm_R[s], kmesh)
    else:
        raise ValueError("unknown shape of dm_R: %s" % str(dm_R.shape))
    return dm_k


def k2R(dm_k, kmesh, tol=IMAG_DISCARD_TOL, warn=None):
    """fourier.py:144-158."""
    if dm_k.ndim == 3:
        dm_R = FFTtoT(dm_k, kmesh, tol=tol, warn=warn)
    elif dm_k.ndim == 4:
        dm_R = np.zeros_like(dm_k)
        for s in range(dm_R.shape[0]):
            dm_R[s] = FFTtoT(dm_k[s], kmesh, tol=tol, warn=warn)
    else:
        raise ValueError("unknown shape of dm_k: %s" % str(dm_k.shape))
    return dm_R.real


class StripeLattice(object):
    """The part of libdmet/system/lattice.py:31-56,192-231,304-351,399-411 the path reads: cell grid, cell
    index arithmetic, expand/extract_stripe and the k2R/R2k method wrappers."""

    def __init__(self, kmesh, nscsites):
        self.kmesh = list(kmesh)
        self.csize = np.asarray(kmesh)
        self.ncells = int(np.prod(self.csize))
        self.nkpts = self.ncells
        self.nscsites = self.nao = int(nscsites)
        self.cells = get_R_vec_rel(kmesh)
        self.celldict = dict(zip(map(tuple, self.cells), range(self.ncells)))
        self.kpts_scaled = make_kpts_scaled(kmesh)
        self.phase_R2k = get_phase_R2k_scaled(kmesh, self.kpts_scaled)
        self.phase_k2R = self.phase_R2k.conj().T / self.nkpts

    def cell_idx2pos(self, idx):
        return self.cells[idx % self.ncells]

    def cell_pos2idx(self, pos):
        return self.celldict[tuple(pos % self.csize)]

    def add(self, i, j):
        return self.cell_pos2idx(self.cell_idx2pos(i) + self.cell_idx2pos(j))

    def subtract(self, i, j):
        return self.cell_pos2idx(self.cell_idx2pos(i) - self.cell_idx2pos(j))

    def k2R(self, A, tol=IMAG_DISCARD_TOL):
        return k2R(A, self.kmesh, tol=tol)

    def R2k(self, B):
        return R2k(B, self.kmesh)

    k2R_basis = k2R
    R2k_basis = R2k

    def expand(self, A, dense=False):
        """lattice.py:304-337."""
        assert A.shape[-3] == self.ncells
        nscsites = A.shape[-1]
        nsites = A.shape[-1] * A.shape[-3]
        if A.ndim == 3:
            bigA = np.zeros((nsites, nsites), dtype=A.dtype)
            rng = range(self.ncells) if dense else \
                [j for j in range(self.ncells) if not np.allclose(A[j], 0.0)]
            for i, j in it.product(rng, range(self.ncells)):
                idx = self.add(i, j)
                bigA[idx*nscsites:(idx+1)*nscsites, j*nscsites:(j+1)*nscsites] = A[i]
        elif A.ndim == 4:
            spin = A.shape[0]
            bigA = np.zeros((spin, nsites, nsites), dtype=A.dtype)
            rng = range(self.ncells) if dense else \
                [j for j in range(self.ncells) if not np.allclose(A[:, j], 0.0)]
            for i, j in it.product(rng, range(self.ncells)):
                idx = self.add(i, j)
                bigA[:, idx*nscsites:(idx+1)*nscsites, j*nscsites:(j+1)*nscsites] = A[:, i]
        else:
            raise ValueError("unknown shape of A, %s" % (A.shape,))
        return bigA

    def extract_stripe(self, A):
        """lattice.py:339-351."""
        ncells = self.ncells
        nscsites = A.shape[-1] // ncells
        if A.ndim == 2:
            return A.reshape((ncells, nscsites, ncells, nscsites))[:, :, 0]
        elif A.ndim == 3:
            spin = A.shape[0]
            return A.reshape((spin, ncells, nscsites, ncells, nscsites))[:, :, :, 0]
        raise ValueError("unknown shape of A, %s" % (A.shape,))
