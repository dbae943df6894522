"""Oracle: lattice Fourier transforms and k-point helpers.

Behaviour restated from libdmet/system/fourier.py:39-65 (cell vectors, fftfreq-ordered scaled k-points,
round_to_FBZ), :73-81 (kpt_member), :112-121 (R2k phase), :129-177 (R2k / k2R / FFTtoK / FFTtoT: scipy fftn / ifftn
over the mesh axes, no normalisation R->k, 1/Nk k->R, k2R keeps the real part) and from libdmet/system/lattice.py:
44-56 (cell grid, phases), 192-207 (cell index arithmetic), 304-351 (expand / extract_stripe), 399-411.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
from scipy import fft as scifft

from . import pyscf_lib as lib
from .pyscf_lib import KPT_DIFF_TOL

IMAG_DISCARD_TOL = 1e-7  # libdmet/settings.py:4


def max_abs(x):
    """largest magnitude of an array (libdmet/utils/misc.py:34-41)"""
    x = np.asarray(x)
    return float(np.abs(x).max()) if x.size else 0.0


def get_R_vec_rel(kmesh):
    """integer cell positions in C order (fourier.py:39-44 before multiplying by the lattice vectors)"""
    return lib.cartesian_prod([np.arange(n) for n in kmesh])


def make_kpts_scaled(kmesh):
    """scaled k-points in numpy.fft order (fourier.py:46-53)"""
    return lib.cartesian_prod([scifft.fftfreq(int(n), 1.0) for n in kmesh])


def round_to_FBZ(kpts, tol=1e-10, wrap_around=True):
    """fold fractional k-points into [-0.5, 0.5) (or [0, 1) without wrap-around)   (fourier.py:55-65)"""
    folded = np.asarray(kpts, dtype=float) - np.floor(kpts)
    if wrap_around:
        return np.where(folded > 0.5 - tol, folded - 1.0, folded)
    return np.where(folded > 1.0 - tol, 0.0, folded)


def kpt_member(kpt, kpts, tol=KPT_DIFF_TOL):
    """indices of kpts equal to kpt modulo reciprocal lattice vectors (fourier.py:73-81)"""
    delta = np.asarray(kpts).reshape(len(kpts), -1) - np.ravel(kpt)
    dist = np.linalg.norm(delta - np.round(delta), axis=1)
    return np.flatnonzero(dist < tol)


def get_phase_R2k_scaled(kmesh, kpts_scaled):
    """exp(-i R.k), shape (ncells, nkpts) (fourier.py:112-121); R.k = 2 pi R_rel . k_scaled"""
    return np.exp(-2.0j * np.pi * np.dot(get_R_vec_rel(kmesh), np.asarray(kpts_scaled).T))


def _mesh_view(A, kmesh):
    return A.reshape(tuple(kmesh) + A.shape[-2:]), tuple(range(len(kmesh)))


def FFTtoK(A, kmesh):
    """cells -> k-points, unnormalised (fourier.py:160-166)"""
    view, axes = _mesh_view(A, kmesh)
    return scifft.fftn(view, axes=axes).reshape(A.shape)


def FFTtoT(B, kmesh, tol=IMAG_DISCARD_TOL, warn=None):
    """k-points -> cells with 1/Nk, real part; the imaginary-part warning of fourier.py:174-175 is appended to
    `warn` when a list is passed (fourier.py:168-177)"""
    view, axes = _mesh_view(B, kmesh)
    A = scifft.ifftn(view, axes=axes).reshape(B.shape)
    worst = max_abs(A.imag)
    if warn is not None and worst > tol:
        warn.append(worst)
    return A.real


def _per_spin(fn, x, what):
    if x.ndim == 3:
        return fn(x)
    if x.ndim == 4:
        return np.stack([fn(xs) for xs in x])
    raise ValueError("unknown shape of %s: %s" % (what, str(x.shape)))


def R2k(dm_R, kmesh):
    """fourier.py:129-142"""
    return _per_spin(lambda a: FFTtoK(a, kmesh), dm_R, "dm_R")


def k2R(dm_k, kmesh, tol=IMAG_DISCARD_TOL, warn=None):
    """fourier.py:144-158"""
    return _per_spin(lambda b: FFTtoT(b, kmesh, tol=tol, warn=warn), dm_k, "dm_k").real


class StripeLattice(object):
    """The parts of the reference's Lattice the path reads: cell grid and index arithmetic, phases, k2R / R2k
    wrappers, expand / extract_stripe (lattice.py:31-56, 192-231, 304-351, 399-411)."""

    def __init__(self, kmesh, nscsites):
        self.kmesh = [int(n) for n in kmesh]
        self.csize = np.asarray(self.kmesh)
        self.ncells = self.nkpts = int(np.prod(self.csize))
        self.nscsites = self.nao = int(nscsites)
        self.cells = get_R_vec_rel(self.kmesh)
        self.celldict = {tuple(c): i for i, c in enumerate(self.cells)}
        self.kpts_scaled = make_kpts_scaled(self.kmesh)
        self.phase_R2k = get_phase_R2k_scaled(self.kmesh, self.kpts_scaled)
        self.phase_k2R = self.phase_R2k.conj().T / self.nkpts

    # cell index arithmetic
    def cell_idx2pos(self, idx):
        return self.cells[idx % self.ncells]

    def cell_pos2idx(self, pos):
        return self.celldict[tuple(np.mod(pos, self.csize))]

    def add(self, i, j):
        return self.cell_pos2idx(self.cells[i] + self.cells[j])

    def subtract(self, i, j):
        return self.cell_pos2idx(self.cells[i] - self.cells[j])

    # Fourier wrappers
    def k2R(self, A, tol=IMAG_DISCARD_TOL):
        return k2R(A, self.kmesh, tol=tol)

    def R2k(self, B):
        return R2k(B, self.kmesh)

    k2R_basis = k2R
    R2k_basis = R2k

    def expand(self, A, dense=False):
        """stripe (.., ncells, n, n) -> full translation-invariant (.., ncells*n, ncells*n): block (i+j, j) = A[i].
        Without `dense`, all-zero stripes are skipped like the reference does (lattice.py:304-337)."""
        assert A.shape[-3] == self.ncells
        n = A.shape[-1]
        lead = A.shape[:-3]
        stripes = A.reshape((-1, self.ncells, n, n))
        full = np.zeros((stripes.shape[0], self.ncells * n, self.ncells * n), dtype=A.dtype)
        for i in range(self.ncells):
            if not dense and np.allclose(stripes[:, i], 0.0):
                continue
            for j in range(self.ncells):
                r = self.add(i, j)
                full[:, r * n:(r + 1) * n, j * n:(j + 1) * n] = stripes[:, i]
        return full.reshape(lead + full.shape[1:])

    def extract_stripe(self, A):
        """first block column of a full matrix (lattice.py:339-351)"""
        n = A.shape[-1] // self.ncells
        if A.ndim not in (2, 3):
            raise ValueError("unknown shape of A, %s" % (A.shape,))
        blocks = A.reshape(A.shape[:-2] + (self.ncells, n, self.ncells, n))
        return blocks[..., :, :, 0, :]
