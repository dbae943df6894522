"""Lattice Fourier transforms k <-> R with the reference's signatures (libdmet/system/fourier.py:39-177).

`R2k` / `k2R` / `FFTtoK` / `FFTtoT` run on the GPU as dense phase-matrix products (csrc/aux_kernels.cuh,
phase_transform_kernel): for the k-meshes of this path (Nk <= a few hundred) the direct sum streams the stack once
and is HBM-bound, and it matches scipy's FFT to rounding.  Inputs may be numpy arrays (result returned as numpy, as
the reference does) or torch CUDA tensors (result stays on the device).
"""
import warnings

import numpy as np
import torch

from .device import get_device
from .schedule import make_kpts_scaled, cell_vectors, round_to_FBZ, kpt_member, KPT_DIFF_TOL  # noqa: F401

IMAG_DISCARD_TOL = 1e-7   # libdmet/settings.py:4

_phase_cache = {}


def get_phase_R2k_scaled(kmesh, kpts_scaled=None):
    """exp(-i R.k), shape (ncells, nkpts) (fourier.py:112-121).  With the mesh's own k-points the exponent is an
    exact rational multiple of 2 pi, evaluated from integers."""
    kmesh = [int(x) for x in kmesh]
    R = cell_vectors(kmesh)
    if kpts_scaled is None:
        kidx = cell_vectors(kmesh)          # k = kidx / kmesh (mod 1) in fftfreq order
        num = np.zeros((len(R), len(kidx)), dtype=np.int64)
        lcm = int(np.lcm.reduce(kmesh))
        for d, n in enumerate(kmesh):
            num += np.outer(R[:, d], kidx[:, d]) * (lcm // n)
        num %= lcm
        ang = -2.0 * np.pi * num / lcm
    else:
        ang = -2.0 * np.pi * np.einsum("Ru,ku->Rk", R, np.asarray(kpts_scaled, dtype=float))
    return np.cos(ang) + 1j * np.sin(ang)


def get_phase_R2k(cell, kpts, kmesh=None):
    """fourier.py:112-121 for a (duck-typed) cell and absolute k-points."""
    scaled = cell.get_scaled_kpts(kpts)
    if kmesh is None:
        kmesh = [len(np.unique(scaled.round(8)[:, d])) for d in range(scaled.shape[-1])]   # fourier.py:83-89
    return get_phase_R2k_scaled(kmesh, scaled)


def get_phase_k2R(cell, kpts, kmesh=None):
    """exp(+ikR) / nkpts (fourier.py:123-127)."""
    return get_phase_R2k(cell, kpts, kmesh=kmesh).conj().T / len(kpts)


def _phase_dev(kmesh, forward):
    """device phase matrix W[out][in]: forward (R -> k): W[k][R] = exp(-ikR); backward: W[R][k] = exp(+ikR)."""
    key = (tuple(int(x) for x in kmesh), bool(forward))
    dev = get_device()
    if key not in _phase_cache:
        ph = get_phase_R2k_scaled(kmesh)               # (R, k)
        W = ph.T if forward else ph.conj()               # (k, R) / (R, k)
        _phase_cache[key] = dev.to_device(np.ascontiguousarray(W), torch.complex128)
    return _phase_cache[key]


def _as_dev(a):
    dev = get_device()
    if isinstance(a, torch.Tensor):
        return a.contiguous(), True
    a = np.asarray(a)
    if np.iscomplexobj(a):
        return dev.to_device(a.astype(np.complex128, copy=False), torch.complex128), False
    return dev.to_device(a.astype(np.float64, copy=False), torch.float64), False


def _transform(A, kmesh, forward, out_real, tol):
    """A: ((spin,) ncells, n, m).  Returns the same shape, complex (forward) or real (backward)."""
    nk = int(np.prod(kmesh))
    x, on_dev = _as_dev(A)
    assert x.shape[-3] == nk, "first (non-spin) dimension must be the number of cells / k-points"
    dev = get_device()
    x4 = x.reshape((-1,) + tuple(x.shape[-3:]))
    if len(kmesh) <= 3 and max(int(v) for v in kmesh) <= 8:
        # factorised, HBM-bound kernel on the mesh's own k-points
        out, imag = dev.lattice_dft(x4, kmesh, forward, out_real=out_real, scale=1.0 if forward else 1.0 / nk)
    else:
        W = _phase_dev(kmesh, forward)
        out, imag = dev.phase_transform(x4, W, out_real=out_real, scale=1.0 if forward else 1.0 / nk)
    out = out.reshape(tuple(x.shape))
    if out_real and imag is not None and imag > tol:
        warnings.warn("k2R: non-zero imaginary part: %15.8g" % imag)     # fourier.py:174-175
    if on_dev:
        return out
    return out.cpu().numpy()


def FFTtoK(A, kmesh):
    """fourier.py:160-166: (ncells, n, m) -> (nkpts, n, m) complex, no normalisation."""
    assert A.ndim == 3
    return _transform(A, kmesh, True, False, None)


def FFTtoT(B, kmesh, tol=IMAG_DISCARD_TOL):
    """fourier.py:168-177: (nkpts, n, m) -> (ncells, n, m) real part, 1/Nk normalisation, warns on imag > tol."""
    assert B.ndim == 3
    return _transform(B, kmesh, False, True, tol)


def R2k(dm_R, kmesh):
    """fourier.py:129-142."""
    if dm_R.ndim not in (3, 4):
        raise ValueError("unknown shape of dm_R: %s" % str(tuple(dm_R.shape)))
    return _transform(dm_R, kmesh, True, False, None)


def k2R(dm_k, kmesh, tol=IMAG_DISCARD_TOL):
    """fourier.py:144-158."""
    if dm_k.ndim not in (3, 4):
        raise ValueError("unknown shape of dm_k: %s" % str(tuple(dm_k.shape)))
    return _transform(dm_k, kmesh, False, True, tol)


def get_kmesh(cell, kpts):
    """fourier.py:83-89."""
    scaled_k = cell.get_scaled_kpts(kpts).round(8)
    return [len(np.unique(scaled_k[:, d])) for d in range(scaled_k.shape[-1])]
