"""Oracle: AO->LO one-body transforms, restated with numpy einsum.

What is restated (behaviour, not code): libdmet/basis_transform/make_basis.py:524-558 `transform_h1_to_lo`
(h_lo = C^dagger h C per k-point and spin, with spin broadcasting, scalar / per-spin-scalar shortcuts and numpy
result-type promotion), :923-962 `multiply_basis` (C_ao_eo = C_ao_lo . C_lo_eo per k-point, spin broadcast), and the
helpers libdmet/utils/misc.py:43-86 (`mdot`, `kdot`, `get_spin_dim`, `add_spin_dim`).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np


def mdot(*mats):
    """chained matrix product (misc.py:43-47)"""
    out = mats[0]
    for m in mats[1:]:
        out = np.dot(out, m)
    return out


def kdot(a, b):
    """per-k matrix product of two (nk, ., .) stacks (misc.py:49-59)"""
    assert a.shape[0] == b.shape[0]
    return np.einsum("kij,kjl->kil", a, b).astype(np.result_type(a.dtype, b.dtype), copy=False)


def get_spin_dim(arrays, non_spin_dim=3):
    """largest leading spin dimension among arrays that carry one (misc.py:61-74)"""
    dims = [1]
    for x in arrays:
        nd = np.ndim(x)
        if nd == non_spin_dim + 1:
            dims.append(np.shape(x)[0])
        elif nd != non_spin_dim:
            raise ValueError("array of rank %d is neither a k-stack nor a spin stack of k-stacks" % nd)
    return max(dims)


def add_spin_dim(H, spin, non_spin_dim=3):
    """give H a leading spin axis and replicate a single spin block up to `spin` (misc.py:76-86)"""
    H = np.asarray(H)
    H = H[None] if H.ndim == non_spin_dim else H
    assert H.ndim == non_spin_dim + 1
    return np.asarray([H[0]] * spin) if H.shape[0] < spin else H


def _spin_broadcast(x, C):
    """(x4, C4, squeeze): both with a spin axis of common length; squeeze tells the caller to drop it again"""
    squeeze = (np.ndim(x) == 3 and np.ndim(C) == 3)
    spin = get_spin_dim((x, C))
    return add_spin_dim(x, spin), add_spin_dim(C, spin), squeeze


def transform_h1_to_lo(h_ao_ao, C_ao_lo):
    """h_lo[s,k] = C[s,k]^dagger h[s,k] C[s,k]   (make_basis.py:524-558)"""
    h = np.asarray(h_ao_ao)
    C = np.asarray(C_ao_lo)
    nk, nlo = C.shape[-3], C.shape[-1]
    out_dtype = np.result_type(h.dtype, C.dtype)
    if h.ndim == 0:                              # a bare number stands for h = const (l.536-537)
        return np.full((nk, nlo, nlo), h, dtype=out_dtype)
    if h.ndim == 1:                              # one number per spin (l.538-543)
        return np.stack([np.full((nk, nlo, nlo), v, dtype=out_dtype) for v in h])
    h4, C4, squeeze = _spin_broadcast(h, C)
    res = np.einsum("skpm,skpq,skqn->skmn", C4.conj(), h4, C4).astype(out_dtype, copy=False)
    return res[0] if squeeze else res


def multiply_basis(C_ao_lo, C_lo_eo):
    """C_ao_eo[s,k] = C_ao_lo[s,k] . C_lo_eo[s,k]   (make_basis.py:923-962)"""
    A = np.asarray(C_ao_lo)
    B = np.asarray(C_lo_eo)
    if A.ndim not in (3, 4) or B.ndim not in (3, 4):
        raise ValueError("invalid shape for multiply_basis: C_ao_lo shape %s, C_lo_eo shape: %s" % (A.shape, B.shape))
    out_dtype = np.result_type(A.dtype, B.dtype)
    A4, B4, squeeze = _spin_broadcast(A, B)
    res = np.einsum("skpl,skln->skpn", A4, B4).astype(out_dtype, copy=False)
    return res[0] if squeeze else res
