"""Oracle: AO->LO one-body transforms.  Restates libdmet/basis_transform/make_basis.py:524-558,923-962 and
libdmet/utils/misc.py:43-86.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from functools import reduce
import numpy as np


def mdot(*args):
    """misc.py:43-47."""
    return reduce(np.dot, args)


def kdot(a, b):
    """misc.py:49-59."""
    ka, s1_a, s2_a = a.shape
    kb, s1_b, s2_b = b.shape
    assert ka == kb
    res = np.zeros((ka, s1_a, s2_b), dtype=np.result_type(a.dtype, b.dtype))
    for k in range(ka):
        np.dot(a[k], b[k], out=res[k])
    return res


def get_spin_dim(arrays, non_spin_dim=3):
    """misc.py:61-74."""
    spin = 1
    for a in arrays:
        a = np.asarray(a)
        if a.ndim == non_spin_dim:
            continue
        elif a.ndim == non_spin_dim + 1:
            spin = max(spin, a.shape[0])
        else:
            raise ValueError
    return spin


def add_spin_dim(H, spin, non_spin_dim=3):
    """misc.py:76-86."""
    H = np.asarray(H)
    if H.ndim == non_spin_dim:
        H = H[None]
    assert H.ndim == (non_spin_dim + 1)
    if H.shape[0] < spin:
        H = np.asarray((H[0],) * spin)
    return H


def transform_h1_to_lo(h_ao_ao, C_ao_lo):
    """make_basis.py:524-558: h^{LO} = C^dagger h^{AO} C per k (and spin)."""
    h_ao_ao = np.asarray(h_ao_ao)
    C_ao_lo = np.asarray(C_ao_lo)
    nkpts = C_ao_lo.shape[-3]
    nlo = C_ao_lo.shape[-1]
    res_type = np.result_type(h_ao_ao.dtype, C_ao_lo.dtype)
    if h_ao_ao.ndim == 0:
        return np.ones((nkpts, nlo, nlo), dtype=res_type) * h_ao_ao
    elif h_ao_ao.ndim == 1:
        spin = len(h_ao_ao)
        h_lo_lo = np.ones((spin, nkpts, nlo, nlo), dtype=res_type)
        for s in range(spin):
            h_lo_lo[s] *= h_ao_ao[s]
        return h_lo_lo
    if C_ao_lo.ndim == 3 and h_ao_ao.ndim == 3:
        h_lo_lo = np.zeros((nkpts, nlo, nlo), dtype=res_type)
        for k in range(nkpts):
            h_lo_lo[k] = mdot(C_ao_lo[k].conj().T, h_ao_ao[k], C_ao_lo[k])
    else:
        spin = get_spin_dim((h_ao_ao, C_ao_lo))
        h_ao_ao = add_spin_dim(h_ao_ao, spin)
        C_ao_lo = add_spin_dim(C_ao_lo, spin)
        assert h_ao_ao.ndim == C_ao_lo.ndim
        h_lo_lo = np.zeros((spin, nkpts, nlo, nlo), dtype=res_type)
        for s in range(spin):
            for k in range(nkpts):
                h_lo_lo[s, k] = mdot(C_ao_lo[s, k].conj().T, h_ao_ao[s, k], C_ao_lo[s, k])
    return h_lo_lo


def multiply_basis(C_ao_lo, C_lo_eo):
    """make_basis.py:923-962: C_ao_eo = C_ao_lo . C_lo_eo per k (and spin)."""
    C_ao_lo = np.asarray(C_ao_lo)
    C_lo_eo = np.asarray(C_lo_eo)
    nkpts, nlo, neo = C_lo_eo.shape[-3:]
    nao = C_ao_lo.shape[-2]
    if C_ao_lo.ndim == 3 and C_lo_eo.ndim == 3:
        C_ao_eo = kdot(C_ao_lo, C_lo_eo)
    else:
        if C_ao_lo.ndim == 3 and C_lo_eo.ndim == 4:
            spin = C_lo_eo.shape[0]
            C_ao_lo = add_spin_dim(C_ao_lo, spin)
        elif C_ao_lo.ndim == 4 and C_lo_eo.ndim == 3:
            spin = C_ao_lo.shape[0]
            C_lo_eo = add_spin_dim(C_lo_eo, spin)
        elif C_ao_lo.ndim == 4 and C_lo_eo.ndim == 4:
            spin = max(C_ao_lo.shape[0], C_lo_eo.shape[0])
            C_ao_lo = add_spin_dim(C_ao_lo, spin)
            C_lo_eo = add_spin_dim(C_lo_eo, spin)
        else:
            raise ValueError("invalid shape for multiply_basis: C_ao_lo shape %s, C_lo_eo shape: %s"
                             % (C_ao_lo.shape, C_lo_eo.shape))
        C_ao_eo = np.zeros((spin, nkpts, nao, neo), dtype=np.result_type(C_ao_lo.dtype, C_lo_eo.dtype))
        for s in range(spin):
            C_ao_eo[s] = kdot(C_ao_lo[s], C_lo_eo[s])
    return C_ao_eo
