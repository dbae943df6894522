"""AO <-> LO one-body transforms with the reference's signatures (libdmet/basis_transform/make_basis.py:524-644,
923-962), computed as batched complex GEMMs on the FP64 tensor cores (csrc/zgemm_tn.cuh).

Every transform here is a per-(spin, k) sandwich  out = A1 . H . A2.  With the TN kernel it is two launches:
    V^T[c][n'] = sum_n  A2^T[c][n] H[n'][n]            (= (H A2)[n'][c])
    out[r][c]  = sum_n' A1[r][n'] V^T[c][n']
so only k-contiguous copies of the coefficient matrices are needed (one batched transpose).
"""
import numpy as np
import torch

from .device import get_device


def get_spin_dim(arrays, non_spin_dim=3):
    """libdmet/utils/misc.py:61-74."""
    spin = 1
    for a in arrays:
        nd = a.ndim
        if nd == non_spin_dim:
            continue
        elif nd == non_spin_dim + 1:
            spin = max(spin, a.shape[0])
        else:
            raise ValueError
    return spin


def add_spin_dim(H, spin, non_spin_dim=3):
    """libdmet/utils/misc.py:76-86 (numpy or torch)."""
    if isinstance(H, torch.Tensor):
        if H.dim() == non_spin_dim:
            H = H[None]
        assert H.dim() == non_spin_dim + 1
        if H.shape[0] < spin:
            H = H[:1].expand((spin,) + tuple(H.shape[1:]))
        return H
    H = np.asarray(H)
    if H.ndim == non_spin_dim:
        H = H[None]
    assert H.ndim == (non_spin_dim + 1)
    if H.shape[0] < spin:
        H = np.asarray((H[0],) * spin)
    return H


def _zdev(a):
    """-> contiguous complex128 device tensor; second value tells whether the input lived on the device."""
    dev = get_device()
    if isinstance(a, torch.Tensor):
        t = a if a.dtype == torch.complex128 else a.to(torch.complex128)
        return t.contiguous(), True
    return dev.to_device(np.asarray(a).astype(np.complex128, copy=False), torch.complex128), False


def sandwich(A1, conj1, A2T, conj2, H, h_index=None, a_index=None):
    """out[b] = op1(A1[a(b)]) . H[h(b)] . op2(A2T[a(b)])^T  for b in range(nb).

    A1: (na, r, n), A2T: (na, c, n), H: (nh, n, n) complex128 device tensors; h_index / a_index map the batch
    entry to the slice of H / of the coefficient arrays (spin broadcasting without copies)."""
    dev = get_device()
    nb = len(h_index)
    r, n = A1.shape[-2], A1.shape[-1]
    c = A2T.shape[-2]
    VT = dev.empty((nb, c, n), torch.complex128)
    segs = np.zeros((nb, 4), dtype=np.int32)
    segs[:, 0] = a_index
    segs[:, 1] = h_index
    segs[:, 2] = int(conj2)
    dev.zgemm_tn(A2T, H, segs, VT, c_off=np.arange(nb, dtype=np.int64) * c * n, s_outer=n, nbatch=nb, nseg=1)
    out = dev.empty((nb, r, c), torch.complex128)
    segs2 = np.zeros((nb, 4), dtype=np.int32)
    segs2[:, 0] = a_index
    segs2[:, 1] = np.arange(nb)
    segs2[:, 2] = int(conj1)
    dev.zgemm_tn(A1, VT, segs2, out, c_off=np.arange(nb, dtype=np.int64) * r * c, s_outer=c, nbatch=nb, nseg=1)
    return out


def _finish(out_dev, shape, res_type, on_dev):
    out_dev = out_dev.reshape(shape)
    if on_dev:
        return out_dev if np.issubdtype(res_type, np.complexfloating) else out_dev.real.contiguous()
    out = out_dev.cpu().numpy()
    if not np.issubdtype(res_type, np.complexfloating):
        out = np.ascontiguousarray(out.real)
    return out


def _np_dtype(a):
    if isinstance(a, torch.Tensor):
        return np.complex128 if a.is_complex() else np.float64
    return np.asarray(a).dtype


def _batch_maps(spin, nk, h_spin, c_spin):
    b = np.arange(spin * nk)
    s, k = b // nk, b % nk
    h_index = np.minimum(s, h_spin - 1) * nk + k
    a_index = np.minimum(s, c_spin - 1) * nk + k
    return h_index, a_index


def _prep(x, C_ao_lo):
    """common shape handling: returns (x4, C4, spin, nkpts, squeeze)"""
    squeeze = (C_ao_lo.ndim == 3 and x.ndim == 3)
    spin = get_spin_dim((x, C_ao_lo))
    x4 = x if x.ndim == 4 else x[None]
    C4 = C_ao_lo if C_ao_lo.ndim == 4 else C_ao_lo[None]
    return x4, C4, spin, C4.shape[1], squeeze


def transform_h1_to_lo(h_ao_ao, C_ao_lo):
    r"""make_basis.py:524-558:  h^{LO} = C^\dagger h^{AO} C  per k (and spin)."""
    if not isinstance(h_ao_ao, torch.Tensor):
        h_ao_ao = np.asarray(h_ao_ao)
    if not isinstance(C_ao_lo, torch.Tensor):
        C_ao_lo = np.asarray(C_ao_lo)
    nkpts = C_ao_lo.shape[-3]
    nlo = C_ao_lo.shape[-1]
    res_type = np.result_type(_np_dtype(h_ao_ao), _np_dtype(C_ao_lo))
    if h_ao_ao.ndim == 0:       # scalar shortcut (l.536-537)
        return np.ones((nkpts, nlo, nlo), dtype=res_type) * h_ao_ao
    elif h_ao_ao.ndim == 1:     # [0, 0] shortcut (l.538-543)
        spin = len(h_ao_ao)
        h_lo_lo = np.ones((spin, nkpts, nlo, nlo), dtype=res_type)
        for s in range(spin):
            h_lo_lo[s] *= h_ao_ao[s]
        return h_lo_lo
    h4, C4, spin, nk, squeeze = _prep(h_ao_ao, C_ao_lo)
    dev = get_device()
    hd, on1 = _zdev(h4)
    Cd, on2 = _zdev(C4)
    CT = dev.ztranspose(Cd.reshape(-1, Cd.shape[-2], nlo))          # (cs*nk, nlo, nao)
    h_index, a_index = _batch_maps(spin, nk, h4.shape[0], C4.shape[0])
    out = sandwich(CT, True, CT, False, hd.reshape(-1, hd.shape[-2], hd.shape[-1]), h_index, a_index)
    shape = (nk, nlo, nlo) if squeeze else (spin, nk, nlo, nlo)
    return _finish(out, shape, res_type, on1 and on2)


def _c_inverse(Cd, Sd, nk):
    """C^{-1} = C^dagger S per (spin, k): (cs*nk, nlo, nao)."""
    dev = get_device()
    cs = Cd.shape[0]
    nao, nlo = Cd.shape[-2], Cd.shape[-1]
    CT = dev.ztranspose(Cd.reshape(-1, nao, nlo))
    nb = cs * nk
    segs = np.zeros((nb, 4), dtype=np.int32)
    segs[:, 0] = np.arange(nb)
    segs[:, 1] = np.arange(nb) % nk
    segs[:, 2] = 1      # conj(C^T) = C^dagger
    segs[:, 3] = 1      # S^T = conj(S) for Hermitian S
    Cinv = dev.empty((nb, nlo, nao), torch.complex128)
    dev.zgemm_tn(CT, Sd.reshape(-1, nao, nao), segs, Cinv, c_off=np.arange(nb, dtype=np.int64) * nlo * nao,
                 s_outer=nao, nbatch=nb, nseg=1)
    return Cinv


def transform_rdm1_to_lo(dm_ao_ao, C_ao_lo, S_ao_ao):
    r"""make_basis.py:591-620:  \gamma^{LO} = C^{-1} \gamma^{AO} C^{-1\dagger},  C^{-1} = C^\dagger S."""
    dm_ao_ao = dm_ao_ao if isinstance(dm_ao_ao, torch.Tensor) else np.asarray(dm_ao_ao)
    C_ao_lo = C_ao_lo if isinstance(C_ao_lo, torch.Tensor) else np.asarray(C_ao_lo)
    res_type = np.result_type(_np_dtype(dm_ao_ao), _np_dtype(C_ao_lo), _np_dtype(S_ao_ao))
    d4, C4, spin, nk, squeeze = _prep(dm_ao_ao, C_ao_lo)
    nlo = C4.shape[-1]
    dd, on1 = _zdev(d4)
    Cd, on2 = _zdev(C4)
    Sd, _ = _zdev(S_ao_ao)
    Cinv = _c_inverse(Cd, Sd, nk)
    h_index, a_index = _batch_maps(spin, nk, d4.shape[0], C4.shape[0])
    out = sandwich(Cinv, False, Cinv, True, dd.reshape(-1, dd.shape[-2], dd.shape[-1]), h_index, a_index)
    shape = (nk, nlo, nlo) if squeeze else (spin, nk, nlo, nlo)
    return _finish(out, shape, res_type, on1 and on2)


def transform_h1_to_ao(h_lo_lo, C_ao_lo, S_ao_ao):
    r"""make_basis.py:560-589:  h^{AO} = C^{-1\dagger} h^{LO} C^{-1}."""
    h_lo_lo = h_lo_lo if isinstance(h_lo_lo, torch.Tensor) else np.asarray(h_lo_lo)
    C_ao_lo = C_ao_lo if isinstance(C_ao_lo, torch.Tensor) else np.asarray(C_ao_lo)
    res_type = np.result_type(_np_dtype(h_lo_lo), _np_dtype(C_ao_lo), _np_dtype(S_ao_ao))
    h4, C4, spin, nk, squeeze = _prep(h_lo_lo, C_ao_lo)
    nao = C4.shape[-2]
    dev = get_device()
    hd, on1 = _zdev(h4)
    Cd, on2 = _zdev(C4)
    Sd, _ = _zdev(S_ao_ao)
    CinvT = dev.ztranspose(_c_inverse(Cd, Sd, nk))                   # (cs*nk, nao, nlo)
    h_index, a_index = _batch_maps(spin, nk, h4.shape[0], C4.shape[0])
    out = sandwich(CinvT, True, CinvT, False, hd.reshape(-1, hd.shape[-2], hd.shape[-1]), h_index, a_index)
    shape = (nk, nao, nao) if squeeze else (spin, nk, nao, nao)
    return _finish(out, shape, res_type, on1 and on2)


def transform_rdm1_to_ao(dm_lo_lo, C_ao_lo):
    r"""make_basis.py:622-644:  \gamma^{AO} = C \gamma^{LO} C^\dagger."""
    dm_lo_lo = dm_lo_lo if isinstance(dm_lo_lo, torch.Tensor) else np.asarray(dm_lo_lo)
    C_ao_lo = C_ao_lo if isinstance(C_ao_lo, torch.Tensor) else np.asarray(C_ao_lo)
    res_type = np.result_type(_np_dtype(dm_lo_lo), _np_dtype(C_ao_lo))
    d4, C4, spin, nk, squeeze = _prep(dm_lo_lo, C_ao_lo)
    nao = C4.shape[-2]
    dd, on1 = _zdev(d4)
    Cd, on2 = _zdev(C4)
    C3 = Cd.reshape(-1, nao, Cd.shape[-1])
    h_index, a_index = _batch_maps(spin, nk, d4.shape[0], C4.shape[0])
    out = sandwich(C3, False, C3, True, dd.reshape(-1, dd.shape[-2], dd.shape[-1]), h_index, a_index)
    shape = (nk, nao, nao) if squeeze else (spin, nk, nao, nao)
    return _finish(out, shape, res_type, on1 and on2)


def multiply_basis(C_ao_lo, C_lo_eo):
    """make_basis.py:923-962:  C_ao_eo = C_ao_lo . C_lo_eo per k (and spin)."""
    C_ao_lo = C_ao_lo if isinstance(C_ao_lo, torch.Tensor) else np.asarray(C_ao_lo)
    C_lo_eo = C_lo_eo if isinstance(C_lo_eo, torch.Tensor) else np.asarray(C_lo_eo)
    if C_ao_lo.ndim not in (3, 4) or C_lo_eo.ndim not in (3, 4):
        raise ValueError("invalid shape for multiply_basis: C_ao_lo shape %s, C_lo_eo shape: %s"
                         % (tuple(C_ao_lo.shape), tuple(C_lo_eo.shape)))
    res_type = np.result_type(_np_dtype(C_ao_lo), _np_dtype(C_lo_eo))
    squeeze = C_ao_lo.ndim == 3 and C_lo_eo.ndim == 3
    A4 = C_ao_lo if C_ao_lo.ndim == 4 else C_ao_lo[None]
    B4 = C_lo_eo if C_lo_eo.ndim == 4 else C_lo_eo[None]
    spin = max(A4.shape[0], B4.shape[0])
    nk, nlo, neo = B4.shape[-3:]
    nao = A4.shape[-2]
    dev = get_device()
    Ad, on1 = _zdev(A4)
    Bd, on2 = _zdev(B4)
    BT = dev.ztranspose(Bd.reshape(-1, nlo, neo))                    # (bs*nk, neo, nlo)
    nb = spin * nk
    b = np.arange(nb)
    s, k = b // nk, b % nk
    segs = np.zeros((nb, 4), dtype=np.int32)
    segs[:, 0] = np.minimum(s, A4.shape[0] - 1) * nk + k
    segs[:, 1] = np.minimum(s, B4.shape[0] - 1) * nk + k
    out = dev.empty((nb, nao, neo), torch.complex128)
    dev.zgemm_tn(Ad.reshape(-1, nao, nlo), BT, segs, out, c_off=b.astype(np.int64) * nao * neo, s_outer=neo,
                 nbatch=nb, nseg=1)
    shape = (nk, nao, neo) if squeeze else (spin, nk, nao, neo)
    return _finish(out, shape, res_type, on1 and on2)
