"""k-space four-index integrals from the GDF tensor on the device -- drop-in for
`libdmet.routine.pbc_helper.get_eri_7d` (pbc_helper.py:276-294; used by the reference's
basis_transform/test/test_transform_gdf.py:117 on the LO-basis GDF and by `Lattice.get_H2`, lattice.py:748).

    eri_7d[i, j, k][p, q, r, s] = sum_L L(k_i, k_j)[L, p, q] . L(k_k, k_l)[L, r, s],      k_i - k_j + k_k - k_l = G

(PySCF `GDF.get_eri`, general-k branch: no conjugation over the auxiliary index.)  Every stored block is transposed
once so that the auxiliary index is contiguous (`ldm_ztranspose`); the nkpts^3 products are complex TN GEMMs on the
FP64 tensor cores (`ldm_zgemm_tn`, one launch per k_i with nkpts^2 batch entries).  There is no CPU fallback."""
import numpy as np
import torch

from .device import get_device
from .schedule import kpt_member


def get_kconserv(kpts_scaled):
    """kconserv[i, j, k] = l with k_i - k_j + k_k - k_l a reciprocal lattice vector"""
    ks = np.asarray(kpts_scaled, dtype=float)
    nk = len(ks)
    out = np.zeros((nk, nk, nk), dtype=np.int64)
    for i in range(nk):
        for j in range(nk):
            d = ks[i] - ks[j]
            for k in range(nk):
                hit = kpt_member(d + ks[k], ks)
                assert len(hit) == 1
                out[i, j, k] = hit[0]
    return out


def get_eri_7d(cell, xdf, kpts=None, compact=False, return_device=False):
    """(nkpts, nkpts, nkpts, nao, nao, nao, nao) complex128.  `xdf`: anything `eri_transform.as_provider` accepts
    (in-memory provider, cderi file behind a GDF-like object, `LoGDF`)."""
    if compact:
        raise NotImplementedError("compact=True (packed real integrals) is not built")
    from . import eri_transform as et
    prov = et.as_provider(cell, xdf)
    if kpts is not None and len(kpts) != len(prov.kpts_scaled):
        raise NotImplementedError("band k-points other than the GDF mesh")
    dev = get_device()
    nao, naux, nk = int(prov.nao), int(prov.naux), len(prov.kpts_scaled)
    n2 = nao * nao
    kconserv = get_kconserv(prov.kpts_scaled)
    # LT[(i, j)] = L(k_i, k_j)^T : (nao^2, naux), auxiliary index contiguous
    LT = dev.empty((nk * nk, n2, naux), torch.complex128)
    blk = dev.empty((1, naux, n2), torch.complex128)
    synth = hasattr(prov, "keys") and hasattr(prov, "scale")
    for i in range(nk):
        for j in range(nk):
            if synth:
                dev.synth_block(blk[0].reshape(naux, nao, nao), naux, nao, prov.keys(i, j), prov.scale)
            else:
                L = prov.load(i, j)
                L = L if isinstance(L, torch.Tensor) else torch.from_numpy(
                    np.ascontiguousarray(L, dtype=np.complex128))
                blk[0].copy_(L.reshape(naux, n2))
            LT[i * nk + j].copy_(dev.ztranspose(blk)[0])
    out = dev.empty((nk, nk, nk, n2, n2), torch.complex128)
    for i in range(nk):
        segs = np.zeros((nk * nk, 4), dtype=np.int32)
        for j in range(nk):
            for k in range(nk):
                segs[j * nk + k] = (i * nk + j, k * nk + int(kconserv[i, j, k]), 0, 0)
        dev.zgemm_tn(LT, LT, segs, out[i], c_off=np.arange(nk * nk, dtype=np.int64) * n2 * n2, s_outer=n2,
                     nbatch=nk * nk, nseg=1)
    out = out.reshape((nk, nk, nk) + (nao,) * 4)
    return out if return_device else dev.to_host(out)
