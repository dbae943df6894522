"""Oracle: k-space four-index integrals from the GDF tensor.

Numpy restatement of libdmet/routine/pbc_helper.py:276-294 (`get_eri_7d`) and 314-351 (`get_jk_from_eri_7d`).  The
reference obtains every (k_i, k_j, k_k, k_l) block from PySCF's `GDF.get_eri`, whose general-k branch contracts the
two three-index tensors WITHOUT conjugation over the auxiliary index (pyscf/pbc/df/df_ao2mo.py, `zdotNN` over
`sr_loop(kpti_kptj)` and `sr_loop(kptk_kptl)`):
    eri_7d[i, j, k][p, q, r, s] = sum_L L(k_i, k_j)[L, p, q] . L(k_k, k_l)[L, r, s],   k_l from momentum conservation
PySCF itself is absent here, so that convention is pinned by a consumer the reference does ship: J and K built from
the 7-d integrals with the reference's own formulas (`get_jk_from_eri_7d`) must equal the J and K of the supercell
integrals from `get_emb_eri` (tests/test_eri_7d.py).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import numpy as np

from .fourier import kpt_member


def get_kconserv(kpts_scaled):
    """kconserv[i, j, k] = l with k_i - k_j + k_k - k_l a reciprocal lattice vector (PySCF kpts_helper.get_kconserv)"""
    ks = np.asarray(kpts_scaled, dtype=float)
    nk = len(ks)
    out = np.zeros((nk, nk, nk), dtype=int)
    for i in range(nk):
        for j in range(nk):
            for k in range(nk):
                hit = kpt_member(ks[i] - ks[j] + ks[k], ks)
                assert len(hit) == 1
                out[i, j, k] = hit[0]
    return out


def get_eri_7d(cell, xdf, kpts=None, compact=False):
    """(nkpts, nkpts, nkpts, nao, nao, nao, nao) complex128 (pbc_helper.py:276-294)"""
    assert not compact
    nao, nk = xdf.nao, len(xdf.kpts_scaled)
    kconserv = get_kconserv(xdf.kpts_scaled)
    out = np.zeros((nk, nk, nk, nao, nao, nao, nao), dtype=np.complex128)
    for i in range(nk):
        for j in range(nk):
            Lij = np.asarray(xdf.load(i, j)).reshape(-1, nao * nao)
            for k in range(nk):
                Lkl = np.asarray(xdf.load(k, kconserv[i, j, k])).reshape(-1, nao * nao)
                out[i, j, k] = Lij.T.dot(Lkl).reshape((nao,) * 4)
    return out


def get_jk_from_eri_7d(eri, dm, with_j=True, with_k=True):
    """J and K per k-point from the spinless 7-d integrals (pbc_helper.py:314-351)"""
    eri, dm = np.asarray(eri), np.asarray(dm)
    shape = dm.shape
    dm = dm[None] if dm.ndim == 3 else dm
    spin, nk, nao, _ = dm.shape
    vj = np.zeros((spin, nk, nao, nao), dtype=np.complex128)
    vk = np.zeros((spin, nk, nao, nao), dtype=np.complex128)
    for s in range(spin):
        for k in range(nk):
            if with_j:
                vj[s] += np.einsum("Rpqrs,qp->Rrs", eri[k, k], dm[s, k])
            if with_k:
                vk[s] += np.einsum("Ppqrs,qr->Pps", eri[:, k, k], dm[s, k])
    return (vj.reshape(shape) / nk if with_j else None), (vk.reshape(shape) / nk if with_k else None)
