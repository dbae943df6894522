"""Numpy restatements of the PySCF helpers the hot path calls (PySCF itself is absent from this image).

Semantics follow PySCF's public documentation/behaviour (pyscf>=2.0 as pinned by /root/reference/pyproject.toml:17);
call sites in the reference are cited per function.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np

# pyscf.pbc.lib.kpts_helper.KPT_DIFF_TOL (imported at libdmet/basis_transform/eri_transform.py:19)
KPT_DIFF_TOL = 1e-6

HERMITIAN = 1
ANTIHERMI = 2
SYMMETRIC = 3


def cartesian_prod(arrays):
    """pyscf.lib.cartesian_prod: C-ordered cartesian product (used at libdmet/system/fourier.py:42,52)."""
    arrays = [np.asarray(a) for a in arrays]
    grids = np.meshgrid(*arrays, indexing="ij")
    return np.stack([g.ravel() for g in grids], axis=-1)


def pack_tril(a, out=None):
    """pyscf.lib.pack_tril: a[..., i, j] (i >= j) -> [..., i*(i+1)/2 + j]   (eri_transform.py:133,139,375)."""
    a = np.asarray(a)
    n = a.shape[-1]
    idx = np.tril_indices(n)
    res = a[..., idx[0], idx[1]]
    if out is not None:
        flat = out.reshape(-1)[: res.size]
        flat[:] = res.reshape(-1)
        return flat.reshape(res.shape)
    return np.ascontiguousarray(res)


def unpack_tril(tril, filltriu=HERMITIAN):
    """pyscf.lib.unpack_tril: inverse of pack_tril; default fills the upper triangle with the conjugate of
    the lower one (eri_transform.py:217)."""
    tril = np.asarray(tril)
    npair = tril.shape[-1]
    n = int((np.sqrt(8 * npair + 1) - 1) // 2)
    assert n * (n + 1) // 2 == npair
    out = np.zeros(tril.shape[:-1] + (n, n), dtype=tril.dtype)
    idx = np.tril_indices(n)
    if filltriu == HERMITIAN:
        out[..., idx[1], idx[0]] = tril.conj()
    elif filltriu == SYMMETRIC:
        out[..., idx[1], idx[0]] = tril
    elif filltriu == ANTIHERMI:
        out[..., idx[1], idx[0]] = -tril.conj()
    out[..., idx[0], idx[1]] = tril            # lower triangle and diagonal are the stored values
    return out


def hermi_sum(a, axes=None, hermi=HERMITIAN, inplace=False):
    """pyscf.lib.hermi_sum: a + a.T.conj() (HERMITIAN) or a + a.T (SYMMETRIC) over the last two axes
    (eri_transform.py:372-373 uses axes=(0,2,1), hermi=SYMMETRIC, inplace=True)."""
    at = np.swapaxes(a, -1, -2)
    if hermi == HERMITIAN:
        at = at.conj()
    res = a + at
    if inplace:
        a[...] = res
        return a
    return res


def dot(a, b, alpha=1, c=None, beta=0):
    """pyscf.lib.dot: c <- alpha * a @ b + beta * c   (eri_transform.py:455-485)."""
    ab = np.dot(a, b)
    if alpha != 1:
        ab = ab * alpha
    if c is None:
        return ab
    if beta == 0:
        c[...] = ab
    else:
        if beta != 1:
            c *= beta
        c += ab
    return c


def conc_mos(moi, moj):
    """pyscf.ao2mo.incore._conc_mos(moi, moj)[2:] -> (mo_pq, (0, ni, ni, ni+nj))   (eri_transform.py:432)."""
    ni, nj = moi.shape[1], moj.shape[1]
    return np.hstack((moi, moj)), (0, ni, ni, ni + nj)


def r_e2(Lpq, mo, pqslice, out=None):
    """pyscf.ao2mo._ao2mo.r_e2 (complex second half-transformation, no pair symmetry):
    out[L, i*nj + j] = sum_pq conj(mo_i[p, i]) * Lpq[L, p, q] * mo_j[q, j]     (eri_transform.py:433)."""
    i0, i1, j0, j1 = pqslice
    nao = mo.shape[0]
    L = np.asarray(Lpq).reshape(-1, nao, nao)
    ci = mo[:, i0:i1]
    cj = mo[:, j0:j1]
    # two large zgemm calls (all BLAS threads busy) instead of 2 small ones per auxiliary row
    nL, ni, nj = L.shape[0], ci.shape[1], cj.shape[1]
    half = np.dot(L.reshape(nL * nao, nao), cj).reshape(nL, nao, nj)            # (L, p, j)
    res = np.dot(ci.conj().T, half.transpose(1, 0, 2).reshape(nao, nL * nj))    # (i, L*j)
    res = res.reshape(ni, nL, nj).transpose(1, 0, 2).reshape(nL, ni * nj)
    if out is not None:
        out[...] = res
        return out
    return res


def restore(symmetry, eri, norb):
    """pyscf.ao2mo.restore for a 4-fold (npair, npair) input: target 1 / 4 / 8   (eri_transform.py:529,543)."""
    eri = np.asarray(eri)
    npair = norb * (norb + 1) // 2
    symmetry = int(str(symmetry).replace("s", ""))
    if eri.size == norb ** 4:
        eri1 = eri.reshape(norb, norb, norb, norb)
        if symmetry == 1:
            return eri1
        idx = np.tril_indices(norb)
        eri4 = eri1[idx[0], idx[1]][:, idx[0], idx[1]]
    elif eri.size == npair * npair:
        eri4 = eri.reshape(npair, npair)
    elif eri.size == npair * (npair + 1) // 2:
        eri4 = unpack_tril(eri.reshape(-1), SYMMETRIC)
    else:
        raise ValueError("restore: unknown eri size")
    if symmetry == 4:
        return eri4
    if symmetry == 8:
        return pack_tril(eri4)
    if symmetry == 1:
        tri = np.zeros((norb, norb), dtype=np.int64)
        idx = np.tril_indices(norb)
        tri[idx[0], idx[1]] = np.arange(npair)
        tri[idx[1], idx[0]] = np.arange(npair)
        return np.ascontiguousarray(eri4[tri.ravel()][:, tri.ravel()]).reshape(norb, norb, norb, norb)
    raise ValueError("restore: unknown symmetry %s" % symmetry)


def dot_eri_dm(eri, dm, hermi=0, with_j=True, with_k=True):
    """pyscf.scf.hf.dot_eri_dm: J_ij = sum_kl (ij|kl) D_kl ; K_jk = sum_il (ij|kl) D_il
    (convention quoted at libdmet/solver/scf.py:269-271).  eri is s1 / s4 / s8; dm is (n,n) or (nset,n,n)."""
    dm = np.asarray(dm)
    single = dm.ndim == 2
    dms = dm[None] if single else dm
    n = dms.shape[-1]
    npair = n * (n + 1) // 2
    eri = np.asarray(eri)
    if eri.size == npair * (npair + 1) // 2 and eri.size != npair * npair:
        eri = restore(4, eri, n)
    vj = vk = None
    if eri.size == npair * npair and eri.ndim <= 2:
        eri4 = eri.reshape(npair, npair)
        idx = np.tril_indices(n)
        if with_j:
            vj = []
            for d in dms:
                dp = d + d.T
                dp[np.diag_indices(n)] *= 0.5
                vj.append(unpack_tril(eri4 @ dp[idx], SYMMETRIC))
            vj = np.asarray(vj)
        if with_k:
            eri1 = restore(1, eri4, n)
            vk = np.asarray([np.einsum("ijkl,il->jk", eri1, d) for d in dms])
    else:
        eri1 = eri.reshape(n, n, n, n)
        if with_j:
            vj = np.asarray([np.einsum("ijkl,kl->ij", eri1, d) for d in dms])
        if with_k:
            vk = np.asarray([np.einsum("ijkl,il->jk", eri1, d) for d in dms])
    if single:
        vj = None if vj is None else vj[0]
        vk = None if vk is None else vk[0]
    return vj, vk


# ---------------------------------------------------------------------------------------------------------
# cderi file access: pyscf.pbc.df.df._load3c (called at eri_transform.py:221) and pyscf.df.addons.load (171)
# ---------------------------------------------------------------------------------------------------------
def kpts_member(kpt, kpts):
    """pyscf.pbc.lib.kpts_helper.member: positions of `kpt` in `kpts` (rows compared within KPT_DIFF_TOL)"""
    kpts = np.asarray(kpts).reshape(len(kpts), -1)
    return np.where(np.abs(kpts - np.asarray(kpt).reshape(1, -1)).max(axis=1) < KPT_DIFF_TOL)[0]


class KPairLoader(object):
    """What `_load3c.__enter__` hands out for one k-point pair: `.shape` and `[rows]` over the stored entry.  Column
    segments '0', '1', ... of a group are stacked horizontally; when only the swapped pair is stored, full
    (nao*nao column) data come back conjugate-transposed in the two AO indices and Hermitian-packed data
    conjugated."""

    def __init__(self, entry, swapped, nao):
        self.segs = [entry[str(n)] for n in range(len(entry))] if hasattr(entry, "keys") else [entry]
        self.swapped, self.nao = swapped, nao

    @property
    def shape(self):
        return (self.segs[0].shape[0], sum(s.shape[1] for s in self.segs))

    def __getitem__(self, rows):
        v = np.hstack([np.asarray(s[rows]) for s in self.segs])
        if not self.swapped:
            return v
        nao = self.nao
        if v.shape[-1] == nao * nao:
            return np.ascontiguousarray(v.reshape(-1, nao, nao).transpose(0, 2, 1).conj()).reshape(-1, nao * nao)
        return v.conj()


def load3c(feri, label, kpti_kptj, kptij_label, nao):
    """`_load3c(cderi, label, kpti_kptj, kptij_label)` on an open h5py-like file: the stored pair, else the swapped
    one.  Files without the pair list (PySCF >= 2.1: `kpts` + `j3c/<ki * nkpts + kj>`) are addressed by index."""
    kpti_kptj = np.asarray(kpti_kptj)
    if kptij_label in feri:
        kptij = np.asarray(feri[kptij_label][...])
        hit = kpts_member(kpti_kptj, kptij)
        if len(hit):
            return KPairLoader(feri["%s/%d" % (label, hit[0])], False, nao)
        hit = kpts_member(kpti_kptj[[1, 0]], kptij)
        if not len(hit):
            raise KeyError("%s for the k-point pair %s is not in the file" % (label, kpti_kptj))
        return KPairLoader(feri["%s/%d" % (label, hit[0])], True, nao)
    kpts = np.asarray(feri["kpts"][...])
    ki, kj = int(kpts_member(kpti_kptj[0], kpts)[0]), int(kpts_member(kpti_kptj[1], kpts)[0])
    for key, swapped in (("%s/%d" % (label, ki * len(kpts) + kj), False),
                         ("%s/%d" % (label, kj * len(kpts) + ki), True)):
        if key in feri:
            return KPairLoader(feri[key], swapped, nao)
    raise KeyError("%s for the k-point pair (%d, %d) is not in the file" % (label, ki, kj))
