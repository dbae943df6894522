"""Oracle: embedding ERI from Gaussian-density-fitting integrals.

Line-by-line numpy restatement of libdmet/basis_transform/eri_transform.py:44-112 (dispatch), 118-157, 195-227
(sr_loop chunking), 235-399 (get_emb_eri_fast_gdf, incore), 403-434, 436-485, 523-544.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`mydf` is any GDF provider with
    .kpts_scaled (nkpts, 3)   scaled k-points (the reference obtains them via cell.get_scaled_kpts(mydf.kpts), l.266)
    .kmesh                    k-mesh (for get_phase_R2k, l.291)
    .nao, .naux               (cell.nao_nr() l.256; get_naoaux l.263)
    .load(ki, kj)             (naux, nao, nao) complex128 block L(k_i, k_j) as PySCF's _load3c returns it (l.221)
    .blockdim                 PySCF GDF.blockdim (l.333), default 240
"""
import numpy as np

from . import pyscf_lib as lib
from .pyscf_lib import KPT_DIFF_TOL
from .fourier import round_to_FBZ, kpt_member, get_phase_R2k_scaled, max_abs
from .make_basis import multiply_basis, add_spin_dim

ERI_IMAG_TOL = 1e-6   # eri_transform.py:32


def get_basis_k(basis, phase_R2k):
    """eri_transform.py:118-126."""
    spin = basis.shape[0]
    basis_k = np.empty_like(basis, dtype=np.complex128)
    for s in range(spin):
        basis_k[s] = np.einsum('Rim,Rk->kim', basis[s], phase_R2k)
    return basis_k


def get_weights_t_reversal(kpts_scaled, tol=KPT_DIFF_TOL):
    """eri_transform.py:142-157."""
    nkpts = len(kpts_scaled)
    kpts_round = round_to_FBZ(np.array(kpts_scaled, dtype=float), tol=tol)
    weights = np.ones(nkpts, dtype=int)
    for i, ki in enumerate(kpts_round):
        if weights[i] == 1:
            for j in range(i + 1, nkpts):
                sum_ij = ki + kpts_round[j]
                sum_ij -= np.round(sum_ij)
                if max_abs(sum_ij) < tol:
                    weights[i] = 2
                    weights[j] = 0
                    break
    assert np.sum(weights) == nkpts
    return weights


def sr_loop(mydf, ki, kj, blksize):
    """eri_transform.py:195-227 with compact=False: yields (<=blksize, nao*nao) complex128 chunks."""
    Lfull = mydf.load(ki, kj)
    naux = Lfull.shape[0]
    for b0 in range(0, naux, blksize):
        b1 = min(naux, b0 + blksize)
        yield np.asarray(Lfull[b0:b1], dtype=np.complex128).reshape(b1 - b0, -1)


def transform_ao_to_emb(Lpq, basis, kp, kq, Lpq_beta=None):
    """eri_transform.py:403-434."""
    if basis.ndim == 3:
        basis = basis[np.newaxis]
    spin, ncells, nlo, nemb = basis.shape
    if Lpq_beta is None:
        Lpq = [Lpq for s in range(spin)]
    else:
        Lpq = [Lpq, Lpq_beta]
    nL = Lpq[0].shape[0]
    Lij = np.empty((spin, nL, nemb * nemb), dtype=np.complex128)
    for s in range(spin):
        mopq, pqslice = lib.conc_mos(basis[s, kp], basis[s, kq])
        lib.r_e2(Lpq[s], mopq, pqslice, out=Lij[s])
    return Lij


def _Lij_s4_to_eri(Lij_s4, eri, weight=1, t_reversal_symm=False):
    """eri_transform.py:436-485 (incore branch)."""
    if Lij_s4.ndim == 2:
        Lij_s4 = Lij_s4[np.newaxis]
    spin, nL, nemb_pair = Lij_s4.shape
    if t_reversal_symm:
        if spin == 1:
            Lij_loc = np.asarray(Lij_s4[0].real, order='C')
            if weight == 1:
                lib.dot(Lij_loc.T, Lij_loc, 1.0, eri[0], 1)
            elif weight == 2:
                lib.dot(Lij_loc.T, Lij_loc, 2.0, eri[0], 1)
                Lij_loc = np.asarray(Lij_s4[0].imag, order='C')
                lib.dot(Lij_loc.T, Lij_loc, 2.0, eri[0], 1)
            else:
                raise ValueError
        else:
            Lij_loc_a, Lij_loc_b = np.asarray(Lij_s4.real, order='C')
            if weight == 1:
                lib.dot(Lij_loc_a.T, Lij_loc_a, 1.0, eri[0], 1)
                lib.dot(Lij_loc_a.T, Lij_loc_b, 1.0, eri[1], 1)
                lib.dot(Lij_loc_b.T, Lij_loc_b, 1.0, eri[2], 1)
            elif weight == 2:
                lib.dot(Lij_loc_a.T, Lij_loc_a, 2.0, eri[0], 1)
                lib.dot(Lij_loc_a.T, Lij_loc_b, 2.0, eri[1], 1)
                lib.dot(Lij_loc_b.T, Lij_loc_b, 2.0, eri[2], 1)
                Lij_loc_a, Lij_loc_b = np.asarray(Lij_s4.imag, order='C')
                lib.dot(Lij_loc_a.T, Lij_loc_a, 2.0, eri[0], 1)
                lib.dot(Lij_loc_a.T, Lij_loc_b, 2.0, eri[1], 1)
                lib.dot(Lij_loc_b.T, Lij_loc_b, 2.0, eri[2], 1)
            else:
                raise ValueError
    else:
        if spin == 1:
            lib.dot(Lij_s4[0].conj().T, Lij_s4[0], 1, eri[0], 1)
        else:
            lib.dot(Lij_s4[0].conj().T, Lij_s4[0], 1, eri[0], 1)
            lib.dot(Lij_s4[0].conj().T, Lij_s4[1], 1, eri[1], 1)
            lib.dot(Lij_s4[1].conj().T, Lij_s4[1], 1, eri[2], 1)


def eri_restore(eri, symmetry, nemb):
    """eri_transform.py:523-544."""
    spin_pair = eri.shape[0]
    if spin_pair == 1:
        eri_res = lib.restore(symmetry, eri[0].real, nemb)[np.newaxis]
    else:
        if symmetry == 4:
            nemb_pair = nemb * (nemb + 1) // 2
            if eri.size == spin_pair * nemb_pair * nemb_pair:
                return eri.real.reshape(spin_pair, nemb_pair, nemb_pair)
            eri_res = np.empty((spin_pair, nemb_pair, nemb_pair))
        elif symmetry == 1:
            if eri.size == spin_pair * nemb ** 4:
                return eri.real.reshape(spin_pair, nemb, nemb, nemb, nemb)
            eri_res = np.empty((spin_pair, nemb, nemb, nemb, nemb))
        else:
            raise ValueError("Spin unrestricted ERI does not support 8-fold symmetry.")
        for s in range(spin_pair):
            eri_res[s] = lib.restore(symmetry, eri[s].real, nemb)
    return eri_res


def build_C_ao_emb(mydf, C_ao_lo=None, basis=None, C_ao_eo=None, unit_eri=False):
    """eri_transform.py:270-300: the (spin, nkpts, nao, nemb) coefficients, already scaled by nkpts**-0.75."""
    nao = mydf.nao
    nkpts = len(mydf.kpts_scaled)
    if C_ao_eo is None:
        if C_ao_lo is None:
            C_ao_lo = np.zeros((nkpts, nao, nao), dtype=np.complex128)
            C_ao_lo[:, range(nao), range(nao)] = 1.0
        C_ao_lo = np.asarray(C_ao_lo)
        if C_ao_lo.ndim == 3:
            C_ao_lo = C_ao_lo[np.newaxis]
        if basis is None:
            basis = np.eye(nkpts * nao).reshape(1, nkpts, nao, nkpts * nao)
        if basis.shape[0] < C_ao_lo.shape[0]:
            basis = add_spin_dim(basis, C_ao_lo.shape[0])
        if C_ao_lo.shape[0] < basis.shape[0]:
            C_ao_lo = add_spin_dim(C_ao_lo, basis.shape[0])
        if unit_eri:
            C_ao_emb = C_ao_lo / (nkpts ** 0.75)
        else:
            phase = get_phase_R2k_scaled(mydf.kmesh, mydf.kpts_scaled)
            C_ao_emb = multiply_basis(C_ao_lo, get_basis_k(basis, phase)) / (nkpts ** 0.75)
    else:
        if C_ao_lo is not None:
            raise ValueError("Don't pass both `C_ao_lo` and `C_ao_eo`.")
        C_ao_eo = np.asarray(C_ao_eo)
        if C_ao_eo.ndim == 3:
            C_ao_eo = C_ao_eo[np.newaxis]
        assert (nkpts, nao) == C_ao_eo.shape[1:3]
        C_ao_emb = C_ao_eo / (nkpts ** 0.75)
    return C_ao_emb


def get_emb_eri_fast_gdf(cell, mydf, C_ao_lo=None, basis=None, feri=None,
                         kscaled_center=None, symmetry=4, max_memory=None,
                         C_ao_eo=None, kconserv_tol=KPT_DIFF_TOL, unit_eri=False, swap_idx=None,
                         t_reversal_symm=True, incore=True, fout="H2.h5", kL_subset=None, restore=True):
    """eri_transform.py:235-399, incore.  `kL_subset` / `restore=False` are oracle-only hooks used by the
    multi-rank tests (the reference's MPI variant shards the same loop, eri_transform_mpi.py:151-157)."""
    assert incore, "oracle restates the incore branch only"
    nao = mydf.nao
    nkpts = len(mydf.kpts_scaled)
    naux = mydf.naux
    kscaled = np.array(mydf.kpts_scaled, dtype=float)
    if kscaled_center is not None:
        kscaled = kscaled - kscaled_center

    C_ao_emb = build_C_ao_emb(mydf, C_ao_lo, basis, C_ao_eo, unit_eri)
    spin, _, _, nemb = C_ao_emb.shape
    nemb_pair = nemb * (nemb + 1) // 2
    res_shape = (spin * (spin + 1) // 2, nemb_pair, nemb_pair)

    if t_reversal_symm:
        # NOTE the reference calls get_weights_t_reversal(cell, kpts) on the UNSHIFTED k-points (l.309)
        weights = get_weights_t_reversal(mydf.kpts_scaled)
        eri = np.zeros(res_shape)
    else:
        weights = np.ones((nkpts,), dtype=int)
        eri = np.zeros(res_shape, dtype=np.complex128)

    if max_memory is None:
        max_memory = 2000
    blksize = max_memory * 1e6 / 16 / (nao ** 2 * 2)
    blksize = max(16, min(int(blksize), getattr(mydf, "blockdim", 240)))
    Lij_s4 = np.empty((spin, naux, nemb_pair), dtype=np.complex128)

    for kL in range(nkpts):
        if weights[kL] <= 0:
            continue
        if kL_subset is not None and kL not in kL_subset:
            continue
        Lij_s4[:] = 0.0
        i_visited = np.zeros((nkpts,), dtype=bool)
        for i in range(nkpts):
            if i_visited[i]:
                continue
            i_visited[i] = True
            for j in range(nkpts):
                kconserv = -kscaled[i] + kscaled[j] + kscaled[kL]
                if max_abs(np.round(kconserv) - kconserv) > kconserv_tol:
                    continue
                if t_reversal_symm:
                    jm = kpt_member(-kscaled[j], kscaled)
                    assert len(jm) == 1
                    jm = jm[0]
                step0, step1 = 0, 0
                for Lpq in sr_loop(mydf, i, j, blksize):
                    lchunk = Lpq.shape[0]
                    step0, step1 = step1, step1 + lchunk
                    Lij_loc = transform_ao_to_emb(Lpq, C_ao_emb, i, j).reshape(-1, nemb, nemb)
                    if t_reversal_symm and (not i_visited[jm]):
                        lib.hermi_sum(Lij_loc, axes=(0, 2, 1), hermi=lib.SYMMETRIC, inplace=True)
                    buf = lib.pack_tril(Lij_loc)
                    Lij_s4[:, step0:step1] += buf.reshape(spin, lchunk, nemb_pair)
                if t_reversal_symm:
                    i_visited[jm] = True
        _Lij_s4_to_eri(Lij_s4, eri, weight=weights[kL], t_reversal_symm=t_reversal_symm)

    if not restore:
        return eri
    if not t_reversal_symm:
        eri = eri.real
    eri = eri_restore(eri, symmetry, nemb)
    return eri


def get_emb_eri(cell, mydf, C_ao_lo=None, basis=None, unit_eri=False, symmetry=4, t_reversal_symm=True,
                max_memory=None, swap_idx=None, feri=None, kscaled_center=None, kconserv_tol=KPT_DIFF_TOL,
                incore=True, fout="H2.h5", **kwargs):
    """eri_transform.py:44-94 (GDF branch only)."""
    return get_emb_eri_fast_gdf(cell, mydf, C_ao_lo=C_ao_lo, basis=basis, feri=feri,
                                kscaled_center=kscaled_center, symmetry=symmetry, max_memory=max_memory,
                                kconserv_tol=kconserv_tol, unit_eri=unit_eri, swap_idx=swap_idx,
                                t_reversal_symm=t_reversal_symm, incore=incore, fout=fout)


def get_unit_eri(cell, mydf, C_ao_lo=None, symmetry=4, t_reversal_symm=True, max_memory=None, swap_idx=None,
                 feri=None, kscaled_center=None, kconserv_tol=KPT_DIFF_TOL, incore=True, fout="H2.h5", **kwargs):
    """eri_transform.py:96-112."""
    C_ao_lo = np.asarray(C_ao_lo)
    if C_ao_lo.ndim == 3:
        C_ao_lo = C_ao_lo[np.newaxis]
    basis = np.empty_like(C_ao_lo)
    return get_emb_eri(cell, mydf, C_ao_lo=C_ao_lo, basis=basis, feri=feri, kscaled_center=kscaled_center,
                       symmetry=symmetry, max_memory=max_memory, kconserv_tol=kconserv_tol, unit_eri=True,
                       swap_idx=swap_idx, t_reversal_symm=t_reversal_symm, incore=incore, fout=fout, **kwargs)


def _Lij_s4_to_eri_gso(Lij_s4, eri, weight=1, t_reversal_symm=False):
    """eri_transform.py:1252-1284 (incore branch)."""
    if t_reversal_symm:
        Lij_loc_a, Lij_loc_b = np.asarray(Lij_s4.real, order='C')
        if weight == 1:
            lib.dot(Lij_loc_a.T, Lij_loc_a, 1.0, eri[0], 1)
            lib.dot(Lij_loc_b.T, Lij_loc_b, 1.0, eri[0], 1)
            lib.dot(Lij_loc_a.T, Lij_loc_b, -1.0, eri[0], 1)
            lib.dot(Lij_loc_b.T, Lij_loc_a, -1.0, eri[0], 1)
        elif weight == 2:
            for part in (Lij_s4.real, Lij_s4.imag):
                Lij_loc_a, Lij_loc_b = np.asarray(part, order='C')
                lib.dot(Lij_loc_a.T, Lij_loc_a, 2.0, eri[0], 1)
                lib.dot(Lij_loc_b.T, Lij_loc_b, 2.0, eri[0], 1)
                lib.dot(Lij_loc_a.T, Lij_loc_b, -2.0, eri[0], 1)
                lib.dot(Lij_loc_b.T, Lij_loc_a, -2.0, eri[0], 1)
        else:
            raise ValueError
    else:
        lib.dot(Lij_s4[0].conj().T, Lij_s4[0], 1.0, eri[0], 1)
        lib.dot(Lij_s4[1].conj().T, Lij_s4[1], 1.0, eri[0], 1)
        tmp_ab = lib.dot(Lij_s4[0].conj().T, Lij_s4[1], -1.0)
        eri[0] += tmp_ab
        eri[0] += tmp_ab.conj().T


def get_emb_eri_gso(cell, mydf, C_ao_lo=None, basis=None, feri=None, kscaled_center=None, symmetry=4,
                    max_memory=None, kconserv_tol=KPT_DIFF_TOL, unit_eri=False, swap_idx=None,
                    t_reversal_symm=True, basis_k=None, incore=True, fout="H2.h5"):
    """eri_transform.py:1104-1250 (incore): GSO embedding ERI with partial particle-hole transform."""
    assert incore
    nao = mydf.nao
    nkpts = len(mydf.kpts_scaled)
    naux = mydf.naux
    C_ao_lo = add_spin_dim(C_ao_lo, 2)
    kscaled = np.array(mydf.kpts_scaled, dtype=float)
    if kscaled_center is not None:
        kscaled = kscaled - kscaled_center
    if basis_k is None:
        assert basis is not None and basis.ndim == 3
        phase = get_phase_R2k_scaled(mydf.kmesh, mydf.kpts_scaled)
        basis_k = get_basis_k(basis[None], phase)[0]
    if basis_k.ndim == 3:
        nso = basis_k.shape[1] // 2          # separate_basis, libdmet/routine/spinless_helper.py:31-46
        basis_k = np.asarray((basis_k[:, :nso], basis_k[:, nso:]))
    if unit_eri:
        C_ao_emb = C_ao_lo / (nkpts ** 0.75)
    else:
        C_ao_emb = multiply_basis(C_ao_lo, basis_k) / (nkpts ** 0.75)
    spin, _, _, nemb = C_ao_emb.shape
    nemb_pair = nemb * (nemb + 1) // 2
    res_shape = (1, nemb_pair, nemb_pair)
    if t_reversal_symm:
        weights = get_weights_t_reversal(mydf.kpts_scaled)
        eri = np.zeros(res_shape)
    else:
        weights = np.ones((nkpts,), dtype=int)
        eri = np.zeros(res_shape, dtype=np.complex128)
    if max_memory is None:
        max_memory = 2000
    blksize = max_memory * 1e6 / 16 / (nao ** 2 * 2)
    blksize = max(16, min(int(blksize), getattr(mydf, "blockdim", 240)))
    Lij_s4 = np.empty((spin, naux, nemb_pair), dtype=np.complex128)
    for kL in range(nkpts):
        if weights[kL] <= 0:
            continue
        Lij_s4[:] = 0.0
        i_visited = np.zeros((nkpts,), dtype=bool)
        for i in range(nkpts):
            if i_visited[i]:
                continue
            i_visited[i] = True
            for j in range(nkpts):
                kconserv = -kscaled[i] + kscaled[j] + kscaled[kL]
                if max_abs(np.round(kconserv) - kconserv) > kconserv_tol:
                    continue
                if t_reversal_symm:
                    jm = kpt_member(-kscaled[j], kscaled)
                    assert len(jm) == 1
                    jm = jm[0]
                step0, step1 = 0, 0
                for Lpq in sr_loop(mydf, i, j, blksize):
                    lchunk = Lpq.shape[0]
                    step0, step1 = step1, step1 + lchunk
                    Lij_loc = transform_ao_to_emb(Lpq, C_ao_emb, i, j).reshape(-1, nemb, nemb)
                    if t_reversal_symm and (not i_visited[jm]):
                        lib.hermi_sum(Lij_loc, axes=(0, 2, 1), hermi=lib.SYMMETRIC, inplace=True)
                    Lij_s4[:, step0:step1] += lib.pack_tril(Lij_loc).reshape(spin, lchunk, nemb_pair)
                if t_reversal_symm:
                    i_visited[jm] = True
        _Lij_s4_to_eri_gso(Lij_s4, eri, weight=weights[kL], t_reversal_symm=t_reversal_symm)
    if not t_reversal_symm:
        eri = eri.real
    return eri_restore(eri, symmetry, nemb)
