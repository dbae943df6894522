"""Oracle: embedding ERI from Gaussian-density-fitting integrals.

Numpy restatement (own code, same algorithm) of libdmet/basis_transform/eri_transform.py:44-112 (dispatch), 118-157, 195-227
(sr_loop chunking), 235-399 (get_emb_eri_fast_gdf, incore), 403-434, 436-485, 523-544, 1312-1427 (GDF tensor in the LO basis).
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
    """Fourier transform of the embedding basis, basis_k[s, k] = sum_R basis[s, R] exp(-ikR)
    (eri_transform.py:118-126)"""
    return np.einsum("sRim,Rk->skim", np.asarray(basis), phase_R2k).astype(np.complex128)


def get_weights_t_reversal(kpts_scaled, tol=KPT_DIFF_TOL):
    """time-reversal weights (eri_transform.py:142-157): scanning k-points in order, a k whose partner -k appears
    later gets weight 2 and the partner weight 0; self-conjugate points keep 1."""
    kr = round_to_FBZ(np.array(kpts_scaled, dtype=float), tol=tol)
    nk = len(kr)
    w = [1] * nk
    for i in range(nk):
        if w[i] != 1:
            continue
        for j in range(i + 1, nk):
            s = kr[i] + kr[j]
            if max_abs(s - np.round(s)) < tol:
                w[i], w[j] = 2, 0
                break
    assert sum(w) == nk
    return np.asarray(w, dtype=int)


def sr_loop(mydf, ki, kj, blksize):
    """auxiliary-index chunks of L(k_i, k_j), each (<= blksize, nao*nao) complex128 -- what the reference's
    generator yields with compact=False (eri_transform.py:195-227)"""
    block = np.asarray(mydf.load(ki, kj), dtype=np.complex128)
    for start in range(0, block.shape[0], blksize):
        chunk = block[start:start + blksize]
        yield chunk.reshape(chunk.shape[0], -1)


def transform_ao_to_emb(Lpq, basis, kp, kq, Lpq_beta=None):
    """Lij[s, L, (m, n)] = sum_pq conj(C[s, kp][p, m]) Lpq[L, (p, q)] C[s, kq][q, n]   (eri_transform.py:403-434;
    PySCF's r_e2 on the concatenated coefficient matrix)"""
    basis = basis[None] if basis.ndim == 3 else basis
    spin, nemb = basis.shape[0], basis.shape[-1]
    sources = [Lpq] * spin if Lpq_beta is None else [Lpq, Lpq_beta]
    out = np.empty((spin, sources[0].shape[0], nemb * nemb), dtype=np.complex128)
    for s in range(spin):
        mo, sl = lib.conc_mos(basis[s, kp], basis[s, kq])
        lib.r_e2(sources[s], mo, sl, out=out[s])
    return out


def _Lij_s4_to_eri(Lij_s4, eri, weight=1, t_reversal_symm=False):
    """eri += w * Gram products of the packed 3-index tensor (eri_transform.py:436-485, incore).
    Time reversal: real parts only for w = 1, real and imaginary parts with factor 2 for w = 2; spin blocks
    aa -> eri[0], ab -> eri[1], bb -> eri[2].  Otherwise the complex Lambda^dagger Lambda."""
    Lij_s4 = Lij_s4[None] if Lij_s4.ndim == 2 else Lij_s4
    spin = Lij_s4.shape[0]
    targets = [(0, 0, 0)] if spin == 1 else [(0, 0, 0), (0, 1, 1), (1, 1, 2)]     # (bra spin, ket spin, eri block)
    if not t_reversal_symm:
        for a, b, blk in targets:
            lib.dot(Lij_s4[a].conj().T, Lij_s4[b], 1, eri[blk], 1)
        return
    if weight not in (1, 2):
        raise ValueError
    parts = [Lij_s4.real] if weight == 1 else [Lij_s4.real, Lij_s4.imag]
    for part in parts:
        part = np.ascontiguousarray(part)
        for a, b, blk in targets:
            lib.dot(part[a].T, part[b], float(weight), eri[blk], 1)


def eri_restore(eri, symmetry, nemb):
    """s4 -> requested symmetry per spin block (eri_transform.py:523-544)"""
    spin_pair = eri.shape[0]
    if spin_pair > 1 and symmetry not in (1, 4):
        raise ValueError("Spin unrestricted ERI does not support 8-fold symmetry.")
    return np.stack([lib.restore(symmetry, eri[s].real, nemb) for s in range(spin_pair)])


def build_C_ao_emb(mydf, C_ao_lo=None, basis=None, C_ao_eo=None, unit_eri=False):
    """(spin, nkpts, nao, nemb) coefficients scaled by nkpts**-0.75 (eri_transform.py:270-300)"""
    nao, nkpts = mydf.nao, len(mydf.kpts_scaled)
    norm = nkpts ** 0.75
    if C_ao_eo is not None:
        if C_ao_lo is not None:
            raise ValueError("Don't pass both `C_ao_lo` and `C_ao_eo`.")
        C_ao_eo = np.asarray(C_ao_eo)
        C_ao_eo = C_ao_eo[None] if C_ao_eo.ndim == 3 else C_ao_eo
        assert C_ao_eo.shape[1:3] == (nkpts, nao)
        return C_ao_eo / norm
    if C_ao_lo is None:                                   # plain AO basis
        C_ao_lo = np.broadcast_to(np.eye(nao, dtype=np.complex128), (nkpts, nao, nao)).copy()
    C_ao_lo = np.asarray(C_ao_lo)
    C_ao_lo = C_ao_lo[None] if C_ao_lo.ndim == 3 else C_ao_lo
    if unit_eri:
        return C_ao_lo / norm
    if basis is None:                                     # every LO of every cell is an "embedding" orbital
        basis = np.eye(nkpts * nao).reshape(1, nkpts, nao, nkpts * nao)
    spin = max(basis.shape[0], C_ao_lo.shape[0])
    basis, C_ao_lo = add_spin_dim(basis, spin), add_spin_dim(C_ao_lo, spin)
    phase = get_phase_R2k_scaled(mydf.kmesh, mydf.kpts_scaled)
    return multiply_basis(C_ao_lo, get_basis_k(basis, phase)) / norm


def _accumulate_Lij_s4(mydf, C_ao_emb, kL, kscaled, kconserv_tol, t_reversal_symm, blksize):
    """packed 3-index tensor of one transfer momentum: the i / j loops of eri_transform.py:343-382.
    For every not-yet-visited k_i, the k_j with -k_i + k_j + k_L on the reciprocal lattice; with time reversal the
    block is symmetrised (Lij + Lij^T) when its partner pair (-k_j, -k_i) has not been visited, and k_(-j) is marked
    visited afterwards."""
    spin, nkpts, _, nemb = C_ao_emb.shape
    npair = nemb * (nemb + 1) // 2
    out = np.zeros((spin, mydf.naux, npair), dtype=np.complex128)
    seen = np.zeros(nkpts, dtype=bool)
    for i in range(nkpts):
        if seen[i]:
            continue
        seen[i] = True
        for j in range(nkpts):
            q = kscaled[j] - kscaled[i] + kscaled[kL]
            if max_abs(np.round(q) - q) > kconserv_tol:
                continue
            partner = None
            if t_reversal_symm:
                hit = kpt_member(-kscaled[j], kscaled)
                assert len(hit) == 1
                partner = hit[0]
            row = 0
            for Lpq in sr_loop(mydf, i, j, blksize):
                n = Lpq.shape[0]
                Lij = transform_ao_to_emb(Lpq, C_ao_emb, i, j).reshape(-1, nemb, nemb)
                if partner is not None and not seen[partner]:
                    lib.hermi_sum(Lij, axes=(0, 2, 1), hermi=lib.SYMMETRIC, inplace=True)
                out[:, row:row + n] += lib.pack_tril(Lij).reshape(spin, n, npair)
                row += n
            if partner is not None:
                seen[partner] = True
    return out


def _chunk_rows(mydf, max_memory):
    """auxiliary chunk length of the reference (eri_transform.py:330-333)"""
    max_memory = 2000 if max_memory is None else max_memory
    rows = int(max_memory * 1e6 / 16 / (mydf.nao ** 2 * 2))
    return max(16, min(rows, getattr(mydf, "blockdim", 240)))


def get_emb_eri_fast_gdf(cell, mydf, C_ao_lo=None, basis=None, feri=None, kscaled_center=None, symmetry=4,
                         max_memory=None, C_ao_eo=None, kconserv_tol=KPT_DIFF_TOL, unit_eri=False, swap_idx=None,
                         t_reversal_symm=True, incore=True, fout="H2.h5", kL_subset=None, restore=True, info=None):
    """eri_transform.py:235-399 (incore).  `kL_subset` / `restore=False` are oracle-only hooks for the multi-rank
    tests (the reference's MPI variant shards the same kL loop, eri_transform_mpi.py:151-157)."""
    assert incore, "oracle restates the incore branch only"
    nkpts = len(mydf.kpts_scaled)
    kscaled = np.array(mydf.kpts_scaled, dtype=float)
    if kscaled_center is not None:
        kscaled = kscaled - kscaled_center
    C_ao_emb = build_C_ao_emb(mydf, C_ao_lo, basis, C_ao_eo, unit_eri)
    spin, nemb = C_ao_emb.shape[0], C_ao_emb.shape[-1]
    npair = nemb * (nemb + 1) // 2
    # NOTE the weights come from the UNSHIFTED k-points in the reference (l.309)
    weights = get_weights_t_reversal(mydf.kpts_scaled) if t_reversal_symm else np.ones(nkpts, dtype=int)
    eri = np.zeros((spin * (spin + 1) // 2, npair, npair), dtype=float if t_reversal_symm else np.complex128)
    blksize = _chunk_rows(mydf, max_memory)
    for kL in range(nkpts):
        if weights[kL] <= 0 or (kL_subset is not None and kL not in kL_subset):
            continue
        Lij_s4 = _accumulate_Lij_s4(mydf, C_ao_emb, kL, kscaled, kconserv_tol, t_reversal_symm, blksize)
        _Lij_s4_to_eri(Lij_s4, eri, weight=weights[kL], t_reversal_symm=t_reversal_symm)
    if info is not None and not t_reversal_symm:
        info["eri_imag_norm"] = max_abs(eri.imag)          # what the reference logs / warns about (l.390-395)
    if not restore:
        return eri
    return eri_restore(eri.real, symmetry, nemb)


def get_emb_eri(cell, mydf, C_ao_lo=None, basis=None, unit_eri=False, symmetry=4, t_reversal_symm=True,
                max_memory=None, swap_idx=None, feri=None, kscaled_center=None, kconserv_tol=KPT_DIFF_TOL,
                incore=True, fout="H2.h5", **kwargs):
    """dispatcher, GDF branch only (eri_transform.py:44-94)"""
    return get_emb_eri_fast_gdf(cell, mydf, C_ao_lo=C_ao_lo, basis=basis, kscaled_center=kscaled_center,
                                symmetry=symmetry, max_memory=max_memory, kconserv_tol=kconserv_tol,
                                unit_eri=unit_eri, t_reversal_symm=t_reversal_symm, incore=incore)


def get_unit_eri(cell, mydf, C_ao_lo=None, symmetry=4, t_reversal_symm=True, max_memory=None, swap_idx=None,
                 feri=None, kscaled_center=None, kconserv_tol=KPT_DIFF_TOL, incore=True, fout="H2.h5", **kwargs):
    """ERI of the first cell's LOs: the same path with C_ao_emb = C_ao_lo / nkpts^(3/4) (eri_transform.py:96-112)"""
    C = np.asarray(C_ao_lo)
    C = C[None] if C.ndim == 3 else C
    return get_emb_eri(cell, mydf, C_ao_lo=C, basis=None, unit_eri=True, symmetry=symmetry,
                       t_reversal_symm=t_reversal_symm, max_memory=max_memory, kscaled_center=kscaled_center,
                       kconserv_tol=kconserv_tol)


def _Lij_s4_to_eri_gso(Lij_s4, eri, weight=1, t_reversal_symm=False):
    """GSO Gram products with signs (+aa, +bb, -ab, -ba)   (eri_transform.py:1252-1284, incore)"""
    signed = [(0, 0, 1.0), (1, 1, 1.0), (0, 1, -1.0), (1, 0, -1.0)]
    if not t_reversal_symm:
        for a, b, sgn in signed:
            eri[0] += sgn * np.dot(Lij_s4[a].conj().T, Lij_s4[b])
        return
    if weight not in (1, 2):
        raise ValueError
    for part in ([Lij_s4.real] if weight == 1 else [Lij_s4.real, Lij_s4.imag]):
        part = np.ascontiguousarray(part)
        for a, b, sgn in signed:
            lib.dot(part[a].T, part[b], sgn * weight, eri[0], 1)


def get_emb_eri_gso(cell, mydf, C_ao_lo=None, basis=None, feri=None, kscaled_center=None, symmetry=4,
                    max_memory=None, kconserv_tol=KPT_DIFF_TOL, unit_eri=False, swap_idx=None,
                    t_reversal_symm=True, basis_k=None, incore=True, fout="H2.h5"):
    """GSO embedding ERI with partial particle-hole transform (eri_transform.py:1104-1250, incore): two spin
    flavours = alpha / beta rows of the generalised-spin-orbital basis (separate_basis,
    libdmet/routine/spinless_helper.py:31-46), one ERI block."""
    assert incore
    nkpts = len(mydf.kpts_scaled)
    C_ao_lo = add_spin_dim(C_ao_lo, 2)
    kscaled = np.array(mydf.kpts_scaled, dtype=float)
    if kscaled_center is not None:
        kscaled = kscaled - kscaled_center
    if basis_k is None:
        assert basis is not None and basis.ndim == 3
        basis_k = get_basis_k(basis[None], get_phase_R2k_scaled(mydf.kmesh, mydf.kpts_scaled))[0]
    if basis_k.ndim == 3:
        half = basis_k.shape[1] // 2
        basis_k = np.asarray((basis_k[:, :half], basis_k[:, half:]))
    C_ao_emb = (C_ao_lo if unit_eri else multiply_basis(C_ao_lo, basis_k)) / (nkpts ** 0.75)
    nemb = C_ao_emb.shape[-1]
    npair = nemb * (nemb + 1) // 2
    weights = get_weights_t_reversal(mydf.kpts_scaled) if t_reversal_symm else np.ones(nkpts, dtype=int)
    eri = np.zeros((1, npair, npair), dtype=float if t_reversal_symm else np.complex128)
    blksize = _chunk_rows(mydf, max_memory)
    for kL in range(nkpts):
        if weights[kL] <= 0:
            continue
        Lij_s4 = _accumulate_Lij_s4(mydf, C_ao_emb, kL, kscaled, kconserv_tol, t_reversal_symm, blksize)
        _Lij_s4_to_eri_gso(Lij_s4, eri, weight=weights[kL], t_reversal_symm=t_reversal_symm)
    return eri_restore(eri.real, symmetry, nemb)


# ---------------------------------------------------------------------------------------------------------
# GDF tensor in the LO basis (eri_transform.py:1312-1427)
# ---------------------------------------------------------------------------------------------------------
def stored_pairs(mydf):
    """(k_i, k_j) index pairs a PySCF GDF file holds, in file order: j <= i (eri_transform.py:1352-1355 without
    band k-points); providers may carry their own list in `.kptij_idx`"""
    if hasattr(mydf, "kptij_idx"):
        return [tuple(int(x) for x in p) for p in mydf.kptij_idx]
    nk = len(mydf.kpts_scaled)
    return [(i, j) for i in range(nk) for j in range(i + 1)]


def get_mask_kptij_lst(kptij_scaled, tol=KPT_DIFF_TOL):
    """time-reversal partner of every stored pair: mask[a] = b > a when pair b = -pair a (b then gets -2 =
    "filled from its partner"), -1 for pairs that are their own partner or have none (eri_transform.py:1409-1427)"""
    flat = round_to_FBZ(np.asarray(kptij_scaled, dtype=float).reshape(len(kptij_scaled), -1), tol=tol)
    mask = np.full(len(flat), -1, dtype=int)
    for a in range(len(flat)):
        if mask[a] != -1:
            continue
        for b in range(a + 1, len(flat)):
            s = flat[a] + flat[b]
            if max_abs(s - np.round(s)) < tol:
                mask[a], mask[b] = b, -2
                break
    return mask


def transform_gdf_to_lo(mydf, C_ao_lo, t_reversal_symm=True, blksize=240):
    """{pair position: L_lo} in the reference's storage convention (eri_transform.py:1312-1407): for every stored
    pair, L_lo[L, m, n] = sum_pq conj(C_i[p, m]) L[L, p, q] C_j[q, n]; real lower-triangular packed when both
    k-points are Gamma, complex packed when k_i == k_j, full (naux, nlo*nlo) otherwise; with time reversal the
    partner pair receives the complex conjugate."""
    C = np.asarray(C_ao_lo)
    nk, nao, nlo = C.shape
    assert nk == len(mydf.kpts_scaled) and nao == mydf.nao
    pairs = stored_pairs(mydf)
    ks = np.asarray(mydf.kpts_scaled)
    mask = (get_mask_kptij_lst([np.concatenate([ks[i], ks[j]]) for i, j in pairs]) if t_reversal_symm
            else np.full(len(pairs), -1, dtype=int))
    out = {}
    for pos, (i, j) in enumerate(pairs):
        if mask[pos] == -2:
            continue
        Lij = np.zeros((mydf.naux, nlo * nlo), dtype=np.complex128)
        row = 0
        for Lpq in sr_loop(mydf, i, j, blksize):
            Lij[row:row + Lpq.shape[0]] = transform_ao_to_emb(Lpq, C, i, j)[0]
            row += Lpq.shape[0]
        both_gamma = max_abs(ks[i]) < KPT_DIFF_TOL and max_abs(ks[j]) < KPT_DIFF_TOL
        if both_gamma:
            assert max_abs(Lij.imag) < ERI_IMAG_TOL
            stored = lib.pack_tril(Lij.real.reshape(-1, nlo, nlo))
        elif i == j:
            stored = lib.pack_tril(Lij.reshape(-1, nlo, nlo))
        else:
            stored = Lij
        out[pos] = stored
        if mask[pos] != -1:
            out[int(mask[pos])] = stored.conj()
    return out


# ---------------------------------------------------------------------------------------------------------
# GDF tensor on disk (eri_transform.py:159-227)
# ---------------------------------------------------------------------------------------------------------
class FileGDF(object):
    """GDF provider over an open h5py-like cderi file, following the reference's own access pattern: `get_naoaux`
    (159-193: rows of `j3c/<k>` -- or of its segment '0' -- for every stored pair, naux = the maximum) and
    `sr_loop` (195-227: `_load3c`, Hermitian unpacking of k_i == k_j blocks, cast to complex128).  Auxiliary rows
    a pair does not have stay zero, as in the reference's accumulation buffers (l.365-378)."""

    def __init__(self, feri, cell, kpts):
        self.feri, self.cell = feri, cell
        self.kpts = np.asarray(kpts, dtype=float)
        self.kpts_scaled = cell.get_scaled_kpts(self.kpts)
        sk = self.kpts_scaled.round(8)
        self.kmesh = [len(np.unique(sk[:, d])) for d in range(3)]
        self.nao = int(cell.nao_nr())
        self.blockdim = 240
        if "j3c-kptij" in feri:                                                   # l.164-168
            keys = ["j3c/%d" % k for k in range(feri["j3c-kptij"].shape[0])]
        else:
            keys = ["j3c/%s" % k for k in feri["j3c"].keys()]
        rows = []
        for key in keys:                                                          # l.170-176
            entry = feri[key]
            rows.append(entry["0"].shape[0] if hasattr(entry, "keys") else entry.shape[0])
        self.naux = int(max(rows))

    def load(self, ki, kj):
        nao = self.nao
        j3c = lib.load3c(self.feri, "j3c", self.kpts[[ki, kj]], "j3c-kptij", nao)
        Lpq = np.asarray(j3c[0:j3c.shape[0]])
        if ki == kj and Lpq.shape[-1] != nao * nao:                               # l.201, 216-217
            Lpq = lib.unpack_tril(Lpq).reshape(-1, nao * nao)
        out = np.zeros((self.naux, nao, nao), dtype=np.complex128)
        out[:Lpq.shape[0]] = np.asarray(Lpq, dtype=np.complex128).reshape(-1, nao, nao)
        return out
