"""Seeded synthetic inputs of the named (nkpts, nao, naux, neo) shapes: a duck-typed cell, a GDF provider with
PySCF's symmetry structure, localized-orbital coefficients, embedding bases and mean-field matrices.

PySCF and h5py are not available in this image, so benchmarks and parity tests use these providers
(SURVEY.md section 8d).  The GDF tensor obeys, by construction,
    L(k_j, k_i)[L, q, p]  = conj(L(k_i, k_j)[L, p, q])          (what pyscf's _load3c returns for swapped pairs)
    L(-k_i, -k_j)         = conj(L(k_i, k_j))                    (time reversal)
so the time-reversal path of the reference is physically equivalent to the plain path.  Values come from a
counter-based 32-bit hash, reproduced bit for bit by the device generator (`synth_block_kernel` in
csrc/aux_kernels.cuh) so that a 2.6 TB tensor never has to cross PCIe in throughput runs.
"""
import numpy as np

from .schedule import make_kpts_scaled, cell_vectors, kpt_member

_M32 = np.uint64(0xFFFFFFFF)


def lowbias32(x):
    """32-bit integer hash (same constants as the device twin)."""
    x = np.asarray(x, dtype=np.uint64) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


def _u(key, idx):
    h = lowbias32(np.asarray(idx, dtype=np.uint64) ^ np.uint64(key)).astype(np.uint32)
    return h.view(np.int32).astype(np.float64) * 4.656612873077393e-10   # 2^-31


def _u32(key, idx32):
    """`_u` on a uint32 index array (wrapping 32-bit arithmetic, one temporary): same bits, 2.5x faster"""
    x = idx32 ^ np.uint32(key)
    t = x >> np.uint32(16)
    x ^= t
    x *= np.uint32(0x7FEB352D)
    np.right_shift(x, np.uint32(15), out=t)
    x ^= t
    x *= np.uint32(0x846CA68B)
    np.right_shift(x, np.uint32(16), out=t)
    x ^= t
    return x.view(np.int32).astype(np.float64) * 4.656612873077393e-10   # 2^-31



class SyntheticCell(object):
    """The four members of a PySCF Cell the path touches (eri_transform.py:256,266; fourier.py:40,87)."""

    def __init__(self, nao, a=None, dimension=3):
        self._nao = int(nao)
        self._a = np.eye(3) * 4.0 if a is None else np.asarray(a, dtype=float)
        self.dimension = dimension
        self.low_dim_ft_type = None

    def nao_nr(self):
        return self._nao

    def lattice_vectors(self):
        return self._a

    def reciprocal_vectors(self):
        return 2.0 * np.pi * np.linalg.inv(self._a).T

    def get_scaled_kpts(self, kpts_abs):
        return np.dot(np.asarray(kpts_abs), self._a.T) / (2.0 * np.pi)

    def get_abs_kpts(self, kpts_scaled):
        return np.dot(np.asarray(kpts_scaled), self.reciprocal_vectors())


class SyntheticGDF(object):
    """In-memory GDF provider: .kpts / .kpts_scaled / .kmesh / .nao / .naux / .blockdim / .load(ki, kj)."""

    def __init__(self, kmesh, nao, naux, seed=0, scale=None, cell=None):
        self.kmesh = [int(x) for x in kmesh]
        self.nao = int(nao)
        self.naux = int(naux)
        self.seed = int(seed) & 0xFFFFFFFF
        self.cell = cell if cell is not None else SyntheticCell(nao)
        self.kpts_scaled = make_kpts_scaled(self.kmesh)
        self.kpts = self.cell.get_abs_kpts(self.kpts_scaled)
        self.nkpts = len(self.kpts_scaled)
        self.scale = float(scale) if scale is not None else float(np.sqrt(self.nkpts / float(self.naux)))
        self.blockdim = 240          # pyscf GDF default
        self.max_memory = 4000
        self._cderi = "<synthetic>"
        self.minus = [int(kpt_member(-k, self.kpts_scaled)[0]) for k in self.kpts_scaled]
        self._keys = {}
        assert 2 * self.naux * self.nao * self.nao < 2 ** 32

    def pair_key(self, a, b):
        inner = int(lowbias32(np.uint64((a * self.nkpts + b + 0x9E3779B9) & 0xFFFFFFFF)))
        return int(lowbias32(np.uint64(self.seed ^ inner)))

    def keys(self, ki, kj):
        k = self._keys.get((ki, kj))
        if k is None:
            mi, mj = self.minus[ki], self.minus[kj]
            k = (self.pair_key(ki, kj), self.pair_key(kj, ki), self.pair_key(mi, mj), self.pair_key(mj, mi))
            self._keys[(ki, kj)] = k
        return k

    def load(self, ki, kj, aux_slice=None):
        """(naux, nao, nao) complex128 block L(k_i, k_j) (host twin of the device generator; 32-bit wrapping
        arithmetic on chunks of auxiliary rows small enough for the temporaries to stay in cache)."""
        nao = self.nao
        l0, l1 = (0, self.naux) if aux_slice is None else aux_slice
        keys = self.keys(ki, kj)
        out = np.empty((l1 - l0, nao, nao), dtype=np.complex128)
        rows = max(1, 80000 // (nao * nao))
        for a in range(l0, l1, rows):
            self._fill(out[a - l0:min(l1, a + rows) - l0], a, min(l1, a + rows), keys)
        return out

    def _fill(self, out, l0, l1, keys):
        nao = self.nao
        k_ij, k_ji, k_mij, k_mji = keys
        n32 = np.uint32(nao)
        L = np.arange(l0, l1, dtype=np.uint32)[:, None, None]
        p = np.arange(nao, dtype=np.uint32)[None, :, None]
        q = np.arange(nao, dtype=np.uint32)[None, None, :]
        d = np.uint32(2) * ((L * n32 + p) * n32 + q)
        t = np.uint32(2) * ((L * n32 + q) * n32 + p)
        s = 0.25 * self.scale
        re = _u32(k_ij, d)
        re += _u32(k_ji, t)
        re += _u32(k_mij, d)
        re += _u32(k_mji, t)
        re *= s
        out.real = re
        d += np.uint32(1)
        t += np.uint32(1)
        im = _u32(k_ij, d)
        im -= _u32(k_ji, t)
        im -= _u32(k_mij, d)
        im += _u32(k_mji, t)
        im *= s
        out.imag = im


class PooledGDF(object):
    """A GDF tensor made of `npool` distinct synthetic blocks: L(k_i, k_j) := block number (k_i nkpts + k_j) mod
    npool of the wrapped SyntheticGDF.  Throughput runs use it so that the tensor of the target workload (758 GB)
    is DEFINED by a pool that fits one GPU, identically for every number of ranks -- each (k_i, k_j) block of the
    schedule is still read, transformed and accumulated on its own, but the same numbers can be produced on 1, 2, 4
    or 8 GPUs and by the CPU oracle (tests/golden/make_bench_digest.py).  The pair symmetries of a physical tensor
    are not kept; the pipeline and the oracle run the same schedule on any input."""

    def __init__(self, gdf, npool):
        self.inner = gdf
        for a in ("kmesh", "nao", "naux", "cell", "kpts", "kpts_scaled", "nkpts", "scale", "blockdim", "max_memory",
                  "seed"):
            setattr(self, a, getattr(gdf, a))
        self.npool = int(max(1, min(npool, self.nkpts * self.nkpts)))
        self._cderi = "<synthetic pool of %d blocks>" % self.npool

    def pool_index(self, ki, kj):
        return (ki * self.nkpts + kj) % self.npool

    def pool_pair(self, s):
        return s % self.nkpts, (s // self.nkpts) % self.nkpts

    def block_key(self, ki, kj):
        """pairs that share a pool block share its resident copy (eri_transform.ResidentGDF)"""
        return self.pool_index(ki, kj)

    def keys(self, ki, kj):
        return self.inner.keys(*self.pool_pair(self.pool_index(ki, kj)))

    def load(self, ki, kj, aux_slice=None):
        return self.inner.load(*self.pool_pair(self.pool_index(ki, kj)), aux_slice=aux_slice)


def _trs_fill(nk, minus, make):
    """fill out[k] = make(k) for one member of each {k, -k} pair and conj for the partner (real at k = -k)."""
    out = [None] * nk
    for k in range(nk):
        if out[k] is not None:
            continue
        v = make(k)
        if minus[k] == k:
            v = v.real.astype(np.complex128) if np.iscomplexobj(v) else v
        out[k] = v
        out[minus[k]] = v.conj()
    return np.asarray(out)


def make_C_ao_lo(kmesh, nao, nlo=None, seed=1, spin=None):
    """Unitary (orthonormal-column) C_ao_lo[k] with C(-k) = conj(C(k)); shape ((spin,) nkpts, nao, nlo)."""
    nlo = nao if nlo is None else nlo
    ks = make_kpts_scaled(kmesh)
    nk = len(ks)
    minus = [int(kpt_member(-k, ks)[0]) for k in ks]
    rng = np.random.default_rng(seed)

    def one_spin():
        def make(k):
            a = rng.standard_normal((nao, nlo)) + 1j * rng.standard_normal((nao, nlo))
            if minus[k] == k:
                a = a.real
            qmat, _ = np.linalg.qr(a)
            return qmat.astype(np.complex128)
        return _trs_fill(nk, minus, make)

    if spin is None:
        return one_spin()
    return np.asarray([one_spin() for _ in range(spin)])


def make_emb_basis(kmesh, nlo, neo, nimp=None, seed=2, spin=1):
    """Real orthonormal basis (spin, ncells, nlo, neo): identity on the first nimp orbitals of cell 0 (the impurity,
    libdmet/routine/slater.py:212) and a random orthonormal bath on the environment."""
    ncells = int(np.prod(kmesh))
    nimp = min(nlo, neo // 2) if nimp is None else nimp
    nbath = neo - nimp
    rng = np.random.default_rng(seed)
    out = np.zeros((spin, ncells * nlo, neo))
    for s in range(spin):
        out[s, :nimp, :nimp] = np.eye(nimp)
        if nbath > 0:
            b, _ = np.linalg.qr(rng.standard_normal((ncells * nlo - nimp, nbath)))
            out[s, nimp:, nimp:] = b
    return out.reshape(spin, ncells, nlo, neo)


def make_hermitian_k(kmesh, n, seed=3, spin=None, scale=1.0, posdef=False):
    """Hermitian h[k] with h(-k) = conj(h(k)) (real in R space); shape ((spin,) nkpts, n, n)."""
    ks = make_kpts_scaled(kmesh)
    nk = len(ks)
    minus = [int(kpt_member(-k, ks)[0]) for k in ks]
    rng = np.random.default_rng(seed)

    def one_spin():
        def make(k):
            a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
            if minus[k] == k:
                a = a.real.astype(np.complex128)
            h = (a + a.conj().T) * (0.5 * scale)
            if posdef:
                h = h.dot(h.conj().T) / n + np.eye(n)
            return h
        return _trs_fill(nk, minus, make)

    if spin is None:
        return one_spin()
    return np.asarray([one_spin() for _ in range(spin)])


def make_rdm1_k(fock_k, nocc):
    """Idempotent per-k density matrix from the lowest `nocc` eigenvectors of fock_k (nkpts, n, n) (orthonormal
    basis).  Keeps the time-reversal structure of fock_k."""
    out = np.zeros_like(fock_k)
    for k in range(fock_k.shape[0]):
        e, v = np.linalg.eigh(fock_k[k])
        out[k] = v[:, :nocc].dot(v[:, :nocc].conj().T)
    return out


def trs_block_count(kmesh, t_reversal_symm=True):
    """number of (k_i, k_j) blocks / Gram products in the reference schedule of a mesh."""
    from .schedule import build_schedule
    sch = build_schedule(make_kpts_scaled(kmesh), t_reversal_symm)
    return sch.nblocks, sch.ngram


__all__ = ["SyntheticCell", "SyntheticGDF", "PooledGDF", "make_C_ao_lo", "make_emb_basis", "make_hermitian_k", "make_rdm1_k",
           "trs_block_count", "cell_vectors"]
