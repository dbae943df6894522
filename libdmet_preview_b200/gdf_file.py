"""On-disk GDF tensors: the PySCF `cderi` file layouts the reference consumes, served as a GDF provider.

What the reference does with the file (`libdmet/basis_transform/eri_transform.py`):
    get_naoaux  (159-193)  rows of `j3c/<pair>[/0]` for every stored pair; naux = the maximum ("aux basis drop")
    sr_loop     (195-227)  PySCF `_load3c(cderi, 'j3c', kpti_kptj, 'j3c-kptij')`: the stored (k_i, k_j) entry, or the
                           conjugate transpose of the stored (k_j, k_i) one; column segments `j3c/<pair>/<seg>` are
                           concatenated; Hermitian-packed (s2) k_i == k_j blocks are unpacked (216-217); everything
                           is cast to complex128 (218)
Two generations of the layout exist and both are read here:
    v1   `j3c-kptij` (npairs, 2, 3) absolute k-point pairs, `j3c/<position in that list>/<seg>`
    v2   `kpts` (nkpts, 3), `aosym` ('s1' | 's2'), `j3c/<k_i * nkpts + k_j>/<seg>`      (PySCF >= 2.1)
(v1 is what the reference itself relies on -- it passes 'j3c-kptij' to `_load3c` -- and is pinned by golden results of
the reference's own code, tests/golden/eri_file_*.npz; v2 follows PySCF's `CDERIArray` key convention and is checked
against this module's own writer only.)
A `j3c/<pair>` entry may also be a plain dataset instead of a group of segments (old PySCF, and what
`transform_gdf_to_lo` writes: `j3c/<pair>/0` only).

The file is parsed by `h5lite` (h5py is used instead when the file needs a feature outside that subset and h5py is
importable).  Block payloads of full, single-segment, correctly oriented pairs are read by one positioned read
straight into the destination buffer; packed, transposed and multi-segment blocks are assembled on the host the way
`_load3c` does.
"""
import numpy as np

from . import h5lite
from .schedule import KPT_DIFF_TOL

_KPT_FILE_TOL = 1e-6     # PySCF kpts_helper.member tolerance (KPT_DIFF_TOL)


STORED_SWAPPED = 1       # include/ldm_b200.h: LDM_STORED_SWAPPED
STORED_REAL = 2          # LDM_STORED_REAL
STORED_CONJ = 4          # LDM_STORED_CONJ: entry of the time-reversed pair (-k_i, -k_j), plain conjugate


class StoredEntry(object):
    """rows of one cderi entry as they lie in the file + the flags the device unpacker needs"""
    __slots__ = ("data", "flags")

    def __init__(self, data, flags):
        self.data, self.flags = data, flags

    def expand(self, naux, nao):
        """what `ldm_unpack_stored` produces from this entry, (naux, nao, nao) complex128 -- layout only, used by the
        tests to check the device kernel and the stored-entry bookkeeping"""
        out = np.zeros((naux, nao, nao), dtype=np.complex128)
        a = self.data
        rows = a.shape[0]
        if a.shape[1] == nao * nao:
            full = a.reshape(rows, nao, nao).astype(np.complex128)
        else:
            r, c = np.tril_indices(nao)
            full = np.empty((rows, nao, nao), dtype=np.complex128)
            full[:, c, r] = a.conj()
            full[:, r, c] = a
        if self.flags & STORED_SWAPPED:
            # full entry of the pair (k_j, k_i): conjugate transpose; packed one: plain conjugate (PySCF _load3c)
            full = full.conj().transpose(0, 2, 1) if a.shape[1] == nao * nao else full.conj()
        if self.flags & STORED_CONJ:
            full = full.conj()
        out[:rows] = full
        return out


def _open(path):
    try:
        return h5lite.File(path)
    except h5lite.H5FormatError:
        try:
            import h5py
        except ImportError:
            raise
        return h5py.File(path, "r")


def _segments(entry):
    """datasets of one stored pair in column order: a bare dataset, or the members '0', '1', ... of a group"""
    if hasattr(entry, "keys"):
        return [entry[str(n)] for n in range(len(entry))]
    return [entry]


def _match(k, kpts, tol=_KPT_FILE_TOL):
    hit = np.flatnonzero(np.abs(np.asarray(kpts) - np.asarray(k)).reshape(len(kpts), -1).max(axis=1) < tol)
    return int(hit[0]) if len(hit) else -1


class GDFFile(object):
    """GDF provider over a PySCF cderi file: `.kpts .kpts_scaled .kmesh .nao .naux .blockdim .kptij_idx
    .load(ki, kj)` (the duck type `get_emb_eri`, `transform_gdf_to_lo`, `ResidentGDF` and the oracle accept).

    cell            anything with `lattice_vectors()` (and `nao_nr()`, `dimension` when present); or None with
                    `lattice_vectors=` given
    kpts            absolute k-points in the caller's order (mydf.kpts); default: the file's own order
    """

    fills_out = True         # `load(ki, kj, out=buffer)` writes into a caller buffer (pinned staging ring of the pipeline)

    def __init__(self, path, cell=None, kpts=None, lattice_vectors=None, label="j3c"):
        self.path = self._cderi = path
        self.cell = cell
        self.label = label
        self.blockdim = 240
        self.max_memory = 4000
        self._f = _open(path)
        f = self._f
        if label not in f:
            raise KeyError("%s holds no '%s' group: not a GDF cderi file" % (path, label))
        if cell is not None and getattr(cell, "dimension", 3) == 2 \
                and getattr(cell, "low_dim_ft_type", None) != "inf_vacuum":
            raise NotImplementedError("2-D cells with a negative-definite j3c- part (eri_transform.py:226-227)")
        # -- which pairs are stored, under which key
        if "j3c-kptij" in f:
            self.version = "v1"
            kptij = np.asarray(f["j3c-kptij"][...], dtype=float).reshape(-1, 2, 3)
            if kpts is None:
                diag = [n for n in range(len(kptij)) if np.abs(kptij[n, 0] - kptij[n, 1]).max() < _KPT_FILE_TOL]
                kpts = kptij[diag, 0]
            kpts = np.asarray(kpts, dtype=float).reshape(-1, 3)
            idx = []
            for n in range(len(kptij)):
                i, j = _match(kptij[n, 0], kpts), _match(kptij[n, 1], kpts)
                if i < 0 or j < 0:
                    raise ValueError("%s: stored pair %d is not on the k-mesh" % (path, n))
                idx.append((i, j))
            keys = [str(n) for n in range(len(kptij))]
        elif "kpts" in f:
            self.version = "v2"
            fk = np.asarray(f["kpts"][...], dtype=float).reshape(-1, 3)
            kpts = fk if kpts is None else np.asarray(kpts, dtype=float).reshape(-1, 3)
            perm = [_match(k, kpts) for k in fk]                  # file index -> caller index
            if min(perm) < 0 or len(fk) != len(kpts):
                raise ValueError("%s: the file's k-points differ from mydf.kpts" % path)
            idx, keys = [], []
            for key in sorted(f[label].keys(), key=int):
                a, b = divmod(int(key), len(fk))
                idx.append((perm[a], perm[b]))
                keys.append(key)
        else:
            raise KeyError("%s has neither 'j3c-kptij' nor 'kpts'" % path)
        self.kpts = kpts
        self.nkpts = len(kpts)
        self.kptij_idx = idx
        self._key = dict(zip(idx, keys))
        self._minus = None
        # -- geometry
        if lattice_vectors is None:
            if cell is None:
                raise ValueError("GDFFile needs a cell or lattice_vectors to scale the k-points")
            lattice_vectors = cell.lattice_vectors()
        a = np.asarray(lattice_vectors, dtype=float)
        self.kpts_scaled = np.dot(kpts, a.T) / (2.0 * np.pi)         # cell.get_scaled_kpts
        sk = self.kpts_scaled.round(8)
        self.kmesh = [len(np.unique(sk[:, d])) for d in range(3)]    # fourier.py:83-89
        # -- sizes: get_naoaux (159-193) and nao from the column count of an unpacked pair
        rows, cols_full = [], None
        for (i, j) in idx:
            segs = _segments(f[label][self._key[(i, j)]])
            rows.append(int(segs[0].shape[0]))
            if i != j and cols_full is None:
                cols_full = sum(int(s.shape[1]) for s in segs)
        self.naux_of = dict(zip(idx, rows))
        self.naux = max(rows)
        self.aux_drop = len(set(rows)) != 1                          # the reference warns here (l.191-192)
        if cell is not None and hasattr(cell, "nao_nr"):
            self.nao = int(cell.nao_nr())
        elif cols_full is not None:
            self.nao = int(round(np.sqrt(cols_full)))
        else:
            ncol = sum(int(s.shape[1]) for s in _segments(f[label][keys[0]]))
            n2 = int(round(np.sqrt(ncol)))
            self.nao = n2 if n2 * n2 == ncol else int(round((np.sqrt(8 * ncol + 1) - 1) / 2))
        self._tril = None

    def _resolve(self, ki, kj):
        """(stored pair, flags) serving block (k_i, k_j): the pair itself; the swapped pair (conjugate transpose,
        PySCF `_load3c`); or -- files that keep only one member of each time-reversal pair, as recent PySCF writes
        them -- the pair (-k_i, -k_j) (plain conjugate, L(-k_i, -k_j) = conj L(k_i, k_j)) or (-k_j, -k_i)."""
        if (ki, kj) in self._key:
            return (ki, kj), 0
        if (kj, ki) in self._key:
            return (kj, ki), STORED_SWAPPED
        if self._minus is None:
            ks = np.asarray(self.kpts_scaled)
            self._minus = []
            for k in ks:
                d = ks + k[None]
                hit = np.flatnonzero(np.abs(d - np.round(d)).max(axis=1) < KPT_DIFF_TOL)
                self._minus.append(int(hit[0]) if len(hit) else -1)
        mi, mj = self._minus[ki], self._minus[kj]
        if mi >= 0 and mj >= 0:
            if (mi, mj) in self._key:
                return (mi, mj), STORED_CONJ
            if (mj, mi) in self._key:
                return (mj, mi), STORED_SWAPPED | STORED_CONJ
        raise KeyError("k-point pair (%d, %d) is not stored in %s" % (ki, kj, self.path))

    # -- block assembly (PySCF _load3c / _KPair3CLoader + sr_loop's unpack and cast) ---------------------------
    def _stored(self, key, out):
        """read the stored entry `key` into out[:rows] ((naux, nao, nao) complex128, rows beyond the stored ones
        are left untouched); returns (rows, packed?)"""
        nao = self.nao
        segs = _segments(self._f[self.label][key])
        rows = int(segs[0].shape[0])
        ncol = sum(int(s.shape[1]) for s in segs)
        flat = out.reshape(out.shape[0], nao * nao)
        if ncol == nao * nao:
            if len(segs) == 1 and segs[0].dtype == np.complex128 and hasattr(segs[0], "read_direct") \
                    and rows == out.shape[0]:
                segs[0].read_direct(flat)                       # one positioned read, no intermediate copy
            else:
                c0 = 0
                for s in segs:
                    w = int(s.shape[1])
                    flat[:rows, c0:c0 + w] = s[...]
                    c0 += w
            return rows, False
        if ncol != nao * (nao + 1) // 2:
            raise ValueError("%s: pair %s has %d columns, expected %d or %d" % (self.path, key, ncol, nao * nao,
                                                                              nao * (nao + 1) // 2))
        packed = np.concatenate([np.asarray(s[...]) for s in segs], axis=1) if len(segs) > 1 \
            else np.asarray(segs[0][...])
        if self._tril is None:
            self._tril = np.tril_indices(nao)
        r, c = self._tril
        v = out[:rows]
        v[:, c, r] = packed.conj()                              # unpack_tril, HERMITIAN fill (l.216-217)
        v[:, r, c] = packed
        return rows, True

    def load_stored(self, ki, kj, l0, l1, buf):
        """Rows [l0, l1) of the entry that serves block (k_i, k_j), AS STORED, read into the byte storage of `buf`
        (any C-contiguous array of at least naux * nao * nao * 16 bytes, e.g. a pinned staging buffer).  Returns a
        `StoredEntry` whose `.data` is the (rows, ncols) view and `.flags` say how the device has to interpret it
        (`ldm_eri_block_stored`).  Single-segment entries are one positioned read; column segments are read one
        after the other into their column ranges."""
        pair, flags = self._resolve(ki, kj)
        key = self._key[pair]
        segs = _segments(self._f[self.label][key])
        dtype = np.dtype(segs[0].dtype)
        if dtype not in (np.dtype(np.complex128), np.dtype(np.float64)):
            raise TypeError("%s: entry %s has dtype %s" % (self.path, key, dtype))
        if dtype == np.float64:
            flags |= STORED_REAL
        nrow = int(segs[0].shape[0])
        ncol = sum(int(s.shape[1]) for s in segs)
        nao = self.nao
        if ncol not in (nao * nao, nao * (nao + 1) // 2):
            raise ValueError("%s: pair %s has %d columns, expected %d or %d" % (self.path, key, ncol, nao * nao,
                                                                              nao * (nao + 1) // 2))
        r0, r1 = min(l0, nrow), min(l1, nrow)
        raw = buf.reshape(-1).view(np.uint8)
        data = raw[:(r1 - r0) * ncol * dtype.itemsize].view(dtype).reshape(r1 - r0, ncol)
        if r1 > r0:
            if len(segs) == 1 and hasattr(segs[0], "read_direct"):
                segs[0].read_direct(data, slice(r0, r1))
            else:
                c0 = 0
                for s in segs:
                    w = int(s.shape[1])
                    data[:, c0:c0 + w] = s[r0:r1]
                    c0 += w
        return StoredEntry(data, flags)

    def load(self, ki, kj, out=None):
        """(naux, nao, nao) complex128 block L(k_i, k_j); auxiliary rows a pair does not have are zero.
        `out`: optional destination (e.g. a pinned buffer) of that shape."""
        nao = self.nao
        if out is None:
            out = np.empty((self.naux, nao, nao), dtype=np.complex128)
        pair, flags = self._resolve(ki, kj)
        if not flags & STORED_SWAPPED:
            rows, _ = self._stored(self._key[pair], out)
        else:
            tmp = np.empty((self.naux_of[pair], nao, nao), dtype=np.complex128)
            rows, _ = self._stored(self._key[pair], tmp)
            np.conjugate(tmp.transpose(0, 2, 1), out=out[:rows])
        if flags & STORED_CONJ:
            np.conjugate(out[:rows], out=out[:rows])
        if rows < out.shape[0]:
            out[rows:] = 0.0
        return out

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def write_gdf_file(path, provider, version="v1", nsegments=1, pack_diagonal=True, real_gamma=True, pairs=None,
                   naux_of=None):
    """Write the GDF tensor of `provider` as a PySCF-layout cderi file (module header).

    pairs           stored (k_i, k_j) index pairs; default j <= i for every i (PySCF's list)
    nsegments       column segments per pair (`j3c/<pair>/0 .. nsegments-1`)
    pack_diagonal   store k_i == k_j blocks Hermitian-packed (lower triangle, row-major), as PySCF's aosym 's2'
    real_gamma      store the (Gamma, Gamma) block as float64 when its imaginary part vanishes
    naux_of         {(k_i, k_j): rows} to store fewer auxiliary rows for some pairs (PySCF drops linearly dependent
                    auxiliary functions per k-point pair)
    """
    nk = len(provider.kpts_scaled)
    nao = int(provider.nao)
    kpts = np.asarray(provider.kpts, dtype=float)
    if pairs is None:
        pairs = [(i, j) for i in range(nk) for j in range(i + 1)]
    r, c = np.tril_indices(nao)
    with h5lite.Writer(path) as w:
        if version == "v1":
            w["j3c-kptij"] = np.asarray([(kpts[i], kpts[j]) for i, j in pairs])
        elif version == "v2":
            w["kpts"] = kpts
            w["aosym"] = "s2" if pack_diagonal else "s1"
        else:
            raise ValueError("unknown cderi layout %s" % version)
        for pos, (i, j) in enumerate(pairs):
            blk = np.asarray(provider.load(i, j), dtype=np.complex128).reshape(-1, nao, nao)
            if naux_of is not None and (i, j) in naux_of:
                blk = blk[:naux_of[(i, j)]]
            if i == j and pack_diagonal:
                data = np.ascontiguousarray(blk[:, r, c])
            else:
                data = blk.reshape(len(blk), nao * nao)
            gamma = max(np.abs(provider.kpts_scaled[i]).max(), np.abs(provider.kpts_scaled[j]).max()) < KPT_DIFF_TOL
            if real_gamma and gamma and not np.any(data.imag):
                data = np.ascontiguousarray(data.real)
            key = str(pos) if version == "v1" else str(i * nk + j)
            bounds = np.linspace(0, data.shape[1], nsegments + 1).astype(int)
            for s in range(nsegments):
                w["j3c/%s/%d" % (key, s)] = data[:, bounds[s]:bounds[s + 1]]
    return path
