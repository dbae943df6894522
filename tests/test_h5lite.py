"""CPU: the HDF5-subset reader/writer (libdmet_preview_b200/h5lite.py).

The reader is pinned against a file written by the real HDF5 library: scipy ships `testhdf5_7.4_GLNX86.mat`, a
MATLAB v7.3 file (= HDF5 with a 512-byte user block, superblock 0, symbol-table root group, version-1 object header,
version-2 contiguous layout) and the same variable as a classic .mat file readable by scipy.io.  The writer is checked
against the reader (round trips of every structure a cderi file uses)."""
import os
import struct

import numpy as np
import pytest

from libdmet_preview_b200 import h5lite


def _scipy_file(name):
    import scipy.io
    return os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", name)


def test_reads_a_file_written_by_libhdf5():
    import scipy.io
    path = _scipy_file("testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's HDF5 sample file is not installed")
    want = scipy.io.loadmat(_scipy_file("testdouble_7.4_GLNX86.mat"))["testdouble"]
    with h5lite.File(path) as f:
        assert f._base == 512 and f.keys() == ["testdouble"] and "testdouble" in f and "nothing" not in f
        ds = f["testdouble"]
        assert ds.shape == (9, 1) and ds.dtype == np.float64          # MATLAB stores column-major
        assert np.array_equal(ds[...].T, want)
        assert np.array_equal(ds[2:5], want.T[2:5])
        assert np.array_equal(ds.memmap(), want.T)
        off, n = ds.file_extent()
        with open(path, "rb") as raw:
            raw.seek(off)
            assert np.array_equal(np.frombuffer(raw.read(n), "<f8"), want.ravel())
        buf = np.empty((4, 1))
        ds.read_direct(buf, slice(5, 9))
        assert np.array_equal(buf, want.T[5:9])
        with pytest.raises(ValueError):
            ds.read_direct(np.empty((3, 1)), slice(5, 9))


def test_writer_emits_the_same_float_type_message_as_libhdf5():
    """byte-for-byte: IEEE f64 datatype message of the library-written sample vs ours"""
    path = _scipy_file("testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's HDF5 sample file is not installed")
    with h5lite.File(path) as f:
        msgs = dict(f._messages(f["testdouble"]._addr))
    assert bytes(msgs[0x03][:20]) == h5lite._encode_dtype(np.float64)


def test_not_hdf5(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(h5lite.H5FormatError):
        h5lite.File(str(p))
    with pytest.raises(ValueError):
        h5lite.File(str(p), "w")


def test_unsupported_features_fail_by_name(tmp_path):
    """a superblock of a version this reader does not know must not be mis-parsed"""
    p = tmp_path / "v9.h5"
    p.write_bytes(h5lite.SIGNATURE + bytes([9]) + bytes(200))
    with pytest.raises(h5lite.H5FormatError, match="superblock version 9"):
        h5lite.File(str(p))


@pytest.mark.parametrize("nmembers", [1, 8, 9, 300, 2100])
def test_group_btree_depths(tmp_path, nmembers):
    """1 symbol-table node; a full one; two; two B-tree levels (> 256 members); three levels (> 8192 would be needed
    for a fourth) -- 2080 is the number of stored pairs of a 4x4x4 mesh"""
    p = str(tmp_path / "g.h5")
    rng = np.random.default_rng(nmembers)
    vals = rng.standard_normal(nmembers)
    with h5lite.Writer(p) as w:
        for n in range(nmembers):
            w["j3c/%d/0" % n] = vals[n:n + 1]
    with h5lite.File(p) as f:
        g = f["j3c"]
        assert len(g) == nmembers and sorted(g.keys(), key=int) == [str(n) for n in range(nmembers)]
        assert g.keys() == sorted(g.keys())                     # B-tree order = byte order of the names
        for n in (0, nmembers // 2, nmembers - 1):
            assert f["j3c/%d/0" % n][0] == vals[n] and f["/j3c"][str(n)]["0"][...][0] == vals[n]
        with pytest.raises(KeyError):
            f["j3c/%d" % nmembers]
        ntype, level, used, _ = f._btree_node(g._btree)
        assert ntype == 0 and level == (0 if nmembers <= 256 else 1 if nmembers <= 8192 else 2)


def test_dtypes_shapes_and_partial_reads(tmp_path):
    p = str(tmp_path / "d.h5")
    rng = np.random.default_rng(5)
    data = {
        "c128": rng.standard_normal((7, 3, 3)) + 1j * rng.standard_normal((7, 3, 3)),
        "c64": (rng.standard_normal((4, 2)) + 1j * rng.standard_normal((4, 2))).astype(np.complex64),
        "f64": rng.standard_normal((5, 2, 3)),
        "f32": rng.standard_normal(6).astype(np.float32),
        "i32": np.arange(-5, 5, dtype=np.int32),
        "u8": np.arange(200, dtype=np.uint8).reshape(20, 10),
        "i64": np.asarray(-(1 << 40)),
        "bytes": np.asarray(b"s2"),
        "empty": np.zeros((0, 4)),
        "bigendian": np.arange(4, dtype=">f8"),
        "noncontig": np.arange(24.0).reshape(4, 6)[:, ::2],
    }
    with h5lite.Writer(p) as w:
        for k, v in data.items():
            w["grp/" + k] = v
        w["aosym"] = "s2"
        with pytest.raises(ValueError):
            w["grp/f64"] = data["f64"]
        with pytest.raises(ValueError):
            w["grp/f64/below"] = data["f64"]
    with h5lite.File(p) as f:
        assert f["aosym"][()] == "s2"
        for k, v in data.items():
            d = f["grp/" + k]
            got = d[...]
            assert got.shape == v.shape and np.array_equal(got, v), k
            assert got.dtype == (v.dtype.newbyteorder("<") if v.dtype.byteorder == ">" else v.dtype), k
            assert d.nbytes == v.nbytes and d.ndim == v.ndim
        c = f["grp/c128"]
        assert np.array_equal(c[2:5], data["c128"][2:5]) and np.array_equal(c[-2:], data["c128"][-2:])
        assert np.array_equal(c[1:6, 1], data["c128"][1:6, 1]) and np.array_equal(c[3], data["c128"][3])
        assert np.array_equal(np.asarray(c), data["c128"]) and len(c) == 7
        assert np.array_equal(c.memmap(), data["c128"])
        out = np.zeros((3, 3, 3), dtype=np.complex128)
        c.read_direct(out, slice(4, 7))
        assert np.array_equal(out, data["c128"][4:7])
        with pytest.raises(ValueError):
            c.read_direct(np.zeros((3, 3, 6), dtype=np.complex128)[:, :, ::2], slice(4, 7))     # not contiguous
        assert f["grp/empty"].file_extent() is None and f["grp/empty"].memmap().shape == (0, 4)


def test_large_payload_alignment_and_eof(tmp_path):
    """payloads of 1 MiB and more start on a 4 KiB boundary (direct I/O friendly); the end-of-file address in the
    superblock equals the file size"""
    p = str(tmp_path / "big.h5")
    a = np.random.default_rng(0).standard_normal((300, 500)) + 0j
    with h5lite.Writer(p) as w:
        w["j3c/0/0"] = a
    with h5lite.File(p) as f:
        off, n = f["j3c/0/0"].file_extent()
        assert off % 4096 == 0 and n == a.nbytes and f._eof == os.path.getsize(p)
        assert np.array_equal(f["j3c/0/0"][...], a)


@pytest.mark.parametrize("chunks,compression,shuffle", [((4, 3), None, False), ((5, 50), "gzip", True),
                                                        (None, "gzip", False), ((16, 16), None, True)])
def test_chunked_and_filtered_datasets(tmp_path, chunks, compression, shuffle):
    p = str(tmp_path / "c.h5")
    rng = np.random.default_rng(3)
    a = rng.standard_normal((13, 10)) + 1j * rng.standard_normal((13, 10))
    b = np.arange(1000, dtype=np.int64).reshape(10, 100)
    with h5lite.Writer(p) as w:
        w.create_dataset("a", a, chunks=chunks if chunks != (5, 50) else (5, 5), compression=compression,
                         shuffle=shuffle)
        w.create_dataset("b", b, chunks=chunks if chunks in (None, (5, 50)) else (4, 25), compression=compression,
                         shuffle=shuffle)
    with h5lite.File(p) as f:
        for k, v in (("a", a), ("b", b)):
            d = f[k]
            assert d._layout[0] == "chunked" and d.file_extent() is None
            assert np.array_equal(d[...], v) and np.array_equal(d[3:9], v[3:9]) and np.array_equal(d[:, 2], v[:, 2])
            out = np.empty((2,) + v.shape[1:], v.dtype)
            d.read_direct(out, slice(7, 9))
            assert np.array_equal(out, v[7:9])
            with pytest.raises(h5lite.H5FormatError):
                d.memmap()


def test_header_continuation_blocks(tmp_path):
    """object headers split over a continuation block (libhdf5 does this when attributes are added later)"""
    p = str(tmp_path / "k.h5")
    a = np.arange(12.0).reshape(3, 4)
    with h5lite.Writer(p) as w:
        w["x"] = a
    raw = bytearray(open(p, "rb").read())
    with h5lite.File(p) as f:
        addr = f["x"]._addr
        msgs = f._messages(addr)
    # rebuild the header: first block = dataspace + continuation, second block (appended to the file) = the rest
    enc = [struct.pack("<HHB3x", t, len(b), 0) + bytes(b) for t, b in msgs]
    tail = b"".join(enc[1:])
    cont_at = len(raw)
    first = enc[0] + struct.pack("<HHB3x", 0x10, 16, 0) + struct.pack("<QQ", cont_at, len(tail))
    assert len(first) <= sum(len(e) for e in enc)
    hsize = sum(len(e) for e in enc)
    raw[addr:addr + 16] = struct.pack("<BxHII4x", 1, len(msgs) + 1, 1, hsize)
    raw[addr + 16:addr + 16 + hsize] = first + struct.pack("<HHB3x", 0, hsize - len(first) - 8, 0) \
        + bytes(hsize - len(first) - 8)
    raw += tail
    raw[addr + 2:addr + 4] = struct.pack("<H", len(msgs) + 2)          # + continuation + padding NIL message
    q = str(tmp_path / "k2.h5")
    open(q, "wb").write(bytes(raw))
    with h5lite.File(q) as f:
        assert np.array_equal(f["x"][...], a)


def test_corrupt_files_fail_cleanly(tmp_path):
    """random byte flips and truncations: every outcome is a clean read, a KeyError (a damaged name no longer
    matches) or H5FormatError -- no hang, no giant allocation from a corrupt length field"""
    import random
    p = str(tmp_path / "a.h5")
    with h5lite.Writer(p) as w:
        for k in range(40):
            w["j3c/%d/0" % k] = np.arange(6.0).reshape(2, 3) + k
        w["j3c-kptij"] = np.zeros((40, 2, 3))
        w.create_dataset("c", np.arange(100.0).reshape(10, 10), chunks=(4, 4), compression="gzip")
    raw = open(p, "rb").read()
    rnd = random.Random(1)
    q = str(tmp_path / "f.h5")
    outcomes = set()
    for trial in range(300):
        b = bytearray(raw)
        if trial % 3 == 0:
            b = b[:rnd.randrange(8, len(b))]
        else:
            for _ in range(rnd.randrange(1, 6)):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        with open(q, "wb") as fh:
            fh.write(bytes(b))
        try:
            with h5lite.File(q) as f:
                for k in f["j3c"].keys():
                    f["j3c"][k]["0"][...]
                f["j3c-kptij"][...]
                f["c"][...]
            outcomes.add("ok")
        except (h5lite.H5FormatError, KeyError) as e:
            outcomes.add(type(e).__name__)
    assert outcomes <= {"ok", "H5FormatError", "KeyError"} and "H5FormatError" in outcomes
