"""GPU: each kernel of libldm_b200.so, called through the C ABI, against numpy on the same seeded inputs.
Tolerances are absolute on O(1)..O(10) data; FP64 throughout."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _z(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.mark.parametrize("za,zb,M,N,K,nseg,nbatch,cA,cB,acc", [
    (1, 1, 64, 40, 8, 1, 1, 0, 0, False),
    (1, 1, 200, 150, 200, 1, 1, 0, 0, False),
    (3, 2, 333, 37, 26, 2, 3, 0, 1, False),
    (3, 2, 1000, 150, 52, 3, 2, 1, 0, True),
    (2, 4, 1031, 100, 201, 2, 2, 1, 1, False),
    (2, 2, 700, 6, 3, 1, 2, 0, 0, False),
    (1, 1, 1, 1, 1, 1, 1, 1, 1, True),
    (2, 3, 129, 161, 9, 2, 1, 0, 0, False),      # two N tiles
    (1, 1, 517, 200, 31, 1, 1, 0, 1, False),     # N = 200 -> 256x40 tile family
    (1, 1, 300, 48, 17, 1, 1, 0, 0, False),      # N = 48 -> b=3 family
])
@pytest.mark.parametrize("m3", [1, 0])
def test_zgemm_tn(dev, za, zb, M, N, K, nseg, nbatch, cA, cB, acc, m3):
    """both forms of the complex product: three real multiplications (default) and the classical four"""
    old = dev.set_option("zgemm_3m", m3)
    try:
        _zgemm_case(dev, za, zb, M, N, K, nseg, nbatch, cA, cB, acc)
    finally:
        dev.set_option("zgemm_3m", old)


def _zgemm_case(dev, za, zb, M, N, K, nseg, nbatch, cA, cB, acc):
    rng = np.random.default_rng(M + N + K)
    A, B = _z(rng, za, M, K), _z(rng, zb, N, K)
    segs = np.zeros((nbatch, nseg, 4), dtype=np.int32)
    segs[..., 0] = rng.integers(0, za, (nbatch, nseg))
    segs[..., 1] = rng.integers(0, zb, (nbatch, nseg))
    segs[..., 2], segs[..., 3] = cA, cB
    C0 = _z(rng, nbatch, M, N)
    ref = C0.copy() if acc else np.zeros_like(C0)
    for b in range(nbatch):
        for s in range(nseg):
            a = A[segs[b, s, 0]].conj() if cA else A[segs[b, s, 0]]
            bb = B[segs[b, s, 1]].conj() if cB else B[segs[b, s, 1]]
            ref[b] += 0.75 * (a @ bb.T)
    Cd = dev.to_device(C0, torch.complex128)
    dev.zgemm_tn(dev.to_device(A, torch.complex128), dev.to_device(B, torch.complex128), segs, Cd,
                 c_off=np.arange(nbatch) * M * N, s_outer=N, alpha=0.75, accumulate=acc, nbatch=nbatch, nseg=nseg)
    assert np.abs(Cd.cpu().numpy() - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("m3", [1, 0])
def test_zgemm_every_width(dev, m3):
    """the tile chooser over the whole range of N (every fragment count of both tile families, the short last n-tile,
    two and three n-tiles), with more tiles than CTAs so that the wave-rotated tile order is exercised -- a wrong
    pairing of tile shape and short-last shortcut would drop whole column fragments"""
    old = dev.set_option("zgemm_3m", m3)
    try:
        rng = np.random.default_rng(77)
        M, K = 9500, 11
        A = _z(rng, 1, M, K)
        Ad = dev.to_device(A, torch.complex128)
        for N in sorted(set(range(1, 210, 9)) | {56, 64, 72, 80, 104, 120, 128, 150, 152, 160, 200}):
            B = _z(rng, 1, N, K)
            Cd = dev.empty((1, M, N), torch.complex128)
            dev.zgemm_tn(Ad, dev.to_device(B, torch.complex128), [[0, 0, 0, 1]], Cd)
            ref = A[0] @ B[0].conj().T
            assert np.abs(Cd.cpu().numpy()[0] - ref).max() < 1e-11 * np.abs(ref).max(), N
    finally:
        dev.set_option("zgemm_3m", old)


def test_zgemm_3m_rounding_only(dev):
    """the 3-multiplication form differs from the 4-multiplication form by rounding: both within a few ulp of
    |A||B| of an exact (integer-valued) product, and the imaginary part has no systematic cancellation error even
    when |Re| >> |Im|"""
    rng = np.random.default_rng(5)
    M, N, K = 256, 150, 200
    A = (rng.integers(-8, 9, (1, M, K)) + 1j * rng.integers(-8, 9, (1, M, K))).astype(np.complex128)
    B = (rng.integers(-8, 9, (1, N, K)) + 1j * rng.integers(-8, 9, (1, N, K))).astype(np.complex128)
    exact = A[0] @ B[0].T                               # integers < 2^53: exact in either form
    outs = []
    for m3 in (1, 0):
        old = dev.set_option("zgemm_3m", m3)
        Cd = dev.empty((1, M, N), torch.complex128)
        dev.zgemm_tn(dev.to_device(A, torch.complex128), dev.to_device(B, torch.complex128), [[0, 0, 0, 0]], Cd)
        dev.set_option("zgemm_3m", old)
        outs.append(Cd.cpu().numpy()[0])
        assert np.array_equal(outs[-1], exact)
    # generic data, nearly real product: Im is ~1e-6 of Re
    A = _z(rng, 1, M, K)
    B = _z(rng, 1, N, K)
    B.imag *= 1e-6
    A.imag *= 1e-6
    ref = A[0] @ B[0].T
    scale = np.abs(A[0]) @ np.abs(B[0]).T
    for m3 in (1, 0):
        old = dev.set_option("zgemm_3m", m3)
        Cd = dev.empty((1, M, N), torch.complex128)
        dev.zgemm_tn(dev.to_device(A, torch.complex128), dev.to_device(B, torch.complex128), [[0, 0, 0, 0]], Cd)
        dev.set_option("zgemm_3m", old)
        err = np.abs(Cd.cpu().numpy()[0] - ref) / scale
        assert err.max() < 1e-14, (m3, err.max())
    with pytest.raises(RuntimeError):
        dev.set_option("no_such_option", 1)


def test_zgemm_strided_output(dev):
    """the stage-1a epilogue: rows r = (L, p) written transposed as out[L][c][p]"""
    rng = np.random.default_rng(7)
    naux, nao, neo = 5, 13, 11
    A, B = _z(rng, 1, naux * nao, nao), _z(rng, 1, neo, nao)
    out = dev.empty((naux, neo, nao), torch.complex128)
    dev.zgemm_tn(dev.to_device(A, torch.complex128), dev.to_device(B, torch.complex128), [[0, 0, 0, 0]], out,
                 rdiv=nao, s_outer=neo * nao, s_inner=1, s_col=nao)
    ref = (A[0] @ B[0].T).reshape(naux, nao, neo).transpose(0, 2, 1)
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-12


@pytest.mark.parametrize("M,N,K,lower,acc,pad", [
    (128, 128, 16, False, False, 0), (300, 200, 100, False, True, 0), (1000, 1000, 333, True, True, 1),
    (515, 515, 48, True, False, 0), (130, 77, 5, False, False, 3), (1, 1, 1, True, False, 1),
    # enough 128 x 128 tiles to fill the SMs: the large tile shape (smaller products run with 64 x 64 tiles)
    (2200, 2200, 40, True, True, 0), (1700, 1500, 24, False, True, 1),
])
def test_dgemm_tn_and_mirror(dev, M, N, K, lower, acc, pad):
    rng = np.random.default_rng(M + K)
    A = rng.standard_normal((M, K + pad))
    B = A if lower else rng.standard_normal((N, K + pad))
    C0 = rng.standard_normal((M, N))
    ref = (C0 if acc else 0.0) + 2.0 * (A[:, :K] @ B[:, :K].T)
    Ad = dev.to_device(np.ascontiguousarray(np.pad(A, ((0, 0), (0, (K + pad) % 2)))), torch.float64)
    Bd = Ad if lower else dev.to_device(np.ascontiguousarray(np.pad(B, ((0, 0), (0, (K + pad) % 2)))), torch.float64)
    Cd = dev.to_device(C0, torch.float64)
    dev.dgemm_tn(Ad, Bd, Cd, K=K, alpha=2.0, accumulate=acc, lower_only=lower)
    got = Cd.cpu().numpy()
    if lower:
        # the lower triangle is the contract; what lies above the diagonal tiles (64 or 128 wide) is untouched
        low = np.tril(np.ones((M, M), dtype=bool))
        assert np.abs((got - ref) * low).max() < 1e-11 * max(1.0, np.abs(ref).max())
        tm = np.arange(M) // 128
        mask = tm[:, None] >= tm[None, :]
        assert np.array_equal(got[~mask], C0[~mask])                   # tiles above the diagonal untouched
        dev.mirror_lower(Cd)
        full = np.tril(ref) + np.tril(ref, -1).T
        assert np.abs(Cd.cpu().numpy() - full).max() < 1e-11 * max(1.0, np.abs(ref).max())
    else:
        assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_transpose_d2z_ksum(dev):
    rng = np.random.default_rng(3)
    x = _z(rng, 3, 37, 70)
    xd = dev.to_device(x, torch.complex128)
    assert np.array_equal(dev.ztranspose(xd).cpu().numpy(), x.transpose(0, 2, 1))
    got = dev.ztranspose(xd, conj=True, scale=0.5).cpu().numpy()
    assert np.abs(got - 0.5 * x.conj().transpose(0, 2, 1)).max() < 1e-15
    r = rng.standard_normal((4, 9))
    assert np.array_equal(dev.d2z(dev.to_device(r, torch.float64)).cpu().numpy(), r.astype(np.complex128))
    out, imag = dev.ksum_real(xd, scale=0.25)
    assert np.abs(out.cpu().numpy() - 0.25 * x.sum(0).real).max() < 1e-14
    assert abs(imag - np.abs(x.sum(0).imag).max()) < 1e-13


@pytest.mark.parametrize("n", [1, 4, 11, 30])
def test_restore_and_jk(dev, n):
    from oracle import pyscf_lib as olib
    rng = np.random.default_rng(n)
    npair = n * (n + 1) // 2
    x = rng.standard_normal((npair, npair))
    e4 = x + x.T
    e4d = dev.to_device(e4, torch.float64)
    assert np.array_equal(dev.restore_s1(e4d, n).cpu().numpy(), olib.restore(1, e4, n))
    assert np.array_equal(dev.restore_s8(e4d, n).cpu().numpy(), olib.restore(8, e4, n))
    d = rng.standard_normal((n, n))
    d = d + d.T
    vj, vk = dev.jk_s4(e4d, dev.to_device(d, torch.float64))
    rj, rk = olib.dot_eri_dm(e4, d, hermi=1)
    assert np.abs(vj.cpu().numpy() - rj).max() < 1e-11 * max(1, np.abs(rj).max())
    assert np.abs(vk.cpu().numpy() - rk).max() < 1e-11 * max(1, np.abs(rk).max())
    # non-symmetric ERI block (the ab block): J only, PySCF convention J_ij = sum_kl (ij|kl) D_kl
    eab = rng.standard_normal((npair, npair))
    vj2, none = dev.jk_s4(dev.to_device(eab, torch.float64), dev.to_device(d, torch.float64), with_k=False)
    assert none is None
    assert np.abs(vj2.cpu().numpy() - olib.dot_eri_dm(eab, d, with_k=False)[0]).max() < 1e-11 * max(1, np.abs(rj).max())


@pytest.mark.parametrize("n", [32, 33, 64, 65, 97, 129, 161])
def test_jk_streaming_variants(dev, n):
    """every register-tile width of the streaming J/K kernel (n <= 64, 128, 160, 256), non-symmetric density, against
    the defining sums evaluated row by row on the host (scf.py:269-271: J_ij = sum_kl (ij|kl) D_kl,
    K_jk = sum_il (ij|kl) D_il)"""
    rng = np.random.default_rng(n)
    npair = n * (n + 1) // 2
    x = rng.standard_normal((npair, npair))
    e4 = x + x.T
    d = rng.standard_normal((n, n))
    vj, vk = dev.jk_s4(dev.to_device(e4, torch.float64), dev.to_device(d, torch.float64))
    r, c = np.tril_indices(n)
    dd = np.where(r == c, d[r, c], d[r, c] + d[c, r])
    rj = np.zeros((n, n))
    rj[r, c] = e4 @ dd
    rj[c, r] = rj[r, c]
    rk = np.zeros((n, n))
    M = np.zeros((n, n))
    for P in range(npair):
        i, j = r[P], c[P]
        M[r, c] = e4[P]
        M[c, r] = e4[P]
        rk[j] += M @ d[i]
        if i != j:
            rk[i] += M @ d[j]
    assert np.abs(vj.cpu().numpy() - rj).max() < 1e-11 * np.abs(rj).max()
    assert np.abs(vk.cpu().numpy() - rk).max() < 1e-11 * np.abs(rk).max()
    again = dev.jk_s4(dev.to_device(e4, torch.float64), dev.to_device(d, torch.float64))
    assert torch.equal(again[0], vj) and torch.equal(again[1], vk)          # fixed summation order


@pytest.mark.parametrize("n", [17, 31, 64, 65, 100, 128, 129, 150, 160])
def test_jk_lower_triangle_kernel(dev, n):
    """`ldm_jk_s4_symm`: symmetric ERI block and symmetric density, only the lower triangle of the block is read
    (the upper triangle is poisoned here to prove it) -- against PySCF's dot_eri_dm convention (scf.py:269-271) and
    bit-for-bit reproducible"""
    from oracle import pyscf_lib as olib
    rng = np.random.default_rng(1000 + n)
    npair = n * (n + 1) // 2
    x = rng.standard_normal((npair, npair))
    e4 = x + x.T
    d = rng.standard_normal((n, n))
    d = d + d.T
    rj, rk = olib.dot_eri_dm(e4, d, hermi=1)
    poisoned = np.tril(e4) + np.triu(np.full_like(e4, 1e30), 1)
    e4d, dd = dev.to_device(poisoned, torch.float64), dev.to_device(d, torch.float64)
    vj, vk = dev.jk_s4(e4d, dd, symmetric=True)
    assert np.abs(vj.cpu().numpy() - rj).max() < 1e-11 * np.abs(rj).max()
    assert np.abs(vk.cpu().numpy() - rk).max() < 1e-11 * np.abs(rk).max()
    vj2, none = dev.jk_s4(e4d, dd, with_k=False, symmetric=True)
    assert none is None and torch.equal(vj2, vj)
    again = dev.jk_s4(e4d, dd, symmetric=True)
    assert torch.equal(again[0], vj) and torch.equal(again[1], vk)          # fixed summation order
    # agrees with the general kernel on the full symmetric block
    gj, gk = dev.jk_s4(dev.to_device(e4, torch.float64), dd)
    assert np.abs((gj - vj).cpu().numpy()).max() < 1e-11 * np.abs(rj).max()
    assert np.abs((gk - vk).cpu().numpy()).max() < 1e-11 * np.abs(rk).max()


def test_synth_block_bit_exact(dev):
    from libdmet_preview_b200 import synthetic
    g = synthetic.SyntheticGDF([2, 1, 3], 9, 14, seed=77)
    for (i, j) in [(0, 0), (1, 4), (5, 2)]:
        out = dev.empty((14, 9, 9), torch.complex128)
        dev.synth_block(out, 14, 9, g.keys(i, j), g.scale)
        assert np.array_equal(out.cpu().numpy(), g.load(i, j))           # bit for bit


def test_error_reporting(dev):
    from libdmet_preview_b200._lib import LdmError
    A = dev.empty((1, 4, 4), torch.complex128)
    with pytest.raises(LdmError):
        dev.zgemm_tn(A, A, [[5, 0, 0, 0]], dev.empty((4, 4), torch.complex128))     # slice out of range
    with pytest.raises(LdmError):
        dev.dgemm_tn(dev.empty((4, 3)), dev.empty((4, 3)), dev.empty((4, 4)))       # odd leading dimension


def test_gemm_pipeline_many_tiles_per_cta(dev):
    """Regression for the shared-memory stage hand-over: with many tiles per persistent CTA the consumers used to
    release a stage while its last fragment loads were still in flight (random rows of a tile corrupted, run to run).
    Sizes give ~12 (dgemm) and ~20 (zgemm) tiles per CTA; results must be exact and repeatable."""
    rng = np.random.default_rng(11)
    M, K = 5500, 1000
    X = rng.standard_normal((M, K))
    ref = 2.0 * (X @ X.T)
    Xd = dev.to_device(X, torch.float64)
    tm = np.arange(M) // 128
    mask = tm[:, None] >= tm[None, :]
    first = None
    for rep in range(3):
        E = dev.zeros((M, M))
        dev.dgemm_tn(Xd, Xd, E, alpha=2.0, accumulate=True, lower_only=True)
        got = E.cpu().numpy()
        assert np.abs((got - ref) * mask).max() < 1e-9
        first = got if first is None else first
        assert np.array_equal(got, first)
    A = _z(rng, 1, 200000, 64)
    B = _z(rng, 1, 150, 64)
    Ad, Bd = dev.to_device(A, torch.complex128), dev.to_device(B, torch.complex128)
    refz = A[0] @ B[0].T
    first = None
    for rep in range(3):
        C = dev.zeros((200000, 150), torch.complex128)
        dev.zgemm_tn(Ad, Bd, [[0, 0, 0, 0]], C, accumulate=True)
        got = C.cpu().numpy()
        assert np.abs(got - refz).max() < 1e-10
        first = got if first is None else first
        assert np.array_equal(got, first)


def test_eri_pipeline_misuse_is_reported(dev):
    """error convention of the C ABI: negative status + message, surfaced as LdmError (no crash, no silent result)"""
    import ctypes as C
    from libdmet_preview_b200._lib import LdmError, check
    from libdmet_preview_b200 import eri_transform as et, synthetic
    gdf = synthetic.SyntheticGDF([1, 1, 2], 4, 6)
    CT = et.build_CT(gdf, synthetic.make_C_ao_lo([1, 1, 2], 4))
    eri = dev.zeros((1, 10, 10))
    with et.EriBuild(CT, gdf.naux, eri, 2, 2) as b:
        with pytest.raises(LdmError):
            et.EriBuild(CT, gdf.naux, eri, 2, 2)                 # a second build on the same handle
        with pytest.raises(LdmError):
            b.end_kl(1)                                          # transfer momentum without blocks
        with pytest.raises(LdmError):
            b.block_store(0, 0, 0, 0)                            # no resident store registered
        with pytest.raises(LdmError):
            b.block_synth(0, 7, 0, gdf.keys(0, 0), gdf.scale)    # k index out of range
        with pytest.raises(LdmError):
            check(dev.lib.ldm_eri_set_mode(dev.h, 1))            # GSO mode needs two spin flavours
        b.block_synth(0, 0, 0, gdf.keys(0, 0), gdf.scale)
        with pytest.raises(LdmError):
            b.end_kl(3)                                          # weight must be 0, 1 or 2
        b.end_kl(1)
        b.finish()
    with pytest.raises(LdmError):
        check(dev.lib.ldm_eri_finish(dev.h))                     # no build open
    msg = dev.lib.ldm_last_error()
    assert isinstance(msg, bytes) and len(msg) > 0


def test_large_neo_fallbacks(dev):
    """neo beyond the shared-memory row budget (npair * 8 B > 200 KB, neo > 225): s4 -> s1 and J/K fall back to
    global-memory gathers; the GEMM splits N over several tiles"""
    from oracle import pyscf_lib as olib
    n = 232
    npair = n * (n + 1) // 2
    rng = np.random.default_rng(4)
    e4 = rng.standard_normal((npair, npair))
    e4d = dev.to_device(e4, torch.float64)
    d = rng.standard_normal((n, n))
    d = d + d.T
    vj, vk = dev.jk_s4(e4d, dev.to_device(d, torch.float64))
    e1 = dev.restore_s1(e4d, n)                       # 23 GB: stays on the device, spot-checked
    idx = np.tril_indices(n)
    tri = np.zeros((n, n), dtype=np.int64)
    tri[idx] = np.arange(npair)
    tri[(idx[1], idx[0])] = np.arange(npair)
    for (i, j, k, l) in [(0, 0, 0, 0), (5, 200, 17, 3), (231, 7, 100, 231), (40, 41, 42, 43)]:
        assert e1[i, j, k, l].item() == e4[tri[i, j], tri[k, l]]
    assert torch.equal(e1[3, 9], e1[9, 3]) and torch.equal(e1[3, 9], e1[3, 9].T)
    del e1
    dp = d + d.T
    dp[np.diag_indices(n)] *= 0.5
    rj = np.zeros((n, n))
    rj[idx] = e4 @ dp[idx]
    rj[(idx[1], idx[0])] = rj[idx]
    assert np.abs(vj.cpu().numpy() - rj).max() < 1e-9
    vk = vk.cpu().numpy()
    for (j, k) in [(0, 0), (7, 201), (231, 230), (100, 3)]:            # K_jk = sum_il (ij|kl) D_il
        ref = sum(np.dot(e4[tri[i, j], tri[k, :]], d[i, :]) for i in range(n))
        assert abs(vk[j, k] - ref) < 1e-9


def test_zgemm_many_n_tiles(dev):
    rng = np.random.default_rng(6)
    A, B = _z(rng, 1, 300, 33), _z(rng, 1, 450, 33)
    out = dev.empty((300, 450), torch.complex128)
    dev.zgemm_tn(dev.to_device(A, torch.complex128), dev.to_device(B, torch.complex128), [[0, 0, 0, 1]], out)
    assert np.abs(out.cpu().numpy() - A[0] @ B[0].conj().T).max() < 1e-11
