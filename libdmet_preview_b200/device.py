"""Device context: one libldm_b200 handle per process (one process per GPU), thin typed wrappers over the C ABI.

torch is used only as plumbing here: it owns device/pinned memory (tensors), the current CUDA stream and, in
`dist.py`, the NCCL process group.  Every arithmetic operation on the path goes through libldm_b200.so.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import check

_ctx = None


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class Device(object):
    def __init__(self, index=None):
        if not torch.cuda.is_available():
            raise RuntimeError("libdmet_preview_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if index is None:
            index = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device()))
        self.index = index
        torch.cuda.set_device(index)
        self.torch_device = torch.device("cuda", index)
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.ldm_create(index, C.byref(h)))
        self.h = h

    # ---- plumbing -------------------------------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.torch_device).cuda_stream)

    def synchronize(self):
        torch.cuda.current_stream(self.torch_device).synchronize()

    def to_device(self, a, dtype=None):
        """numpy / torch -> contiguous device tensor (H2D copy on the current stream)."""
        if isinstance(a, torch.Tensor):
            t = a
            if dtype is not None and t.dtype != dtype:
                t = t.to(dtype)
            return t.to(self.torch_device).contiguous()
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.torch_device, non_blocking=False).contiguous()

    def to_host(self, t):
        """device tensor -> numpy array through pinned memory (a pageable destination runs at a fraction of the
        PCIe rate; torch caches the pinned block, so the page-locking cost is paid once per size)"""
        if t.numel() * t.element_size() < (1 << 20):
            return t.cpu().numpy()
        out = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.torch_device).synchronize()
        return out.numpy()

    def empty(self, shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.torch_device)

    def zeros(self, shape, dtype=torch.float64):
        return torch.zeros(shape, dtype=dtype, device=self.torch_device)

    def set_option(self, name, value):
        """handle option (include/ldm_b200.h: "zgemm_3m" = 1 three-multiplication complex products, 0 classical);
        returns the previous value"""
        old = C.c_int(0)
        check(self.lib.ldm_set_option(self.h, name.encode(), int(value), C.byref(old)))
        return old.value

    def release_workspaces(self):
        """free the ERI pipeline's cached device workspaces (they are kept between builds otherwise)"""
        check(self.lib.ldm_release_workspaces(self.h))

    def launch_count(self):
        return int(self.lib.ldm_launch_count(self.h))

    # ---- kernels --------------------------------------------------------------------------------------
    def zgemm_tn(self, A, B, segs, C_out, M=None, N=None, K=None, c_off=None, rdiv=1, s_outer=None, s_inner=0,
                 s_col=1, alpha=1.0, accumulate=False, nbatch=None, nseg=None):
        """C[b](r,c) (+)= alpha * sum_s sum_k opA(A[az](r,k)) opB(B[bz](c,k)); A (za,M,K), B (zb,N,K) complex128.
        segs: int array (nbatch, nseg, 4) of (az, bz, conjA, conjB)."""
        assert A.dtype == torch.complex128 and B.dtype == torch.complex128 and C_out.dtype == torch.complex128
        assert A.is_contiguous() and B.is_contiguous() and C_out.is_contiguous()
        A3 = A.reshape(-1, A.shape[-2], A.shape[-1])
        B3 = B.reshape(-1, B.shape[-2], B.shape[-1])
        M = A3.shape[1] if M is None else M
        N = B3.shape[1] if N is None else N
        K = A3.shape[2] if K is None else K
        assert B3.shape[2] == K
        segs = np.ascontiguousarray(segs, dtype=np.int32).reshape(-1, 4)
        if nbatch is None:
            nbatch = 1 if nseg is None else segs.shape[0] // nseg
        if nseg is None:
            nseg = segs.shape[0] // nbatch
        assert nbatch * nseg == segs.shape[0]
        if s_outer is None:
            s_outer = N
        off_p = None
        if c_off is not None:
            c_off = np.ascontiguousarray(c_off, dtype=np.int64)
            assert c_off.size == nbatch
            off_p = c_off.ctypes.data_as(_lib.c_i64p)
        check(self.lib.ldm_zgemm_tn(self.h, self.stream, _ptr(A3), A3.shape[0], _ptr(B3), B3.shape[0], M, N, K,
                                    nseg, nbatch, segs.ctypes.data_as(_lib.c_i32p), _ptr(C_out), off_p, rdiv,
                                    s_outer, s_inner, s_col, float(alpha), int(bool(accumulate))))
        return C_out

    def dgemm_tn(self, A, B, C_out, K=None, alpha=1.0, accumulate=False, lower_only=False):
        """C(r,c) (+)= alpha sum_k A(r,k) B(c,k); A (M, lda), B (N, ldb) float64 row-major."""
        assert A.dtype == torch.float64 and B.dtype == torch.float64 and C_out.dtype == torch.float64
        assert A.stride(-1) == 1 and B.stride(-1) == 1 and C_out.stride(-1) == 1
        M, N = A.shape[0], B.shape[0]
        K = A.shape[1] if K is None else K
        check(self.lib.ldm_dgemm_tn(self.h, self.stream, _ptr(A), A.stride(0), _ptr(B), B.stride(0), M, N, K,
                                    _ptr(C_out), C_out.stride(0), float(alpha), int(bool(accumulate)),
                                    int(bool(lower_only))))
        return C_out

    def mirror_lower(self, E):
        assert E.dtype == torch.float64 and E.dim() == 2 and E.stride(1) == 1
        check(self.lib.ldm_mirror_lower(self.h, self.stream, _ptr(E), E.shape[0], E.stride(0)))
        return E

    def phase_transform(self, x, W, out_real=False, scale=1.0, want_imag=True):
        """out[b][k][...] = scale * sum_R W[k][R] x[b][R][...].  x: (batch, nin, ...) real or complex;
        W: (nout, nin) complex128 on the device.  Returns (out, max|imag| or None)."""
        in_real = x.dtype == torch.float64
        assert in_real or x.dtype == torch.complex128
        assert x.is_contiguous() and W.is_contiguous() and W.dtype == torch.complex128
        batch, nin = x.shape[0], x.shape[1]
        nout = W.shape[0]
        assert W.shape[1] == nin
        X = int(np.prod(x.shape[2:]))
        out = self.empty((batch, nout) + tuple(x.shape[2:]), torch.float64 if out_real else torch.complex128)
        imag = C.c_double(0.0)
        check(self.lib.ldm_phase_transform(self.h, self.stream, _ptr(x), _ptr(out), _ptr(W), nin, nout, X, batch,
                                           float(scale), int(in_real), int(out_real),
                                           C.byref(imag) if (out_real and want_imag) else None))
        return out, (imag.value if (out_real and want_imag) else None)

    def lattice_dft(self, x, kmesh, forward, out_real=False, scale=1.0, want_imag=True):
        """factorised lattice DFT on the mesh's own k-points: x (batch, ncells, ...) real or complex ->
        (batch, nkpts, ...) complex (or real part only).  Returns (out, max|imag| or None)."""
        in_real = x.dtype == torch.float64
        assert in_real or x.dtype == torch.complex128
        assert x.is_contiguous() and x.shape[1] == int(np.prod(kmesh))
        km = (list(kmesh) + [1, 1, 1])[:3]
        X = int(np.prod(x.shape[2:]))
        out = self.empty(tuple(x.shape), torch.float64 if out_real else torch.complex128)
        imag = C.c_double(0.0)
        km_arr = (C.c_int32 * 3)(*[int(v) for v in km])
        check(self.lib.ldm_lattice_dft(self.h, self.stream, _ptr(x), _ptr(out), km_arr, X, x.shape[0],
                                       int(bool(forward)), float(scale), int(in_real), int(out_real),
                                       C.byref(imag) if (out_real and want_imag) else None))
        return out, (imag.value if (out_real and want_imag) else None)

    def ztranspose(self, x, conj=False, scale=1.0):
        """(batch, rows, cols) complex128 -> (batch, cols, rows)."""
        assert x.dtype == torch.complex128 and x.is_contiguous() and x.dim() == 3
        out = self.empty((x.shape[0], x.shape[2], x.shape[1]), torch.complex128)
        check(self.lib.ldm_ztranspose(self.h, self.stream, _ptr(x), _ptr(out), x.shape[0], x.shape[1], x.shape[2],
                                      int(bool(conj)), float(scale)))
        return out

    def pack_tril(self, x, out_real=False):
        """(rows, n, n) complex128 -> (rows, n(n+1)/2) lower triangles, complex or (out_real) real part + max|imag|"""
        assert x.dtype == torch.complex128 and x.is_contiguous() and x.dim() == 3 and x.shape[1] == x.shape[2]
        rows, n = x.shape[0], x.shape[1]
        out = self.empty((rows, n * (n + 1) // 2), torch.float64 if out_real else torch.complex128)
        imag = C.c_double(0.0)
        check(self.lib.ldm_pack_tril(self.h, self.stream, _ptr(x), _ptr(out), rows, n, int(bool(out_real)),
                                     C.byref(imag) if out_real else None))
        return out, (imag.value if out_real else None)

    def d2z(self, x):
        assert x.dtype == torch.float64 and x.is_contiguous()
        out = self.empty(tuple(x.shape), torch.complex128)
        check(self.lib.ldm_d2z(self.h, self.stream, _ptr(x), _ptr(out), x.numel()))
        return out

    def ksum_real(self, x, scale=1.0):
        """(nk, ...) complex128 -> (...) float64 = scale * Re sum_k ; returns (out, max|Im sum|)."""
        assert x.dtype == torch.complex128 and x.is_contiguous()
        out = self.empty(tuple(x.shape[1:]), torch.float64)
        imag = C.c_double(0.0)
        check(self.lib.ldm_ksum_real(self.h, self.stream, _ptr(x), _ptr(out), x.shape[0],
                                     int(np.prod(x.shape[1:])), float(scale), C.byref(imag)))
        return out, imag.value

    def restore_s1(self, eri4, n):
        out = self.empty((n, n, n, n))
        check(self.lib.ldm_restore_s1(self.h, self.stream, _ptr(eri4), _ptr(out), n))
        return out

    def restore_s8(self, eri4, n):
        npair = n * (n + 1) // 2
        out = self.empty((npair * (npair + 1) // 2,))
        check(self.lib.ldm_restore_s8(self.h, self.stream, _ptr(eri4), _ptr(out), n))
        return out

    def jk_s4(self, eri4, dm, with_k=True, symmetric=False):
        """J, K of one s4 ERI block and one density matrix.  symmetric=True: `eri4` is a symmetric matrix AND `dm` is
        symmetric (the reference's hermi=1 calls on the restricted / aa / bb blocks) -- only the lower triangle of
        `eri4` is read."""
        n = dm.shape[-1]
        assert eri4.dtype == torch.float64 and eri4.is_contiguous() and dm.is_contiguous()
        vj = self.empty((n, n))
        vk = self.empty((n, n)) if with_k else None
        fn = self.lib.ldm_jk_s4_symm if symmetric else self.lib.ldm_jk_s4
        check(fn(self.h, self.stream, _ptr(eri4), _ptr(dm), _ptr(vj), _ptr(vk) if with_k else None, n))
        return vj, vk

    def scale_eri(self, eri, n, symmetry, weights):
        """in place: s4 (npair, npair) with per-pair impurity counts, or s1 (n,n,n,n) with 0/1 flags (int32)"""
        assert eri.dtype == torch.float64 and eri.is_contiguous() and weights.dtype == torch.int32
        check(self.lib.ldm_scale_eri(self.h, self.stream, _ptr(eri), n, int(symmetry), _ptr(weights)))
        return eri

    def max_abs(self, x):
        assert x.dtype == torch.float64 and x.is_contiguous()
        out = C.c_double(0.0)
        check(self.lib.ldm_max_abs(self.h, self.stream, _ptr(x), x.numel(), C.byref(out)))
        return out.value

    def unpack_stored(self, src, naux, nao, flags=0, out=None):
        """stored cderi entry (rows, ncols) on the device -> (naux, nao, nao) complex128 block
        (ldm_unpack_stored: widen real data, unpack the Hermitian lower triangle, conjugate-transpose swapped
        pairs, zero the auxiliary rows the entry lacks)"""
        rows, ncols = int(src.shape[0]), int(src.shape[1])
        if src.dtype == torch.float64:
            flags |= 2
        elif src.dtype != torch.complex128:
            raise TypeError("stored entries are float64 or complex128")
        src = src.contiguous()
        if out is None:
            out = self.empty((naux, nao, nao), torch.complex128)
        check(self.lib.ldm_unpack_stored(self.h, self.stream, _ptr(src) if rows else None, _ptr(out), int(naux), rows,
                                         int(nao), ncols, int(flags)))
        return out

    def synth_block(self, out, naux, nao, keys, scale, aux_offset=0):
        check(self.lib.ldm_synth_block(self.h, self.stream, _ptr(out), naux, nao, int(aux_offset), int(keys[0]),
                                       int(keys[1]), int(keys[2]), int(keys[3]), float(scale)))
        return out


def get_device(index=None):
    """Process-wide device context (created on first use)."""
    global _ctx
    if _ctx is None:
        _ctx = Device(index)
    return _ctx
