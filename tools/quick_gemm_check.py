"""Quick on-GPU sanity check of the two tensor-core GEMM kernels against torch (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from libdmet_preview_b200.device import get_device

dev = get_device()
torch.manual_seed(0)

def check_z(za, zb, M, N, K, nseg, nbatch, conjA, conjB, accumulate=False, alpha=1.0):
    A = torch.randn(za, M, K, dtype=torch.complex128, device="cuda")
    B = torch.randn(zb, N, K, dtype=torch.complex128, device="cuda")
    rng = np.random.default_rng(1)
    segs = np.zeros((nbatch, nseg, 4), dtype=np.int32)
    segs[..., 0] = rng.integers(0, za, (nbatch, nseg))
    segs[..., 1] = rng.integers(0, zb, (nbatch, nseg))
    segs[..., 2] = conjA
    segs[..., 3] = conjB
    Cout = torch.randn(nbatch, M, N, dtype=torch.complex128, device="cuda")
    ref = Cout.clone() if accumulate else torch.zeros_like(Cout)
    for b in range(nbatch):
        for s in range(nseg):
            a = A[segs[b, s, 0]]; bb = B[segs[b, s, 1]]
            if conjA: a = a.conj()
            if conjB: bb = bb.conj()
            ref[b] += alpha * (a @ bb.T)
    dev.zgemm_tn(A, B, segs, Cout, c_off=np.arange(nbatch) * M * N, rdiv=1, s_outer=N, s_inner=0, s_col=1,
                 alpha=alpha, accumulate=accumulate, nbatch=nbatch, nseg=nseg)
    torch.cuda.synchronize()
    err = (Cout - ref).abs().max().item()
    print("zgemm za=%d zb=%d M=%d N=%d K=%d nseg=%d nbatch=%d cA=%d cB=%d acc=%d: max err %.3e (ref max %.2e)" % (
        za, zb, M, N, K, nseg, nbatch, conjA, conjB, accumulate, err, ref.abs().max().item()))
    return err

def check_d(M, N, K, lower, accumulate=False, alpha=1.0, pad=0):
    A = torch.randn(M, K + pad, dtype=torch.float64, device="cuda")
    B = A if lower else torch.randn(N, K + pad, dtype=torch.float64, device="cuda")
    Cout = torch.randn(M, N, dtype=torch.float64, device="cuda")
    ref = (Cout.clone() if accumulate else torch.zeros_like(Cout)) + alpha * (A[:, :K] @ B[:, :K].T)
    keep = Cout.clone()
    dev.dgemm_tn(A, B, Cout, K=K, alpha=alpha, accumulate=accumulate, lower_only=lower)
    torch.cuda.synchronize()
    if lower:
        # tiles strictly above the diagonal are untouched
        tm = torch.arange(M, device="cuda") // 128
        mask = tm[:, None] >= tm[None, :]
        err = ((Cout - ref) * mask).abs().max().item()
        err2 = ((Cout - keep) * (~mask)).abs().max().item()
        dev.mirror_lower(Cout); torch.cuda.synchronize()
        err3 = (Cout - torch.tril(ref) - torch.tril(ref, -1).T).abs().max().item()
        print("dsyrk M=%d K=%d: err %.3e untouched %.3e mirror %.3e" % (M, K, err, err2, err3))
        return max(err, err2, err3)
    err = (Cout - ref).abs().max().item()
    print("dgemm M=%d N=%d K=%d acc=%d: err %.3e" % (M, N, K, accumulate, err))
    return err

worst = 0.0
worst = max(worst, check_z(1, 1, 64, 40, 8, 1, 1, 0, 0))
worst = max(worst, check_z(1, 1, 200, 150, 200, 1, 1, 0, 0))
worst = max(worst, check_z(3, 2, 333, 37, 26, 2, 3, 0, 1))
worst = max(worst, check_z(3, 2, 1000, 150, 52, 3, 2, 1, 0, accumulate=True, alpha=0.5))
worst = max(worst, check_z(2, 4, 4096, 100, 200, 2, 2, 1, 1))
worst = max(worst, check_z(2, 2, 700, 6, 3, 1, 2, 0, 0))
worst = max(worst, check_d(128, 128, 16, False))
worst = max(worst, check_d(300, 200, 100, False, accumulate=True, alpha=2.0))
worst = max(worst, check_d(1000, 1000, 333, True, accumulate=True, alpha=2.0, pad=3 if False else 1))
worst = max(worst, check_d(11325 // 4, 11325 // 4, 500, True))
print("WORST", worst)

# throughput at the target shapes
def bench_z():
    naux, nao, neo, G = 1000, 200, 150, 4
    A = torch.randn(G, naux * nao, nao, dtype=torch.complex128, device="cuda")
    B = torch.randn(8, neo, nao, dtype=torch.complex128, device="cuda")
    X = torch.empty(G, naux, neo, nao, dtype=torch.complex128, device="cuda")
    segs = np.array([[g, g, 0, 0] for g in range(G)], dtype=np.int32)
    def run():
        dev.zgemm_tn(A, B, segs, X, c_off=np.arange(G) * naux * neo * nao, rdiv=nao, s_outer=neo * nao, s_inner=1,
                     s_col=nao, nbatch=G, nseg=1)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3 * 1e-3
    print("stage 1a: %.3f ms  %.2f TFLOP/s (8 flop/cmac, algorithmic N=150)" % (t * 1e3, 8.0 * G * naux * nao * nao * neo / t / 1e12))
    # stage 1b: chained over G
    S = torch.empty(naux, neo, neo, dtype=torch.complex128, device="cuda")
    segs2 = np.array([[g, g, 0, 1] for g in range(G)], dtype=np.int32)
    Xv = X.reshape(G, naux * neo, nao)
    def run2():
        dev.zgemm_tn(Xv, B, segs2, S, rdiv=1, s_outer=neo, s_inner=0, s_col=1, nbatch=1, nseg=G)
    for _ in range(2): run2()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3): run2()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3 * 1e-3
    print("stage 1b: %.3f ms  %.2f TFLOP/s" % (t * 1e3, 8.0 * G * naux * neo * nao * neo / t / 1e12))

def bench_d():
    npair, K = 11325, 2000
    XT = torch.randn(npair, K, dtype=torch.float64, device="cuda")
    E = torch.zeros(npair, npair, dtype=torch.float64, device="cuda")
    for _ in range(2): dev.dgemm_tn(XT, XT, E, alpha=2.0, accumulate=True, lower_only=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): dev.dgemm_tn(XT, XT, E, alpha=2.0, accumulate=True, lower_only=True)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3 * 1e-3
    print("stage 3 syrk K=2000: %.3f ms  %.2f TFLOP/s (syrk flops K*n*(n+1))" % (t * 1e3, K * npair * (npair + 1.0) / t / 1e12))

bench_z()
bench_d()
