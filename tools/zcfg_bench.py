"""Stage-1 GEMM throughput at the target shape for the current / forced tile configuration (development aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from libdmet_preview_b200.device import get_device
dev = get_device()
naux, nao, neo, G = 1000, 200, int(os.environ.get("NEO", "150")), 4
A = torch.randn(G, naux * nao, nao, dtype=torch.complex128, device="cuda")
B = torch.randn(8, neo, nao, dtype=torch.complex128, device="cuda")
X = torch.empty(G, naux, neo, nao, dtype=torch.complex128, device="cuda")
S = torch.empty(naux, neo, neo, dtype=torch.complex128, device="cuda")
segs = np.array([[g, g, 0, 0] for g in range(G)], dtype=np.int32)
segs2 = np.array([[g, g, 0, 1] for g in range(G)], dtype=np.int32)
Xv = X.reshape(G, naux * neo, nao)
def run1():
    dev.zgemm_tn(A, B, segs, X, c_off=np.arange(G) * naux * neo * nao, rdiv=nao, s_outer=neo * nao, s_inner=1, s_col=nao, nbatch=G, nseg=1)
def run2():
    dev.zgemm_tn(Xv, B, segs2, S, rdiv=1, s_outer=neo, s_inner=0, s_col=1, nbatch=1, nseg=G, accumulate=True)
ref = (A[1] @ B[1].T).reshape(naux, nao, neo).transpose(1, 2)
run1(); torch.cuda.synchronize()
print("check 1a", (X[1] - ref).abs().max().item())
for name, fn, fl in (("1a", run1, 8.0 * G * naux * nao * nao * neo), ("1b", run2, 8.0 * G * naux * neo * nao * neo)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5 * 1e-3
    print("cfg=%s neo=%d stage %s: %.3f ms %.2f TFLOP/s algorithmic" % (os.environ.get("LDM_FORCE_ZCFG", "auto"), neo, name, t * 1e3, fl / t / 1e12))
