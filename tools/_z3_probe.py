import sys; sys.path.insert(0, ".")
import numpy as np, torch
from libdmet_preview_b200.device import get_device
dev = get_device()
naux, nao, neo, G = 1000, 200, 150, 4
A = torch.randn(G, naux * nao, nao, dtype=torch.complex128, device="cuda")
B = torch.randn(8, neo, nao, dtype=torch.complex128, device="cuda")
X = torch.empty(G, naux, neo, nao, dtype=torch.complex128, device="cuda")
segs = np.array([[g, g, 0, 0] for g in range(G)], dtype=np.int32)
for _ in range(2):
    dev.zgemm_tn(A, B, segs, X, c_off=np.arange(G) * naux * neo * nao, rdiv=nao, s_outer=neo * nao, s_inner=1, s_col=nao, nbatch=G, nseg=1)
torch.cuda.synchronize()
