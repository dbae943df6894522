import sys
import numpy as np, torch
sys.path.insert(0, ".")
from libdmet_preview_b200.device import get_device
dev = get_device()
torch.manual_seed(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 11325
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
X = torch.randn(M, K, dtype=torch.float64, device="cuda")
ref = 2.0 * (X @ X.T)
tm = torch.arange(M, device="cuda") // 128
mask = tm[:, None] >= tm[None, :]
for rep in range(4):
    E = torch.zeros(M, M, dtype=torch.float64, device="cuda")
    dev.dgemm_tn(X, X, E, alpha=2.0, accumulate=False, lower_only=True)
    torch.cuda.synchronize()
    d = (E - ref) * mask
    bad = (d.abs() > 1e-9).nonzero()
    tiles = sorted(set((int(r) // 128, int(c) // 128) for r, c in bad.tolist()))
    print("rep", rep, "nbad", bad.shape[0], "tiles", tiles[:10])
    if bad.shape[0]:
        r0, c0 = tiles[0]
        sub = d[r0 * 128:(r0 + 1) * 128, c0 * 128:(c0 + 1) * 128].abs() > 1e-9
        rows = sub.any(1).nonzero().flatten().tolist()
        cols = sub.any(0).nonzero().flatten().tolist()
        print("   rows in tile", rows[:40], "\n   cols in tile", cols[:70])
        lin = r0 * (r0 + 1) // 2 + c0
        print("   lin", lin, "cta", lin % 148, "iteration", lin // 148)
