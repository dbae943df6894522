import sys
import numpy as np, torch
sys.path.insert(0, ".")
from libdmet_preview_b200.device import get_device
dev = get_device()
torch.manual_seed(0)
for M, K in ((1000, 300), (2831, 500), (5000, 64), (11325, 1000)):
    X = torch.randn(M, K, dtype=torch.float64, device="cuda")
    for lower in (True, False):
        E0 = torch.randn(M, M, dtype=torch.float64, device="cuda")
        E = E0.clone()
        dev.dgemm_tn(X, X, E, alpha=2.0, accumulate=True, lower_only=lower)
        ref = E0 + 2.0 * (X @ X.T)
        d = (E - ref)
        if lower:
            tm = torch.arange(M, device="cuda") // 128
            d = d * (tm[:, None] >= tm[None, :])
        bad = (d.abs() > 1e-9).nonzero()
        print("dgemm M=%d K=%d lower=%d err %.3e nbad %d" % (M, K, lower, d.abs().max().item(), bad.shape[0]),
              bad[:4].tolist() if bad.shape[0] else "")
        E = E0.clone()
        dev.dgemm_tn(X, X, E, alpha=2.0, accumulate=False, lower_only=lower)
        d = (E - 2.0 * (X @ X.T))
        if lower:
            d = d * (tm[:, None] >= tm[None, :])
        print("   no-acc err %.3e" % d.abs().max().item())
