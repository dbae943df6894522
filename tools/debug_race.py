import sys, os
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import problem
from libdmet_preview_b200 import eri_transform as et, synthetic
from libdmet_preview_b200.device import get_device
dev = get_device()
gdf, C, basis = problem([1, 1, 2], 200, 1000, 150)
a = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
b = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
print("repeat diff", np.abs(a - b).max(), "max", np.abs(a).max())
for g, kg in ((1, 1), (2, 1), (4, 4)):
    c = et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, group=g, kl_group=kg)
    print("group", g, kg, "diff vs first", np.abs(a - c).max())
g2 = synthetic.SyntheticGDF([1, 1, 2], 200, 1000, seed=gdf.seed, scale=2.0 * gdf.scale)
e2 = et.get_emb_eri(g2.cell, g2, C_ao_lo=C, basis=basis)
print("scale test", np.abs(e2 - 4 * a).max())
# direct dgemm accumulate check at the syrk shape
npair = 11325
X = torch.randn(npair, 3000, dtype=torch.float64, device="cuda")
E1 = torch.zeros(npair, npair, dtype=torch.float64, device="cuda")
dev.dgemm_tn(X, X, E1, K=1000, alpha=1.0, accumulate=True, lower_only=True)
dev.dgemm_tn(X[:, 1000:], X[:, 1000:], E1, K=2000, alpha=2.0, accumulate=True, lower_only=True)
ref = X[:, :1000] @ X[:, :1000].T + 2.0 * (X[:, 1000:] @ X[:, 1000:].T)
print("dgemm acc err", ((E1 - ref).tril()).abs().max().item())
