"""A few launches of the HBM-bound kernels at the target shape, for `ncu -k regex:...` captures (development aid)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from libdmet_preview_b200.device import get_device
from libdmet_preview_b200 import synthetic, eri_transform as et
dev = get_device()
kmesh, nk, n = [4, 4, 4], 64, 200
xr = torch.randn(14, nk, n, n, dtype=torch.float64, device="cuda")
xc = torch.randn(14, nk, n, n, dtype=torch.complex128, device="cuda")
for _ in range(2):
    dev.lattice_dft(xr, kmesh, True)
    dev.lattice_dft(xc, kmesh, True)
    dev.lattice_dft(xc, kmesh, False, out_real=True, scale=1.0 / nk, want_imag=False)
del xr, xc
neo = 150
npair = neo * (neo + 1) // 2
E = torch.randn(npair, npair, dtype=torch.float64, device="cuda")
E = (E + E.T).contiguous()
D = torch.randn(neo, neo, dtype=torch.float64, device="cuda"); D = (D + D.T).contiguous()
for _ in range(2):
    dev.jk_s4(E, D)
    dev.jk_s4(E, D, symmetric=True)
# pack_sym: one transfer momentum of a 1x1x2 mesh at the target block shape
gdf = synthetic.SyntheticGDF([1, 1, 2], 200, 1000, seed=3)
C = synthetic.make_C_ao_lo([1, 1, 2], 200, seed=1)
basis = synthetic.make_emb_basis([1, 1, 2], 200, 150, seed=2)
et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis, return_device=True)
torch.cuda.synchronize()
