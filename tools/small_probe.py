"""One get_emb_eri call at a small shape (development aid for ncu launch lists): python tools/small_probe.py sweep_min"""
import sys
import torch
sys.path.insert(0, ".")
from libdmet_preview_b200 import synthetic, eri_transform as et
shapes = {"c1_hchain": ([1, 1, 3], 4, 30, 6), "c2_graphene": ([3, 3, 1], 26, 150, 40),
          "sweep_min": ([2, 2, 2], 100, 500, 50)}
kmesh, nao, naux, neo = shapes[sys.argv[1] if len(sys.argv) > 1 else "sweep_min"]
gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=1)
C = synthetic.make_C_ao_lo(kmesh, nao, seed=2)
basis = synthetic.make_emb_basis(kmesh, nao, neo, seed=3)
for _ in range(3):
    et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
torch.cuda.synchronize()
