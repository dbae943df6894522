"""Host-side time breakdown of one end-to-end get_emb_eri call (development aid)."""
import sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from libdmet_preview_b200 import synthetic, eri_transform as et
from libdmet_preview_b200.device import get_device
dev = get_device()
kmesh, nao, naux, neo, nspin = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3_nio_uhf")
gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=2026)
C = synthetic.make_C_ao_lo(kmesh, nao, seed=1, spin=(nspin if nspin > 1 else None))
basis = synthetic.make_emb_basis(kmesh, nao, neo, seed=2, spin=nspin)
host = bench.HostPoolProvider(gdf, 16)
for _ in range(2):
    et.get_emb_eri(gdf.cell, host, C_ao_lo=C, basis=basis, source="host")
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
r = et.get_emb_eri(gdf.cell, host, C_ao_lo=C, basis=basis, source="host")
pr.disable()
print("total", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
