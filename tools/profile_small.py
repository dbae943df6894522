"""Where a small get_emb_eri call spends its host time (BASELINE configs[0], [1] and the sweep minimum): wall time per
call and the top Python frames.  Development aid; prints to stdout."""
import cProfile, pstats, sys, time, io
import numpy as np, torch
sys.path.insert(0, ".")
from libdmet_preview_b200 import synthetic, eri_transform as et
for name, (kmesh, nao, naux, neo) in {"c1_hchain": ([1, 1, 3], 4, 30, 6), "c2_graphene": ([3, 3, 1], 26, 150, 40),
                                       "sweep_min": ([2, 2, 2], 100, 500, 50)}.items():
    gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=1)
    C = synthetic.make_C_ao_lo(kmesh, nao, seed=2)
    basis = synthetic.make_emb_basis(kmesh, nao, neo, seed=3)
    for _ in range(5):
        et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    torch.cuda.synchronize()
    n = 200 if nao < 50 else 40
    t = time.perf_counter()
    for _ in range(n):
        et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    torch.cuda.synchronize()
    print("%s: %.3f ms per call" % (name, (time.perf_counter() - t) / n * 1e3), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        et.get_emb_eri(gdf.cell, gdf, C_ao_lo=C, basis=basis)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(16)
    print("\n".join(s.getvalue().splitlines()[:32]), flush=True)
