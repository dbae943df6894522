"""Where one DMET iteration (get_emb_basis + embHam) at the target shape spends its wall time beyond the ERI build:
cProfile of the third iteration over a resident pooled GDF tensor.  Development aid; prints to stdout."""
import cProfile, io, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from libdmet_preview_b200 import synthetic, eri_transform as et, lattice as lat, slater
import bench

kmesh, nao, naux, neo, nspin = bench.workload("target")
gdf = synthetic.PooledGDF(synthetic.SyntheticGDF(kmesh, nao, naux, seed=bench.GDF_SEED), bench.pool_blocks(kmesh, nao, naux))
C_ao_lo_h, basis_h = bench.make_inputs(kmesh, nao, neo, nspin)
base = et.ResidentGDF(gdf)
nval = neo - nao // 2
nimp = neo - nval
Lat = lat.Lattice(base.cell, kmesh)
Lat.set_val_virt_core(nval, nimp - nval, nao - nimp)
hcore = synthetic.make_hermitian_k(kmesh, nao, seed=41)
vhf = synthetic.make_hermitian_k(kmesh, nao, seed=42, scale=0.3)
rdm1 = synthetic.make_rdm1_k(hcore + vhf, max(1, nao // 3)) * 2.0
ovlp = np.asarray([np.eye(nao, dtype=np.complex128)] * len(base.kpts_scaled))
Lat.set_Ham(None, base, C_ao_lo_h, eri_symmetry=4, ovlp=ovlp, hcore=hcore, rdm1=rdm1, vhf=vhf)
for it in range(3):
    pr = cProfile.Profile() if it == 2 else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if pr:
        pr.enable()
    bas = slater.get_emb_basis(Lat, Lat.rdm1_lo_R * 0.5)
    st = {}
    Ham, _ = slater.embHam(Lat, bas, None, stats=st)
    torch.cuda.synchronize()
    if pr:
        pr.disable()
    print("iteration %d: %.3f s (zgemm %.1f ms, dgemm %.1f ms)" % (it, time.perf_counter() - t0, st.get("zgemm_ms", -1),
                                                                   st.get("dgemm_ms", -1)), flush=True)
    del Ham
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print("\n".join(s.getvalue().splitlines()[:45]))
