"""Achieved HBM bandwidth of the streaming kernels at the target shape (Nk=64, nao=nlo=200, neo=150), CUDA events,
inputs larger than or comparable to L2 rotated between iterations.  Writes gpurun_out/hbm_kernels.json."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from libdmet_preview_b200.device import get_device
dev = get_device()
res = {}
peak = 6531.9
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

def timeit(fn, nrot, reps=20):
    for i in range(3): fn(i % nrot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i % nrot)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

def report(name, bytes_alg, t):
    gbs = bytes_alg / t / 1e9
    res[name] = {"algorithmic_bytes": bytes_alg, "seconds": t, "achieved_gbs": gbs, "frac_of_measured_hbm_copy": gbs / peak}
    print("%-42s %8.1f MB %9.1f us %8.1f GB/s  %.2f of %.0f" % (name, bytes_alg / 1e6, t * 1e6, gbs, gbs / peak, peak))

kmesh, nk, n = [4, 4, 4], 64, 200
NR = 2      # rotate over independent stacks; each call moves several hundred MB (>> L2, >> launch overhead)
NB = 14     # the seven (hcore, ovlp, fock, ...) transforms of Lattice.set_Ham x 2 spins, batched
xr = [torch.randn(NB, nk, n, n, dtype=torch.float64, device="cuda") for _ in range(NR)]
xc = [torch.randn(NB, nk, n, n, dtype=torch.complex128, device="cuda") for _ in range(NR)]
t = timeit(lambda i: dev.lattice_dft(xr[i], kmesh, True), NR)
report("R2k real (14,64,200,200) -> complex", NB * (8 + 16) * nk * n * n, t)
t = timeit(lambda i: dev.lattice_dft(xc[i], kmesh, True), NR)
report("R2k complex (14,64,200,200)", NB * 32 * nk * n * n, t)
t = timeit(lambda i: dev.lattice_dft(xc[i], kmesh, False, out_real=True, scale=1.0 / nk, want_imag=False), NR)
report("k2R complex (14,64,200,200) -> real", NB * (16 + 8) * nk * n * n, t)
from libdmet_preview_b200 import fourier
W = fourier._phase_dev(kmesh, True)
t = timeit(lambda i: dev.phase_transform(xc[i], W), NR)
report("dense phase-matrix R2k (fallback kernel)", NB * 32 * nk * n * n, t)
t = timeit(lambda i: dev.ztranspose(xc[i].reshape(-1, n, n)), NR)
report("ztranspose (896,200,200)", NB * 32 * nk * n * n, t)
del xr, xc
neo = 150
npair = neo * (neo + 1) // 2
E = [torch.randn(npair, npair, dtype=torch.float64, device="cuda") for _ in range(2)]
t = timeit(lambda i: dev.mirror_lower(E[i]), 2)
report("mirror_lower (11325^2)", 8 * npair * npair, t)           # reads the lower half, writes the upper half
t = timeit(lambda i: dev.restore_s1(E[i], neo), 2, reps=5)
report("restore s4->s1 (150^4)", 8 * npair * npair + 8 * neo ** 4, t)
t = timeit(lambda i: dev.restore_s8(E[i], neo), 2, reps=5)
report("restore s4->s8", 8 * npair * npair // 2 * 2, t)
D = torch.randn(neo, neo, dtype=torch.float64, device="cuda"); D = (D + D.T).contiguous()
t = timeit(lambda i: dev.jk_s4(E[i], D), 2, reps=5)
report("J/K from s4 ERI (neo=150), general kernel", 8 * npair * npair, t)
t = timeit(lambda i: dev.jk_s4(E[i], D, symmetric=True), 2, reps=5)
report("J/K from s4 ERI (neo=150), lower triangle", 8 * npair * npair, t)
t = timeit(lambda i: dev.jk_s4(E[i], D, with_k=False, symmetric=True), 2, reps=5)
report("J only, lower triangle", 8 * npair * npair, t)
blk = torch.empty(1000, 200, 200, dtype=torch.complex128, device="cuda")
t = timeit(lambda i: dev.synth_block(blk, 1000, 200, (1, 2, 3, 4), 0.25), 1, reps=5)
report("synth_block (1000,200,200) generator", 16 * 1000 * 200 * 200, t)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"hbm_copy_peak_gbs": peak, "kernels": res}, open("gpurun_out/hbm_kernels.json", "w"), indent=1)
