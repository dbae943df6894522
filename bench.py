#!/usr/bin/env python
"""Benchmark of the embedding-ERI hot path (`get_emb_eri`, GDF, restricted, s4, time-reversal symmetry).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload target|small|...]

One "step" = one complete `get_emb_eri` build of the named workload: C_ao_emb construction, stage 1 (two half
transformations per (k_i,k_j) block), symmetrise + pack, stage 3 (Gram products), mirror.  Prints ONE JSON line.

value      algorithmic FP64 TFLOP/s of the whole job, (F1 + F3) / time, inputs resident in HBM (BASELINE.md section 3:
           F1 = 8 naux nao neo (nao+neo) B,  F3 = naux npair (npair+1) G)
e2e        same metric through the public API with HOST inputs: GDF blocks are fetched from a (pinned) host provider
           and copied host->device inside the timed region, the ERI is copied back to the host
roofline   stage-1 complex GEMM kernel (DMMA), timed live with CUDA events on its launch stream
cpu_baseline / --impl reference
           the numpy oracle (the reference cannot be imported here: PySCF / h5py are absent) on the host cores,
           on a bounded sample of the same workload (a few (k_i,k_j) blocks + one Gram product), scaled by block
           and Gram counts to the whole job
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time


def _host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


if "reference" in sys.argv[1:]:
    # the CPU arm uses every host core also under torch.distributed.run (which exports OMP_NUM_THREADS=1, and
    # OpenBLAS falls back to that variable): pinned here, BEFORE numpy loads its BLAS
    for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(_host_cores())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kmesh, nao, naux, neo)
    "target": ([4, 4, 4], 200, 1000, 150),       # BASELINE.json north_star / BASELINE.md section 3
    "mid": ([2, 2, 4], 200, 1000, 150),
    "small": ([2, 2, 2], 100, 500, 50),          # sweep minimum
    "tiny": ([1, 2, 3], 24, 64, 20),
    # shapes of BASELINE.json configs[0..2] (sizes estimated in SURVEY.md section 8; 5th entry = number of spins)
    "c1_hchain": ([1, 1, 3], 4, 30, 6),
    "c2_graphene": ([3, 3, 1], 26, 150, 40),
    "c3_nio_uhf": ([2, 2, 2], 78, 400, 106, 2),
    # BASELINE.json configs[4]: synthetic sweep nkpts 8-64, nao 100-300, naux 500-1500, neo 50-200
    "sweep_222_300_1500_200": ([2, 2, 2], 300, 1500, 200),
    "sweep_224_200_1000_100": ([2, 2, 4], 200, 1000, 100),
    "sweep_333_100_500_150": ([3, 3, 3], 100, 500, 150),
    "sweep_442_200_1000_150": ([4, 4, 2], 200, 1000, 150),
    "sweep_444_100_500_100": ([4, 4, 4], 100, 500, 100),
    "sweep_444_300_1500_200": ([4, 4, 4], 300, 1500, 200),
}


def workload(name):
    w = WORKLOADS[name]
    return (w[0], w[1], w[2], w[3], (w[4] if len(w) > 4 else 1))


# ---- the synthetic inputs of a workload: identical for every N and for the oracle digest -------------------------
GDF_SEED = 2026
POOL_BYTES = 140e9        # the GDF tensor is defined by a pool of distinct blocks that fits ONE GPU (synthetic.PooledGDF)
E2E_POOL = 4              # distinct blocks of the pinned host pool of the end-to-end leg (same for every N)
DIGEST = os.path.join(ROOT, "tests", "golden", "bench_digest.json")


def pool_blocks(kmesh, nao, naux):
    nk = int(np.prod(kmesh))
    return int(max(1, min(nk * nk, POOL_BYTES // (16 * naux * nao * nao))))


def make_inputs(kmesh, nao, neo, nspin):
    from libdmet_preview_b200 import synthetic
    C_ao_lo = synthetic.make_C_ao_lo(kmesh, nao, seed=1, spin=(nspin if nspin > 1 else None))
    basis = synthetic.make_emb_basis(kmesh, nao, neo, seed=2, spin=nspin)
    return C_ao_lo, basis


def sample_orbitals(neo):
    """embedding orbitals whose ERI entries are checked at full size: first, second, middle, last"""
    return sorted({0, min(1, neo - 1), max(0, neo // 2 - 1), neo - 1})


def blas_threads():
    try:
        import threadpoolctl
        return [{"api": i.get("internal_api"), "threads": i.get("num_threads")} for i in threadpoolctl.threadpool_info()
                if i.get("user_api") == "blas"]
    except Exception:
        return None


def flops(kmesh, nao, naux, neo, nspin=1):
    from libdmet_preview_b200.synthetic import trs_block_count
    B, G = trs_block_count(kmesh, True)
    npair = neo * (neo + 1) // 2
    F1 = 8.0 * naux * nao * neo * (nao + neo) * nspin * B
    # unrestricted: aa and bb are syrk products, ab a full product (= two syrk halves)
    F3 = float(naux) * npair * (npair + 1) * G * (1 if nspin == 1 else 4)
    return F1, F3, B, G


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = False
        self.th = threading.Thread(target=self.run, daemon=True)
        # NVML is initialised here, before the timed region: nvmlInit takes ~15 ms and serialises with CUDA calls
        self.nv = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = (nv, nv.nvmlDeviceGetHandleByIndex(index))
        except Exception:
            self.nv = None

    def _nvml(self):
        """in-process NVML sampling (no fork per sample, nothing that stalls the driver for milliseconds)"""
        nv, h = self.nv
        bits = [("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown",
                                        getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8))),
                ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40))),
                ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20))),
                ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap",
                                         getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons",
                              getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mask = int(get_reasons(h)) if get_reasons else 0
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            self.samples.append([str(sm), str(mx), str(pw)] + ["Active" if mask & b else "Not Active" for _, b in bits])
            time.sleep(0.05)

    def run(self):
        if self.nv is not None:
            try:
                self._nvml()
                return
            except Exception:
                pass
        while not self.stop:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], text=True, timeout=5)
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 7 and s[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# CPU baseline (oracle) on a bounded sample
# ---------------------------------------------------------------------------------------------------------
def cpu_sample(kmesh, nao, naux, neo, nblocks=12, budget_s=25.0, ngram=2):
    """Time the oracle's stage 1 (transform_ao_to_emb + hermi_sum + pack_tril + accumulate, in the reference's
    240-row chunks) on `nblocks` blocks (fewer if 60 % of `budget_s` is used up) and `ngram` stage-3 Gram products -- 10 to 20 s of CPU
    work at the target shape -- and scale the best block / Gram time to the whole job."""
    from oracle import eri_transform as o
    from oracle import pyscf_lib as olib
    from libdmet_preview_b200 import synthetic
    F1, F3, B, G = flops(kmesh, nao, naux, neo)
    nk = int(np.prod(kmesh))
    npair = neo * (neo + 1) // 2
    rng = np.random.default_rng(0)
    C_ao_emb = (rng.standard_normal((1, nk, nao, neo)) + 1j * rng.standard_normal((1, nk, nao, neo))) / nk ** 0.75
    Lblk = (rng.standard_normal((naux, nao * nao)) + 1j * rng.standard_normal((naux, nao * nao)))
    Lij_s4 = np.zeros((1, naux, npair), dtype=np.complex128)
    blksize = 240
    t_blocks = []
    t0 = time.perf_counter()
    for b in range(nblocks):
        tb = time.perf_counter()
        for l0 in range(0, naux, blksize):
            Lpq = Lblk[l0:l0 + blksize]
            Lij = o.transform_ao_to_emb(Lpq, C_ao_emb, b % nk, (b + 1) % nk).reshape(-1, neo, neo)
            olib.hermi_sum(Lij, axes=(0, 2, 1), hermi=olib.SYMMETRIC, inplace=True)
            Lij_s4[:, l0:l0 + Lpq.shape[0]] += olib.pack_tril(Lij).reshape(1, -1, npair)
        t_blocks.append(time.perf_counter() - tb)
        if time.perf_counter() - t0 > budget_s * 0.6:
            break
    t_block = min(t_blocks)
    X = np.ascontiguousarray(Lij_s4[0].real)
    eri = np.zeros((npair, npair))
    t_grams = []
    for _ in range(max(1, ngram)):
        tg = time.perf_counter()
        olib.dot(X.T, X, 1.0, eri, 1)
        t_grams.append(time.perf_counter() - tg)
    t_gram = min(t_grams)
    t_job = t_block * B + t_gram * G
    return {"t_block_s": t_block, "t_gram_s": t_gram, "t_job_s": t_job, "tflops": (F1 + F3) / t_job / 1e12,
            "nblocks_timed": len(t_blocks),
            "sample_seconds": float(sum(t_blocks) + sum(t_grams)),
            "sample": "%d of %d (ki,kj) blocks of stage 1 (240-row chunks) + %d of %d Gram products (best of each), "
                      "extrapolated linearly by block/Gram count" % (len(t_blocks), B, len(t_grams), G)}


def run_reference(args):
    kmesh, nao, naux, neo, nspin = workload(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    F1, F3, B, G = flops(kmesh, nao, naux, neo)      # the CPU sample times the restricted path
    times = []
    res = None
    for it in range(args.warmup + args.steps):
        res = cpu_sample(kmesh, nao, naux, neo, nblocks=2 if it < args.warmup else 10, budget_s=25.0)
        if it >= args.warmup:
            times.append(res["t_job_s"])
    t = float(np.mean(times))
    val = (F1 + F3) / t / 1e12
    line = {"impl": "reference", "metric": "get_emb_eri_fp64_tflops", "value": val, "unit": "TFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, kmesh, nao, naux, neo, B, G),
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": _host_cores(), "kind": "port",
                             "blas_threads": blas_threads(),
                             "sample": res["sample"] + "; numpy/OpenBLAS zgemm+dgemm, all host threads (pinned "
                             "before numpy loads, also under torch.distributed.run); the reference itself cannot be "
                             "imported (PySCF, h5py absent)"},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extrapolated": True, "measured_sample_seconds": res["sample_seconds"],
            "get_emb_eri_seconds": t}
    print(json.dumps(line), flush=True)


def config_dict(args, kmesh, nao, naux, neo, B, G):
    """identical in both arms (the driver compares it)"""
    nspin = workload(args.workload)[4]
    kind = "restricted" if nspin == 1 else "unrestricted"
    blk_mb = 16.0 * naux * nao * nao / 1e6
    npair = neo * (neo + 1) // 2
    return {"workload": "%s: get_emb_eri GDF %s s4 time-reversal, kmesh %s nkpts %d nao %d naux %d neo %d "
                        "(%d (ki,kj) blocks, %d Gram products)" % (args.workload, kind, "x".join(map(str, kmesh)),
                                                                  int(np.prod(kmesh)), nao, naux, neo, B, G),
            "kmesh": kmesh, "nao": nao, "naux": naux, "neo": neo, "nspin": nspin, "symmetry": 4,
            "t_reversal_symm": True,
            "l2_policy": "inputs larger than L2 (each GDF block %.0f MB, ERI %.0f MB)" % (blk_mb, npair * npair * 8 / 1e6)}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class HostPoolProvider(object):
    """GDF provider over pinned host memory: the E2E_POOL distinct blocks of `synthetic.PooledGDF(gdf, E2E_POOL)`,
    block (ki,kj) -> pool slot by its fixed map.  What a host-RAM resident cderi looks like to `get_emb_eri` (the
    full 758 GB tensor of the target shape is not materialised on the host; every block still crosses PCIe inside
    the timed region).  The pool is first touched by this process after it has been bound to the GPU's own NUMA
    node (`bind_to_gpu_numa_node`)."""

    def __init__(self, pooled):
        import torch
        from libdmet_preview_b200.device import get_device
        dev = get_device()
        g = pooled
        self.pooled = g
        self.kpts_scaled, self.kmesh, self.nao, self.naux = g.kpts_scaled, g.kmesh, g.nao, g.naux
        self.kpts, self.cell = g.kpts, g.cell
        self.pool = torch.empty((g.npool, g.naux, g.nao, g.nao), dtype=torch.complex128, pin_memory=True)
        tmp = dev.empty((g.naux, g.nao, g.nao), torch.complex128)
        for s in range(g.npool):
            dev.synth_block(tmp, g.naux, g.nao, g.inner.keys(*g.pool_pair(s)), g.scale)
            self.pool[s].copy_(tmp)
        dev.synchronize()

    def load(self, ki, kj):
        return self.pool[self.pooled.pool_index(ki, kj)]


def bind_to_gpu_numa_node(index):
    """CPU affinity of this process := the cores NVML reports as local to GPU `index`, so that pinned staging memory
    is first touched on the GPU's own NUMA node (ranks of a multi-GPU run otherwise all pull through one socket).
    Returns a short description for the JSON line; failures are reported, not raised."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return {"bound": False, "why": "GPU-local cores %d..%d not in this process's cpuset (%d cores allowed)" % (
                min(cpus) if cpus else -1, max(cpus) if cpus else -1, len(allowed))}
        os.sched_setaffinity(0, use)
        return {"bound": True, "cores": len(use), "first": min(use), "last": max(use)}
    except Exception as e:          # noqa: BLE001
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


def fp64_peak_live(seconds=3.0):
    """cuBLAS DGEMM 8192^3 through torch.matmul on this box, now: best of 10 (burst) and back to back for `seconds`
    (sustained); SM clock sampled meanwhile.  Outside every timed region."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    fl = 2.0 * n ** 3
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    with ClockSampler(torch.cuda.current_device()) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, int(seconds / best))
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        sus = e0.elapsed_time(e1) * 1e-3 / reps
    del a, b, c
    torch.cuda.empty_cache()
    return {"dgemm_8192_burst_tflops": fl / best / 1e12, "dgemm_8192_sustained_tflops": fl / sus / 1e12,
            "sustained_seconds": sus * reps, "clocks": clk.summary()}


def independent_sample(torch, dist, world, schedule, items, get_block, C_ao_lo, basis, kmesh, kpts_scaled, orbs,
                       nspin, device="cuda"):
    """ERI entries eri[blk][tri(a,b), tri(c,d)], a,b,c,d in `orbs`, evaluated WITHOUT any kernel of this repository:
    torch complex matmul / einsum (cuBLAS) on the device, straight from the definition
        Lambda^kL[L, (m,n)] = sum_{(i,j) in kL} conj(C_i[:,m]) . L(k_i,k_j)[L] . C_j[:,n]   (+ transpose if symmetrised)
        eri += w (Re Lambda^T Re Lambda [+ Im Lambda^T Im Lambda if w = 2])
    over this rank's work items, summed over ranks.  Returns (n_spin_pair, npr, npr) on every rank."""
    from libdmet_preview_b200.schedule import cell_vectors
    nk = len(kpts_scaled)
    R = torch.as_tensor(cell_vectors(kmesh), dtype=torch.float64, device=device)
    ks = torch.as_tensor(np.asarray(kpts_scaled), dtype=torch.float64, device=device)
    phase = torch.exp(-2j * np.pi * (R @ ks.T)).to(torch.complex128)              # (R, k)
    Cl = torch.as_tensor(np.asarray(C_ao_lo), device=device).to(torch.complex128)
    Cl = Cl[None] if Cl.dim() == 3 else Cl
    bs = torch.as_tensor(np.asarray(basis), device=device).to(torch.complex128)
    bs = bs[None] if bs.dim() == 3 else bs
    o = torch.as_tensor(orbs, device=device)
    Ce = []
    for s in range(nspin):
        bk = torch.einsum("Rlm,Rk->klm", bs[min(s, bs.shape[0] - 1)][..., o], phase)     # (k, nlo, no)
        Ce.append(torch.matmul(Cl[min(s, Cl.shape[0] - 1)], bk) / nk ** 0.75)            # (k, nao, no)
    no = len(orbs)
    ia, ib = np.tril_indices(no)
    ia_t, ib_t = torch.as_tensor(ia, device=device), torch.as_tensor(ib, device=device)
    npr = len(ia)
    pairs = [(0, 0)] if nspin == 1 else [(0, 0), (0, 1), (1, 1)]
    E = torch.zeros((len(pairs), npr, npr), dtype=torch.float64, device=device)
    for (u, l0, l1) in items:
        kL, w, blocks = schedule.units[u]
        Lam = torch.zeros((nspin, l1 - l0, npr), dtype=torch.complex128, device=device)
        for (ki, kj, sym) in blocks:
            Lb = get_block(ki, kj, l0, l1)
            for s in range(nspin):
                X = torch.matmul(Lb, Ce[s][kj])                                           # (rows, nao, no)
                T = torch.einsum("pa,Lpb->Lab", Ce[s][ki].conj(), X)
                if sym:
                    T = T + T.transpose(1, 2)
                Lam[s] += T[:, ia_t, ib_t]
        parts = [Lam.real] if w == 1 else [Lam.real, Lam.imag]
        wt = 2.0 if w == 2 else 1.0
        for part in parts:
            for n, (x, y) in enumerate(pairs):
                E[n] += wt * (part[x].T @ part[y])
    if world > 1:
        dist.all_reduce(E, op=dist.ReduceOp.SUM)
    return E


def sample_index(neo, orbs):
    no = len(orbs)
    ia, ib = np.tril_indices(no)
    return np.asarray([orbs[a] * (orbs[a] + 1) // 2 + orbs[b] for a, b in zip(ia, ib)])


def parity_report(eri_sample, indep, digest):
    """eri_sample: (spin_pair, npr, npr) entries of the kernels' ERI; indep: the torch evaluation; digest: oracle"""
    out = {"entries": int(eri_sample.size), "max_abs_value": float(np.abs(eri_sample).max()),
           "max_abs_vs_independent_torch": float(np.abs(eri_sample - indep).max()) if indep is not None else None,
           "max_abs_vs_oracle_digest": None}
    if digest is not None:
        ref = np.asarray(digest["eri_s4_lower"])
        if ref.shape == eri_sample.shape:
            out["max_abs_vs_oracle_digest"] = float(np.abs(eri_sample - ref).max())
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from libdmet_preview_b200 import synthetic, eri_transform as et
    from libdmet_preview_b200 import dist as ldist
    from libdmet_preview_b200.device import get_device
    from libdmet_preview_b200.schedule import build_schedule

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa else {"bound": False, "why": "--no-numa"}
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up; stdout carries exactly one JSON
        # line, so the banner is sent to stderr (fd level: the print comes from the C library)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = get_device(local_rank)

    kmesh, nao, naux, neo, nspin = workload(args.workload)
    F1, F3, B, G = flops(kmesh, nao, naux, neo, nspin)
    args.group, args.kl_group = et.auto_groups(nao, naux, neo, nspin, args.group or None, args.kl_group or None)
    npool = pool_blocks(kmesh, nao, naux)
    gdf = synthetic.PooledGDF(synthetic.SyntheticGDF(kmesh, nao, naux, seed=GDF_SEED), npool)
    C_ao_lo_h, basis_h = make_inputs(kmesh, nao, neo, nspin)
    C_ao_lo = dev.to_device(C_ao_lo_h, torch.complex128)
    basis = dev.to_device(basis_h, torch.float64)
    schedule = build_schedule(gdf.kpts_scaled, True)
    my_items = ldist.rank_items(schedule, nao, naux, neo, nspin, world)[rank]
    my_blocks = [(l0, l1, blk) for (u, l0, l1) in my_items for blk in schedule.units[u][2]]
    my_rows = sum(l1 - l0 for (l0, l1, _) in my_blocks)           # GDF rows this rank transforms
    orbs = sample_orbitals(neo)
    sidx = torch.as_tensor(sample_index(neo, orbs), device="cuda")
    digest = None
    if os.path.exists(DIGEST):
        with open(DIGEST) as f:
            digest = json.load(f).get(args.workload)

    peak_live = None
    if not args.no_peak:
        peak_live = fp64_peak_live()

    # ---- resident store of L blocks (inputs in HBM before the timed region) ----
    # slot of a block = position of its POOL index among the pool blocks this rank's aux range needs: the tensor
    # L(ki,kj) = pool[(ki nk + kj) mod npool] is the same for every N, only the subset held per rank changes
    npair = neo * (neo + 1) // 2
    ranges = sorted({(l0, l1) for (l0, l1, _) in my_blocks})
    stores, store_map, nslots = {}, {}, 0
    st_t = None
    for (l0, l1) in ranges:
        blks = [b for (a0, a1, b) in my_blocks if (a0, a1) == (l0, l1)]
        ids = sorted({gdf.pool_index(ki, kj) for (ki, kj, _) in blks})
        slot_of = {s: n for n, s in enumerate(ids)}
        st_t = dev.empty((len(ids), l1 - l0, nao, nao), torch.complex128)
        for s, n in slot_of.items():
            dev.synth_block(st_t[n], l1 - l0, nao, gdf.inner.keys(*gdf.pool_pair(s)), gdf.scale, aux_offset=l0)
        for (ki, kj, sym) in blks:
            store_map[(ki, kj, l0)] = slot_of[gdf.pool_index(ki, kj)]
        stores[(l0, l1)] = st_t
        nslots += len(ids)
    blk_bytes = naux * nao * nao * 16
    dev.synchronize()

    debug = bool(os.environ.get("BENCH_DEBUG"))

    def step(collect=None):
        t0 = time.perf_counter()
        CT = et.build_CT(gdf, C_ao_lo, basis)
        eri = et.emb_eri_device(gdf, CT, schedule=schedule, items=my_items, stores=stores, store_map=store_map,
                                group=args.group, kl_group=args.kl_group, stats=collect)
        if debug:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        if world > 1:
            dist.reduce(eri, dst=0, op=dist.ReduceOp.SUM)
        if debug:
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        if rank == 0:
            eri = et.finalize_eri(eri, neo, 4, nspin)
        if debug:
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            print("[rank %d] build %.1f ms (zgemm %.1f dgemm %.1f) reduce %.1f ms finalize %.1f ms" % (
                rank, (t1 - t0) * 1e3, (collect or {}).get("zgemm_ms", -1), (collect or {}).get("dgemm_ms", -1),
                (t2 - t1) * 1e3, (t3 - t2) * 1e3), file=sys.stderr, flush=True)
        return eri

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step({})                  # same code path as the timed steps (kernel-time events included)
    barrier()
    l0 = dev.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    zg_ms = dg_ms = 0.0
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(args.steps):
            st = {}
            out = step(st)
            zg_ms += st["zgemm_ms"]
            dg_ms += st["dgemm_ms"]
        e1.record()
        barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    launches = dev.launch_count() - l0
    checksum = float(out.sum().item()) if rank == 0 else 0.0
    tt = torch.tensor([t_dev, zg_ms, dg_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_step = tt[0].item() / args.steps
    value = (F1 + F3) / t_step / 1e12

    # ---- parity of the timed result at full size (outside the timed region) ----
    parity = None
    if not args.no_parity:
        indep = independent_sample(torch, dist, world, schedule, my_items,
                                   lambda ki, kj, a, b: stores[(a, b)][store_map[(ki, kj, a)]],
                                   C_ao_lo_h, basis_h, kmesh, gdf.kpts_scaled, orbs, nspin)
        if rank == 0:
            got = out[:, sidx][:, :, sidx].cpu().numpy()
            parity = parity_report(got, indep.cpu().numpy(), (digest or {}).get("resident"))
            parity["pool_blocks"] = npool
            parity["orbitals"] = orbs
            parity["eri_sum"] = checksum
    del out

    # ---- roofline of the dominant kernel (stage-1 zgemm), events recorded on its launch stream ----
    F1_mine = F1 * my_rows / float(B * naux)
    with open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")) as f:
        pk = json.load(f)
    peak = pk["dgemm_8192_sustained_tflops"]
    ach = F1_mine * args.steps / (zg_ms * 1e-3) / 1e12 if zg_ms > 0 else None
    traffic = None
    traffic_src = None
    for name in ("zgemm_traffic_r02.json", "zgemm_traffic_r01.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
            traffic_src = "profiles/" + name + " (ncu --set full, one stage-1a launch of %d blocks)" % args.group
            break
    m3 = os.environ.get("LDM_ZGEMM_3M", "1") != "0"
    live = peak_live["dgemm_8192_sustained_tflops"] if peak_live else None
    roofline = {"bound": "tensor", "kernel": "zgemm_tn_kernel (stage 1: both half transformations)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": args.group * (16.0 * naux * nao * nao + 16.0 * naux * nao * neo),
                "peak_live": live, "frac_live": (ach / live) if (ach and live) else None,
                "peak_live_detail": peak_live,
                "complex_product": "3M" if m3 else "4M",
                "executed_tflops": (ach * (0.75 if m3 else 1.0)) if ach else None,
                "note": ("achieved counts ALGORITHMIC flops (8 per complex multiply-add); the kernel evaluates complex "
                         "products with three real multiplications, so the tensor pipe executes 0.75 of them and "
                         "frac can exceed 1") if m3 else None,
                "peak_source": "cuBLAS DGEMM 8192^3 sustained 4 s on this pool's B200 (tools/probe_peaks.py -> "
                               "profiles/fp64_peaks_r01.json; MEASURED_PEAKS.json carries no FP64 figure); peak_live = "
                               "the same probe run by this process before the timed region",
                "share_of_step": zg_ms * 1e-3 / (t_step * args.steps),
                "stage3_share_of_step": dg_ms * 1e-3 / (t_step * args.steps),
                "stage3_dgemm": {"achieved": (F3 * sum((1 if schedule.units[u][1] == 1 else 2) * (l1 - l0)
                                                           for (u, l0, l1) in my_items) / float(G * naux)) *
                                 args.steps / (dg_ms * 1e-3) / 1e12 if dg_ms > 0 else None, "unit": "TFLOP/s"}}

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        stores.clear()
        store_map.clear()
        del st_t
        torch.cuda.empty_cache()
        pooled_e2e = synthetic.PooledGDF(gdf.inner, E2E_POOL)
        host = HostPoolProvider(pooled_e2e)
        n_e2e = max(1, args.e2e_steps)
        st = {}

        def e2e_call(provider, stats=None):
            kw = dict(C_ao_lo=C_ao_lo_h, basis=basis_h, symmetry=4, kl_group=args.kl_group, group=args.group,
                      stats=stats)
            if world > 1:       # the reference's own keyword for its multi-process path (eri_transform.py:71)
                kw["use_mpi"] = True
            if provider is host:
                kw["source"] = "host"
            return et.get_emb_eri(gdf.cell, provider, **kw)

        def timed_e2e(provider, n):
            barrier()
            t0 = time.perf_counter()
            res = None
            stt = {}
            for _ in range(n):
                del res                  # a DMET loop drops the previous eri too: the pinned result buffer is reused
                stt = {}
                res = e2e_call(provider, stt)
            barrier()
            t = (time.perf_counter() - t0) / n
            agg = torch.tensor([t, float(stt.get("h2d_bytes", 0))], dtype=torch.float64, device="cuda")
            if world > 1:
                tmax = agg.clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(agg, op=dist.ReduceOp.SUM)
                return tmax[0].item(), agg[1].item(), res
            return agg[0].item(), agg[1].item(), res

        e2e_call(host)                                                       # warm-up
        t_e2e, h2d, res = timed_e2e(host, n_e2e)
        small = world * (C_ao_lo_h.nbytes + basis_h.nbytes)
        e2e = {"value": (F1 + F3) / t_e2e / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": int(h2d + small),
               "d2h_bytes_per_step": int(npair * npair * 8 * (nspin * (nspin + 1) // 2)), "seconds_per_step": t_e2e,
               "steps": n_e2e, "h2d_gbs_aggregate": h2d / t_e2e / 1e9, "numa": numa,
               "note": "get_emb_eri(cell, host_provider, numpy C_ao_lo, numpy basis%s) -> numpy on rank 0; every "
                       "(ki,kj) block is copied from pinned host memory inside the call (PCIe-bound at this shape: "
                       "the GDF tensor is %.0f GB)" % (", use_mpi=True" if world > 1 else "", B * blk_bytes / 1e9)}
        if not args.no_parity:
            cache = {}

            def host_block(ki, kj, a, b):
                s = pooled_e2e.pool_index(ki, kj)
                if s not in cache:
                    cache[s] = host.pool[s].cuda()
                return cache[s][a:b]
            indep = independent_sample(torch, dist, world, schedule, my_items, host_block, C_ao_lo_h, basis_h, kmesh,
                                       gdf.kpts_scaled, orbs, nspin)
            cache.clear()
            if rank == 0:
                ii = sidx.cpu().numpy()
                e2e["parity"] = parity_report(res[:, ii][:, :, ii], indep.cpu().numpy(), (digest or {}).get("e2e"))
                e2e["parity"]["pool_blocks"] = E2E_POOL
        del res
        # steady state of a DMET loop: the GDF tensor is constant, ResidentGDF keeps what fits in HBM across calls
        if not args.no_steady:
            torch.cuda.empty_cache()
            resident = et.ResidentGDF(host)
            e2e_call(resident)                                               # first iteration: fills the device store
            t_st, h2d_st, res = timed_e2e(resident, n_e2e)
            cached = torch.tensor([float(resident.bytes_cached)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(cached, op=dist.ReduceOp.SUM)
            e2e["steady"] = {"value": (F1 + F3) / t_st / 1e12, "unit": "TFLOP/s", "seconds_per_step": t_st,
                             "h2d_bytes_per_step": int(h2d_st + small), "resident_bytes": int(cached.item()),
                             "resident_fraction_of_tensor": cached.item() / float(B * blk_bytes),
                             "vs_device_resident_seconds": t_st / t_step,
                             "note": "2nd+ call through eri_transform.ResidentGDF (blocks served once stay in HBM up "
                                     "to the budget; the rest keeps streaming over PCIe); reported beside the cold "
                                     "e2e, not instead of it"}
            resident.release()
            del res, resident
        del host
        torch.cuda.empty_cache()

    # ---- GDF tensor read from a cderi FILE, cold page cache (N=1 only; configs[2] size) ----
    gdf_file = None
    if world == 1 and args.gdf_file:
        gdf_file = gdf_file_leg(args)

    # ---- one DMET iteration of this path: get_emb_basis + embHam through the public API (N=1 only) ----
    dmet_iter = None
    if world == 1 and not args.no_dmet and nspin == 1:
        from libdmet_preview_b200 import lattice as lat, slater
        torch.cuda.empty_cache()
        # the GDF tensor of the timed legs (pool of distinct blocks), kept resident in HBM across iterations the way a
        # DMET loop keeps it (eri_transform.ResidentGDF: a pool block shared by several pairs is held once)
        base = et.ResidentGDF(gdf)
        nval = neo - nao // 2 if neo > nao // 2 else max(1, neo // 3)       # impurity = nao/2 orbitals + nval bath
        nimp = neo - nval
        Lat = lat.Lattice(base.cell, kmesh)
        Lat.set_val_virt_core(nval, nimp - nval, nao - nimp)
        hcore = synthetic.make_hermitian_k(kmesh, nao, seed=41)
        vhf = synthetic.make_hermitian_k(kmesh, nao, seed=42, scale=0.3)
        rdm1 = synthetic.make_rdm1_k(hcore + vhf, max(1, nao // 3)) * 2.0
        ovlp = np.asarray([np.eye(nao, dtype=np.complex128)] * len(base.kpts_scaled))
        t0 = time.perf_counter()
        Lat.set_Ham(None, base, C_ao_lo_h, eri_symmetry=4, ovlp=ovlp, hcore=hcore, rdm1=rdm1, vhf=vhf)
        torch.cuda.synchronize()
        t_set = time.perf_counter() - t0
        # three iterations: the first fills the resident store and pays the one-off costs (page-locking the 1 GB
        # result buffer, pipeline workspaces), the last is what every further DMET iteration costs
        first = None
        for it in range(3):
            t0 = time.perf_counter()
            bas = slater.get_emb_basis(Lat, Lat.rdm1_lo_R * 0.5)
            t_basis = time.perf_counter() - t0
            t0 = time.perf_counter()
            stt = {}
            Ham, _ = slater.embHam(Lat, bas, None, group=args.group, kl_group=args.kl_group, stats=stt)
            torch.cuda.synchronize()
            t_ham = time.perf_counter() - t0
            if first is None:
                first = t_basis + t_ham
            del Ham
        dmet_iter = {"seconds": t_basis + t_ham, "get_emb_basis_s": t_basis, "embHam_s": t_ham,
                     "first_iteration_s": first, "set_Ham_once_s": t_set, "neo": int(bas.shape[-1]),
                     "resident_bytes": int(base.bytes_cached), "h2d_bytes": int(stt.get("h2d_bytes", 0)),
                     "note": "ConstructImpHam of this path (libdmet/dmet/HubPhSymm.py:74-100): bath SVD on the host, "
                             "ERI build over the GDF tensor resident in HBM (ResidentGDF; the first iteration fills "
                             "the store), one-body part and J/K on the device, results returned as numpy"}
        base.release()
        del base, Lat

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu and world == 1 and nspin == 1:       # reported at N=1 only (rank 0's host cores are otherwise shared)
        c = cpu_sample(kmesh, nao, naux, neo)
        cpu = {"value": c["tflops"], "unit": "TFLOP/s", "cores": _host_cores(), "kind": "port", "sample": c["sample"],
               "blas_threads": blas_threads(), "seconds_whole_job_extrapolated": c["t_job_s"],
               "sample_seconds": c["sample_seconds"]}

    line = {"metric": "get_emb_eri_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (seeded counter-based GDF tensor generated on the device: L(ki,kj) = pool block "
                    "(ki nk + kj) mod %d, the same tensor for every N; this rank holds %d resident pool pieces for the "
                    "%d block pieces of its schedule)" % (npool, nslots, len(my_blocks)),
            "config": config_dict(args, kmesh, nao, naux, neo, B, G),
            "parallelism": "(kL, aux-range) work items sharded x%d, one NCCL reduce of the s4 ERI" % world,
            "tuning": {"group": args.group, "kl_group": args.kl_group},
            "get_emb_eri_seconds": t_step, "flops_per_step": F1 + F3, "clocks": clk.summary(),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "gdf_file": gdf_file,
            "dmet_iter": dmet_iter, "gpu_launches": int(launches), "checksum": checksum}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gdf_file_leg(args):
    """get_emb_eri with the GDF tensor read from a PySCF-layout cderi FILE (the reference's real input,
    eri_transform.py:195-227) at the configs[2] size: first from a cold page cache (the file's pages are dropped with
    posix_fadvise after an fsync), then warm, then through ResidentGDF.  Wall clock, host numpy in and out."""
    import tempfile
    import torch
    from libdmet_preview_b200 import synthetic, eri_transform as et
    from libdmet_preview_b200.gdf_file import GDFFile, write_gdf_file
    kmesh, nao, naux, neo, nspin = workload("c3_nio_uhf")
    g = synthetic.SyntheticGDF(kmesh, nao, naux, seed=7)
    C, basis = make_inputs(kmesh, nao, neo, nspin)
    F1, F3, B, G = flops(kmesh, nao, naux, neo, nspin)
    d = tempfile.mkdtemp(prefix="ldm_cderi_", dir=args.gdf_dir)
    path = os.path.join(d, "cderi.h5")
    try:
        write_gdf_file(path, g)
        size = os.path.getsize(path)

        def drop_cache():
            fd = os.open(path, os.O_RDONLY)
            try:
                os.fsync(fd)
                os.posix_fadvise(fd, 0, 0, os.POSIX_FADV_DONTNEED)
            finally:
                os.close(fd)

        def cached_fraction():
            try:
                out = subprocess.check_output(["fincore", "--bytes", "--noheadings", "--output", "RES", path],
                                              text=True, timeout=10)
                return float(out.split()[0]) / size
            except Exception:       # noqa: BLE001
                return None

        def call(prov):
            torch.cuda.synchronize()
            st = {}
            t = time.perf_counter()
            e = et.get_emb_eri(g.cell, prov, C_ao_lo=C, basis=basis, stats=st)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            # the result lives in page-locked memory that torch hands out again once the array is dropped (as a DMET
            # loop does with the previous iteration's integrals); keeping it would make the NEXT call page-lock a
            # fresh 0.8 GB buffer inside its timed region, so a pageable copy is kept for the comparisons instead
            keep = np.array(e, copy=True)
            del e
            return dt, keep, st

        call(g)                                     # warm the pipeline workspaces on the device-generated tensor
        t_mem, e_mem, _ = call(g)
        drop_cache()
        frac = cached_fraction()
        t_cold, e_cold, st_cold = call(GDFFile(path, cell=g.cell, kpts=g.kpts))
        t_warm, e_warm, st_warm = call(GDFFile(path, cell=g.cell, kpts=g.kpts))
        res = et.ResidentGDF(GDFFile(path, cell=g.cell, kpts=g.kpts))
        call(res)
        t_res, e_res, _ = call(res)
        res.release()
        return {"workload": "c3_nio_uhf: kmesh 2x2x2 nao %d naux %d neo %d unrestricted, cderi file %.2f GB (PySCF v1 "
                            "layout)" % (nao, naux, neo, size / 1e9),
                "cold_seconds": t_cold, "warm_seconds": t_warm, "resident_seconds": t_res,
                "device_generated_seconds": t_mem, "cold_tflops": (F1 + F3) / t_cold / 1e12,
                "cold_file_gbs": size / t_cold / 1e9, "h2d_bytes": st_cold.get("h2d_bytes"),
                "cold_wait_for_file_s": st_cold.get("provider_wait_s"), "cold_block_calls_s": st_cold.get("block_call_s"),
                "warm_wait_for_file_s": st_warm.get("provider_wait_s"), "warm_block_calls_s": st_warm.get("block_call_s"),
                "page_cache_fraction_before_cold_run": frac,
                "cold_equals_warm_bitwise": bool(np.array_equal(e_cold, e_warm)),
                "max_abs_vs_device_generated": float(np.abs(e_cold - e_mem).max()),
                "max_abs_resident_vs_cold": float(np.abs(e_res - e_cold).max()),
                "dir": args.gdf_dir}
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--group", type=int, default=0, help="blocks per stage-1 launch (0 = auto)")
    ap.add_argument("--kl-group", dest="kl_group", type=int, default=0, help="momenta per stage-3 launch (0 = auto)")
    ap.add_argument("--store-slots", dest="store_slots", type=int, default=0)
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=1)
    ap.add_argument("--no-e2e", dest="no_e2e", action="store_true")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--no-dmet", dest="no_dmet", action="store_true")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true")
    ap.add_argument("--no-peak", dest="no_peak", action="store_true", help="skip the live FP64 DGEMM probe")
    ap.add_argument("--no-steady", dest="no_steady", action="store_true")
    ap.add_argument("--no-numa", dest="no_numa", action="store_true")
    ap.add_argument("--gdf-file", dest="gdf_file", action="store_true",
                    help="add the disk-backed leg (cderi file at the configs[2] size, cold page cache)")
    ap.add_argument("--gdf-dir", dest="gdf_dir", default=os.environ.get("TMPDIR", "/tmp"))
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
