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
def cpu_sample(kmesh, nao, naux, neo, nblocks=2, budget_s=25.0):
    """Time the oracle's stage 1 (transform_ao_to_emb + hermi_sum + pack_tril + accumulate, in the reference's
    240-row chunks) on `nblocks` blocks and one stage-3 Gram product; scale to the whole job."""
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
    tg = time.perf_counter()
    olib.dot(X.T, X, 1.0, eri, 1)
    t_gram = time.perf_counter() - tg
    t_job = t_block * B + t_gram * G
    return {"t_block_s": t_block, "t_gram_s": t_gram, "t_job_s": t_job, "tflops": (F1 + F3) / t_job / 1e12,
            "sample": "%d of %d (ki,kj) blocks of stage 1 (240-row chunks) + 1 of %d Gram products, extrapolated "
                      "linearly by block/Gram count" % (len(t_blocks), B, G)}


def run_reference(args):
    kmesh, nao, naux, neo, nspin = workload(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    F1, F3, B, G = flops(kmesh, nao, naux, neo)      # the CPU sample times the restricted path
    times = []
    res = None
    for it in range(args.warmup + args.steps):
        res = cpu_sample(kmesh, nao, naux, neo, nblocks=1 if it < args.warmup else 2, budget_s=20.0)
        if it >= args.warmup:
            times.append(res["t_job_s"])
    t = float(np.mean(times))
    val = (F1 + F3) / t / 1e12
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    line = {"impl": "reference", "metric": "get_emb_eri_fp64_tflops", "value": val, "unit": "TFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, kmesh, nao, naux, neo, B, G),
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                             "sample": res["sample"] + "; numpy/OpenBLAS zgemm+dgemm, all host threads; the "
                             "reference itself cannot be imported (PySCF, h5py absent)"},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "get_emb_eri_seconds": t}
    print(json.dumps(line), flush=True)


def config_dict(args, kmesh, nao, naux, neo, B, G):
    kind = "restricted" if workload(args.workload)[4] == 1 else "unrestricted"
    return {"workload": "%s: get_emb_eri GDF %s s4 time-reversal, kmesh %s nkpts %d nao %d naux %d neo %d "
                        "(%d (ki,kj) blocks, %d Gram products)" % (args.workload, kind, "x".join(map(str, kmesh)),
                                                                  int(np.prod(kmesh)), nao, naux, neo, B, G),
            "kmesh": kmesh, "nao": nao, "naux": naux, "neo": neo, "symmetry": 4, "t_reversal_symm": True}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class HostPoolProvider(object):
    """GDF provider over pinned host memory: `npool` distinct synthetic blocks, block (ki,kj) -> pool slot by a fixed
    map.  What a host-RAM resident cderi looks like to `get_emb_eri` (the full 758 GB tensor of the target shape is
    not materialised on the host; every block still crosses PCIe inside the timed region)."""

    def __init__(self, gdf, npool):
        import torch
        from libdmet_preview_b200.device import get_device
        dev = get_device()
        self.kpts_scaled, self.kmesh, self.nao, self.naux = gdf.kpts_scaled, gdf.kmesh, gdf.nao, gdf.naux
        self.kpts, self.cell = gdf.kpts, gdf.cell
        nk = len(self.kpts_scaled)
        self.npool = npool
        self.pool = torch.empty((npool, gdf.naux, gdf.nao, gdf.nao), dtype=torch.complex128, pin_memory=True)
        tmp = dev.empty((gdf.naux, gdf.nao, gdf.nao), torch.complex128)
        for s in range(npool):
            dev.synth_block(tmp, gdf.naux, gdf.nao, gdf.keys(s % nk, (s // nk) % nk), gdf.scale)
            self.pool[s].copy_(tmp)
        dev.synchronize()
        self.nk = nk

    def load(self, ki, kj):
        return self.pool[(ki * self.nk + kj) % self.npool]


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
    gdf = synthetic.SyntheticGDF(kmesh, nao, naux, seed=2026)
    C_ao_lo_h = synthetic.make_C_ao_lo(kmesh, nao, seed=1, spin=(nspin if nspin > 1 else None))
    basis_h = synthetic.make_emb_basis(kmesh, nao, neo, seed=2, spin=nspin)
    C_ao_lo = dev.to_device(C_ao_lo_h, torch.complex128)
    basis = dev.to_device(basis_h, torch.float64)
    schedule = build_schedule(gdf.kpts_scaled, True)
    my_items = ldist.rank_items(schedule, nao, naux, neo, nspin, world)[rank]
    my_blocks = [(l0, l1, blk) for (u, l0, l1) in my_items for blk in schedule.units[u][2]]
    my_rows = sum(l1 - l0 for (l0, l1, _) in my_blocks)           # GDF rows this rank transforms

    # ---- resident store of L blocks (inputs in HBM before the timed region) ----
    free_b, total_b = torch.cuda.mem_get_info()
    npair = neo * (neo + 1) // 2
    work_b = nspin * (args.group * naux * neo * nao * 16 + 2 * naux * neo * neo * 16 +
                      npair * args.kl_group * 2 * naux * 8 + 3 * npair * npair * 8) + (6 << 30)
    ranges = sorted({(l0, l1) for (l0, l1, _) in my_blocks})
    budget = free_b - work_b
    stores, store_map, nslots_tot = {}, {}, 0
    for (l0, l1) in ranges:
        blks = [b for (a0, a1, b) in my_blocks if (a0, a1) == (l0, l1)]
        blk_bytes = (l1 - l0) * nao * nao * 16
        share = budget * (len(blks) * (l1 - l0)) / float(max(1, my_rows))
        nslots = int(max(1, min(len(blks), share // blk_bytes)))
        if args.store_slots:
            nslots = min(nslots, args.store_slots)
        st_t = dev.empty((nslots, l1 - l0, nao, nao), torch.complex128)
        for n, (ki, kj, sym) in enumerate(blks):
            if n < nslots:
                dev.synth_block(st_t[n], l1 - l0, nao, gdf.keys(ki, kj), gdf.scale, aux_offset=l0)
            store_map[(ki, kj, l0)] = n % nslots
        stores[(l0, l1)] = st_t
        nslots_tot += nslots
    nslots = nslots_tot
    blk_bytes = naux * nao * nao * 16
    dev.synchronize()

    stats = {}

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
    del out
    tt = torch.tensor([t_dev, zg_ms, dg_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_step = tt[0].item() / args.steps
    value = (F1 + F3) / t_step / 1e12

    # ---- roofline of the dominant kernel (stage-1 zgemm), events recorded on its launch stream ----
    F1_mine = F1 * my_rows / float(B * naux)
    with open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")) as f:
        pk = json.load(f)
    peak = pk["dgemm_8192_sustained_tflops"]
    ach = F1_mine * args.steps / (zg_ms * 1e-3) / 1e12 if zg_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "zgemm_traffic_r01.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    m3 = os.environ.get("LDM_ZGEMM_3M", "1") != "0"
    roofline = {"bound": "tensor", "kernel": "zgemm_tn_kernel (stage 1: both half transformations)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic,
                "complex_product": "3M" if m3 else "4M",
                "executed_tflops": (ach * (0.75 if m3 else 1.0)) if ach else None,
                "note": ("achieved counts ALGORITHMIC flops (8 per complex multiply-add); the kernel evaluates complex "
                         "products with three real multiplications, so the tensor pipe executes 0.75 of them and "
                         "frac can exceed 1") if m3 else None,
                "peak_source": "cuBLAS DGEMM 8192^3 sustained 4 s on this pool's B200 (tools/probe_peaks.py -> "
                               "profiles/fp64_peaks_r01.json); MEASURED_PEAKS.json carries no FP64 figure",
                "share_of_step": zg_ms * 1e-3 / (t_step * args.steps),
                "stage3_share_of_step": dg_ms * 1e-3 / (t_step * args.steps),
                "stage3_dgemm": {"achieved": (F3 * sum((1 if schedule.units[u][1] == 1 else 2) * (l1 - l0)
                                                           for (u, l0, l1) in my_items) / float(G * naux)) *
                                 args.steps / (dg_ms * 1e-3) / 1e12 if dg_ms > 0 else None, "unit": "TFLOP/s"}}

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        stores.clear()
        torch.cuda.empty_cache()
        host = HostPoolProvider(gdf, args.host_pool if world == 1 else max(2, args.host_pool // 4))
        n_e2e = max(1, args.e2e_steps)
        st = {}

        def e2e_call(stats=None):
            kw = dict(C_ao_lo=C_ao_lo_h, basis=basis_h, symmetry=4, source="host", kl_group=args.kl_group, stats=stats)
            if world > 1:       # the reference's own keyword for its multi-process path (eri_transform.py:71)
                return et.get_emb_eri(gdf.cell, host, use_mpi=True, group_blocks=args.group, **kw)
            return et.get_emb_eri(gdf.cell, host, group=args.group, **kw)

        e2e_call()                                                       # warm-up
        barrier()
        t0 = time.perf_counter()
        res = None
        for _ in range(n_e2e):
            del res                      # a DMET loop drops the previous eri too: the pinned result buffer is reused
            res = e2e_call(st)
        barrier()
        t_e2e = (time.perf_counter() - t0) / n_e2e
        agg = torch.tensor([t_e2e, float(st.get("h2d_bytes", 0))], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = agg.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            t_e2e, h2d = tmax[0].item(), agg[1].item()
        else:
            h2d = agg[1].item()
        e2e = {"value": (F1 + F3) / t_e2e / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": int(h2d + world * (C_ao_lo_h.nbytes + basis_h.nbytes)),
               "d2h_bytes_per_step": int(npair * npair * 8), "seconds_per_step": t_e2e, "steps": n_e2e,
               "note": "get_emb_eri(cell, host_provider, numpy C_ao_lo, numpy basis%s) -> numpy on rank 0; every "
                       "(ki,kj) block is copied from pinned host memory inside the call (PCIe-bound at this shape: "
                       "the GDF tensor is 758 GB)" % (", use_mpi=True" if world > 1 else "")}
        del host

    # ---- one DMET iteration of this path: get_emb_basis + embHam through the public API (N=1 only) ----
    dmet_iter = None
    if world == 1 and not args.no_dmet and nspin == 1:
        from libdmet_preview_b200 import lattice as lat, slater
        torch.cuda.empty_cache()
        nval = neo - nao // 2 if neo > nao // 2 else max(1, neo // 3)       # impurity = nao/2 orbitals + nval bath
        nimp = neo - nval
        Lat = lat.Lattice(gdf.cell, kmesh)
        Lat.set_val_virt_core(nval, nimp - nval, nao - nimp)
        hcore = synthetic.make_hermitian_k(kmesh, nao, seed=41)
        vhf = synthetic.make_hermitian_k(kmesh, nao, seed=42, scale=0.3)
        rdm1 = synthetic.make_rdm1_k(hcore + vhf, max(1, nao // 3)) * 2.0
        ovlp = np.asarray([np.eye(nao, dtype=np.complex128)] * len(gdf.kpts_scaled))
        t0 = time.perf_counter()
        Lat.set_Ham(None, gdf, C_ao_lo_h, eri_symmetry=4, ovlp=ovlp, hcore=hcore, rdm1=rdm1, vhf=vhf)
        torch.cuda.synchronize()
        t_set = time.perf_counter() - t0
        # two iterations: the first pays one-off costs (page-locking the 1 GB result buffer, pipeline workspaces),
        # the second is what every further DMET iteration costs
        first = None
        for it in range(2):
            t0 = time.perf_counter()
            bas = slater.get_emb_basis(Lat, Lat.rdm1_lo_R * 0.5)
            t_basis = time.perf_counter() - t0
            t0 = time.perf_counter()
            Ham, _ = slater.embHam(Lat, bas, None, group=args.group, kl_group=args.kl_group)
            torch.cuda.synchronize()
            t_ham = time.perf_counter() - t0
            if first is None:
                first = t_basis + t_ham
            del Ham
        dmet_iter = {"seconds": t_basis + t_ham, "get_emb_basis_s": t_basis, "embHam_s": t_ham,
                     "first_iteration_s": first, "set_Ham_once_s": t_set, "neo": int(bas.shape[-1]),
                     "note": "ConstructImpHam of this path (libdmet/dmet/HubPhSymm.py:74-100): bath SVD on the host, "
                             "ERI build with the GDF blocks generated on the device, one-body part and J/K on the "
                             "device, results returned as numpy"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu and world == 1 and nspin == 1:       # reported at N=1 only (rank 0's host cores are otherwise shared)
        c = cpu_sample(kmesh, nao, naux, neo)
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        cpu = {"value": c["tflops"], "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": c["sample"],
               "seconds_whole_job_extrapolated": c["t_job_s"]}

    line = {"metric": "get_emb_eri_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (seeded counter-based GDF tensor generated on the device; %d resident blocks per GPU, "
                    "the %d block pieces of this rank's schedule cycle over them)" % (nslots, len(my_blocks)),
            "config": dict(config_dict(args, kmesh, nao, naux, neo, B, G), nspin=nspin, parallelism="(kL, aux-range) work items sharded x%d, one NCCL reduce of the s4 ERI" % world,
                           l2_policy="inputs larger than L2 (each GDF block %.0f MB, ERI %.0f MB)" %
                           (blk_bytes / 1e6, npair * npair * 8 / 1e6), group=args.group, kl_group=args.kl_group),
            "get_emb_eri_seconds": t_step, "flops_per_step": F1 + F3, "clocks": clk.summary(),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "dmet_iter": dmet_iter, "gpu_launches": int(launches),
            "checksum": checksum}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--host-pool", dest="host_pool", type=int, default=16)
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=1)
    ap.add_argument("--no-e2e", dest="no_e2e", action="store_true")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--no-dmet", dest="no_dmet", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
