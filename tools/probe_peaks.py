"""Measure the FP64 roofline denominators on the GPU box (cuBLAS DGEMM / ZGEMM through torch.matmul),
host core count / RAM and pinned H2D/D2H bandwidth.  Writes gpurun_out/fp64_peaks.json.
Run: gpurun -- python tools/probe_peaks.py"""
import json, os, subprocess, time
import torch

def clocks():
    try:
        out = subprocess.check_output(
            ["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
             "--format=csv,noheader"], text=True).strip().splitlines()[0]
        return out
    except Exception as e:  # pragma: no cover
        return repr(e)

def bench_mm(dtype, n, flop_per_mac, seconds=None, reps=10):
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    flops = flop_per_mac * n ** 3
    if seconds is None:
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return flops / best / 1e12, None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cnt = 0
    t0 = time.time()
    e0.record()
    mid = None
    while time.time() - t0 < seconds:
        for _ in range(5):
            torch.matmul(a, b); cnt += 1
        torch.cuda.synchronize()
        if mid is None and time.time() - t0 > seconds / 2:
            mid = clocks()
    e1.record(); torch.cuda.synchronize()
    return flops * cnt / (e0.elapsed_time(e1) * 1e-3) / 1e12, mid

res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cpu_count": os.cpu_count(),
       "clocks_idle": clocks()}
try:
    res["sched_affinity"] = len(os.sched_getaffinity(0))
except Exception:
    pass
with open("/proc/meminfo") as f:
    res["meminfo"] = f.readline().strip()
res["dgemm_8192_burst_tflops"], _ = bench_mm(torch.float64, 8192, 2)
res["dgemm_8192_sustained_tflops"], res["clocks_dgemm"] = bench_mm(torch.float64, 8192, 2, seconds=4)
res["zgemm_4096_burst_tflops"], _ = bench_mm(torch.complex128, 4096, 8)
res["zgemm_4096_sustained_tflops"], res["clocks_zgemm"] = bench_mm(torch.complex128, 4096, 8, seconds=4)
# shapes of the hot path: (200000 x 200) @ (200 x 150) complex and syrk-like real
a = torch.randn(200000, 200, device="cuda", dtype=torch.complex128)
b = torch.randn(200, 150, device="cuda", dtype=torch.complex128)
for _ in range(3): a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): a @ b
e1.record(); torch.cuda.synchronize()
res["zgemm_200000x200x150_tflops"] = 8 * 200000 * 200 * 150 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e12
del a, b
x = torch.randn(11325, 1000, device="cuda", dtype=torch.float64)
for _ in range(2): x @ x.T
torch.cuda.synchronize()
e0.record()
for _ in range(5): x @ x.T
e1.record(); torch.cuda.synchronize()
res["dgemm_11325x11325x1000_tflops"] = 2 * 11325 * 11325 * 1000 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e12
del x
# pinned host <-> device bandwidth
nbytes = 1 << 30
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for name, (src, dst) in {"h2d": (h, d), "d2h": (d, h)}.items():
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    e0.record()
    for _ in range(4): dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    res[name + "_pinned_gbs"] = 4 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
# device copy bandwidth (HBM denominator cross-check)
s = torch.empty(1 << 30, dtype=torch.float64, device="cuda"); t = torch.empty_like(s)
t.copy_(s); torch.cuda.synchronize()
e0.record()
for _ in range(5): t.copy_(s)
e1.record(); torch.cuda.synchronize()
res["hbm_copy_gbs"] = 5 * 2 * s.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/fp64_peaks.json", "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
