"""Aggregate pinned host->device bandwidth with one process per GPU copying at the same time -- the ceiling of the
end-to-end (host-streamed) leg of bench.py at N GPUs.  Run under torch.distributed.run; prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe_h2d.py [--no-numa]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
numa = {"bound": False, "why": "--no-numa"} if "--no-numa" in sys.argv else bench.bind_to_gpu_numa_node(lr)
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nbytes = 640 * 1000 * 1000
src = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
for s in src:
    s.fill_(1)                      # first touch after the affinity call
dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
def run(n):
    with torch.cuda.stream(st):
        for i in range(n):
            dst.copy_(src[i % 4], non_blocking=True)
    st.synchronize()
run(4)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
run(24)
dt = time.perf_counter() - t0
mine = 24 * nbytes / dt / 1e9
t = torch.tensor([mine], dtype=torch.float64, device="cuda")
lo, hi, tot = t.clone(), t.clone(), t.clone()
if world > 1:
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, numa)
else:
    gathered = [numa]
if rank == 0:
    print(json.dumps({"tool": "probe_h2d", "n_gpus": world, "per_rank_gbs_min": lo.item(), "per_rank_gbs_max": hi.item(),
                      "aggregate_gbs": tot.item(), "host_cores": os.cpu_count(), "numa": gathered}), flush=True)
if world > 1:
    dist.destroy_process_group()
