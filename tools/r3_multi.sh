#!/bin/bash
# Round 2, final code at N GPUs:  gpurun --gpus N --timeout 900 -- 'bash tools/r3_multi.sh N'
# NCCL parity test on N ranks, then the full bench line (device-resident leg, e2e cold + steady, parity report).
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -q -s 2>&1 | tail -12 | tee $O/r3q_dist_test_$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --steps 3 --warmup 3 2>$O/r3q_bench_$N.err | tail -1 | tee $O/r3q_bench_$N.json | cut -c1-300
echo done
