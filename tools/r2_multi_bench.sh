#!/bin/bash
# bench at N GPUs only:  gpurun --gpus N -- 'bash tools/r2_multi_bench.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/r2n_bench_$N.err | tail -1 | tee gpurun_out/r2n_bench_$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/r2n_bench_ref_$N.json
echo done
