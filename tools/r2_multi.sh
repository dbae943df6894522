#!/bin/bash
# Round 2, multi-GPU call:  gpurun --gpus N --timeout 1500 -- 'bash tools/r2_multi.sh N'
# NCCL parity test on N ranks (tests/test_gpu_dist.py), aggregate H2D probe, bench at N GPUs with its parity report.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2m_topo_$N.txt 2>&1
nproc >> $O/r2m_topo_$N.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -q -s 2>&1 | tail -25 | tee $O/r2m_dist_test_$N.txt
for flag in "" "--no-numa"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      tools/probe_h2d.py $flag 2>/dev/null | tail -1 | tee -a $O/r2m_probe_h2d_$N.json
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --steps 3 --warmup 3 2>$O/r2m_bench_$N.err | tail -1 | tee $O/r2m_bench_$N.json
echo done
