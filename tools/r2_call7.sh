#!/bin/bash
# Round 2, seventh GPU call (1 GPU): full GPU suite on the current code, stage-1 GEMM micro-benchmark, headline bench,
# launch list and per-kernel DRAM traffic under ncu.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/zcfg_bench.py 2>&1 | tee $O/r2g_zcfg.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -14 | tee $O/r2g_gpu_tests.log
timeout 300 python tools/profile_small.py 2>&1 | grep "per call" | tee $O/r2g_small.txt
timeout 900 python bench.py --steps 3 --warmup 3 --gdf-file 2>$O/r2g_bench.err | tail -1 | tee $O/r2g_bench_1gpu.json
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file $O/r2g_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-dmet --no-peak --no-parity > $O/r2g_ncu_bench.log 2>&1
echo done
