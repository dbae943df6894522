#!/bin/bash
# Round 2, fifth GPU call (1 GPU): J/K kernels after pipelining the dd loads, host profile of small shapes.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "jk or zgemm or pipeline" 2>&1 | tail -3 | tee $O/r2e_tests.log
timeout 200 python tools/bench_hbm_kernels.py 2>&1 | tee $O/r2e_hbm_kernels.txt
timeout 300 python tools/profile_small.py 2>&1 | tee $O/r2e_profile_small.txt
timeout 300 python tools/zcfg_bench.py 2>&1 | tee $O/r2e_zcfg.txt
echo done
