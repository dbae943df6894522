#!/bin/bash
# Round 2, sixth GPU call (1 GPU): 3M GEMM with k-permuted B planes (one LDS.128 per plane and fragment).
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -q -x 2>&1 | tail -4 | tee $O/r2f_tests.log
timeout 300 python tools/zcfg_bench.py 2>&1 | tee $O/r2f_zcfg.txt
NEO=100 timeout 300 python tools/zcfg_bench.py 2>&1 | tee -a $O/r2f_zcfg.txt
NEO=200 timeout 300 python tools/zcfg_bench.py 2>&1 | tee -a $O/r2f_zcfg.txt
timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-dmet --no-peak 2>/dev/null | tail -1 | tee $O/r2f_bench_dev.json
echo done
