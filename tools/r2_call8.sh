#!/bin/bash
# Round 2, eighth GPU call (1 GPU): compile-time short last n-tile in the 3M GEMM.
mkdir -p gpurun_out
O=gpurun_out
for sl in 1; do
  echo "short_last $sl" | tee -a $O/r2i_zcfg.txt
  LDM_ZGEMM_SHORT_LAST=$sl timeout 300 python tools/zcfg_bench.py 2>&1 | tee -a $O/r2i_zcfg.txt
done
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_real_sizes.py -q -x -k "not jk" 2>&1 | tail -4 | tee $O/r2i_tests.log
for sl in 1; do
  LDM_ZGEMM_SHORT_LAST=$sl timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-dmet --no-peak 2>/dev/null | tail -1 | tee $O/r2i_bench_sl$sl.json
done
echo done
