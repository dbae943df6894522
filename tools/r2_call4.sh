#!/bin/bash
# Round 2, fourth GPU call (1 GPU): consumer-warp stagger of the stage-1 GEMM (several head starts), J/K again.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r2d_zcfg.txt
for sk in 0 2000 4000 6000 9000; do
  echo "skew $sk" | tee -a $O/r2d_zcfg.txt
  LDM_ZGEMM_SKEW=$sk timeout 200 python tools/zcfg_bench.py 2>&1 | tee -a $O/r2d_zcfg.txt
done
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "zgemm or pipeline or jk" 2>&1 | tail -5 | tee $O/r2d_tests.log
timeout 200 python tools/bench_hbm_kernels.py 2>&1 | grep -i "J" | tee $O/r2d_hbm_kernels.txt
for sk in 0 4000; do
  LDM_ZGEMM_SKEW=$sk timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-dmet --no-peak 2>/dev/null | tail -1 | tee $O/r2d_bench_skew$sk.json
done
timeout 300 python bench.py --workload c3_nio_uhf --gdf-file --steps 2 --warmup 1 --no-e2e --no-cpu --no-dmet --no-peak --no-parity 2>/dev/null | tail -1 | tee $O/r2d_bench_c3_gdffile.json
echo done
