#!/bin/bash
# Round 2, second GPU call (1 GPU): full GPU suite after the fixes, HBM kernels again, stage-1 GEMM with and without
# the padding-fragment skip, device-resident bench, ncu --set full on the kernels under work.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -30 | tee $O/r2b_gpu_tests.log
timeout 200 python tools/bench_hbm_kernels.py 2>&1 | tee $O/r2b_hbm_kernels.txt
for sp in 1 0; do
  LDM_ZGEMM_SKIP_PAD=$sp timeout 200 python tools/zcfg_bench.py 2>&1 | tee -a $O/r2b_zcfg.txt
done
BENCH_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-dmet --no-peak 2>$O/r2b_bench.err | tail -1 | tee $O/r2b_bench_dev.json
LDM_ZGEMM_SKIP_PAD=0 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-dmet --no-peak --no-parity 2>/dev/null | tail -1 | tee $O/r2b_bench_dev_nopadskip.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'lattice_dft|pack_sym|jk_rows_bulk|jk_tri_kernel' -c 14 \
    -o $O/r2b_hbm_full python tools/ncu_targets.py > $O/r2b_ncu_targets.log 2>&1
ncu -i $O/r2b_hbm_full.ncu-rep --page raw --csv > $O/r2b_hbm_full_raw.csv 2>/dev/null
timeout 300 python bench.py --workload c3_nio_uhf --gdf-file --steps 2 --warmup 1 --no-e2e --no-cpu --no-dmet --no-peak 2>/dev/null | tail -1 | tee $O/r2b_bench_c3_gdffile.json
echo done
