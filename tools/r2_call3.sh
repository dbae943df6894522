#!/bin/bash
# Round 2, third GPU call (1 GPU): GPU suite, HBM kernels, source-level ncu profile of the stage-1 GEMM.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 | tee $O/r2c_gpu_tests.log
timeout 200 python tools/bench_hbm_kernels.py 2>&1 | tee $O/r2c_hbm_kernels.txt
timeout 300 python bench.py --workload c3_nio_uhf --gdf-file --steps 2 --warmup 1 --no-e2e --no-cpu --no-dmet --no-peak --no-parity 2>/dev/null | tail -1 | tee $O/r2c_bench_c3_gdffile.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:zgemm_tn_kernel -s 6 -c 2 \
    -o $O/r2c_zgemm python tools/zcfg_bench.py > $O/r2c_ncu_zgemm.log 2>&1
ncu -i $O/r2c_zgemm.ncu-rep --page source --csv --print-source sass > $O/r2c_zgemm_source_sass.csv 2>/dev/null
ncu -i $O/r2c_zgemm.ncu-rep --page details --csv > $O/r2c_zgemm_details.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:'^ldm::pack_sym|pack_sym_kernel|jk_tri_kernel|jk_tri_jsum|jk_reduce' -c 8 \
    -o $O/r2c_pack python tools/ncu_targets.py > $O/r2c_ncu_pack.log 2>&1
ncu -i $O/r2c_pack.ncu-rep --page raw --csv > $O/r2c_pack_raw.csv 2>/dev/null
ls -la $O/*.ncu-rep
echo done
