#!/bin/bash
# Round 2: validation of the final stage-1 kernel (short last n-tile + staged epilogue): full GPU suite, racecheck /
# synccheck / memcheck on the kernel tests, micro-benchmark, headline bench, launch list with DRAM bytes.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/zcfg_bench.py 2>&1 | tee $O/r2j_zcfg.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 | tee $O/r2j_gpu_tests.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q \
    -k "zgemm_tn or dgemm_tn_and_mirror or transpose or restore_and_jk or pipeline_many or strided or many_n_tiles" 2>&1 | tail -8 | tee $O/r2j_racecheck.txt
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q \
    -k "zgemm_tn or dgemm_tn_and_mirror or restore_and_jk or jk_streaming or jk_lower or strided" 2>&1 | tail -8 | tee $O/r2j_synccheck.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py tests/test_zz_gpu_gdf_file.py -x -q \
    -k "zgemm or unpack_stored or jk_lower or time_reversal_reduced or strided or many_n_tiles" 2>&1 | tail -8 | tee $O/r2j_memcheck.txt
timeout 900 python bench.py --steps 3 --warmup 3 --gdf-file 2>$O/r2j_bench.err | tail -1 | tee $O/r2j_bench_1gpu.json
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file $O/r2j_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-dmet --no-peak --no-parity > $O/r2j_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:zgemm_tn_kernel -s 6 -c 2 -o $O/r2j_zgemm python tools/zcfg_bench.py > $O/r2j_ncu_zgemm.log 2>&1
ncu -i $O/r2j_zgemm.ncu-rep --page details --csv > $O/r2j_zgemm_details.csv 2>/dev/null
rm -f $O/r2j_zgemm.ncu-rep
echo done
