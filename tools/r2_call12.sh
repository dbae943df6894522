#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/zcfg_bench.py 2>&1 | tee $O/r2k_zcfg.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q \
    -k "zgemm_tn or dgemm_tn_and_mirror or transpose or restore_and_jk or pipeline_many or strided or many_n_tiles" 2>&1 | tail -8 | tee $O/r2k_racecheck.txt
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_real_sizes.py -q -x 2>&1 | tail -3 | tee $O/r2k_tests.log
echo done
