#!/bin/bash
# Round 2, last validation of the final code (pair-ordered 3M main loop, 64x64 stage-3 tiles, ResidentGDF block keys):
# sanitizers on the kernel tests, headline bench with the disk-backed leg, reference arm.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q \
    -k "zgemm_tn or dgemm_tn_and_mirror or transpose or restore_and_jk or pipeline_many or strided or many_n_tiles" 2>&1 | tail -8 | tee $O/r3l_racecheck.txt
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q \
    -k "zgemm_tn or dgemm_tn_and_mirror or restore_and_jk or jk_streaming or jk_lower or strided" 2>&1 | tail -8 | tee $O/r3l_synccheck.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py tests/test_zz_gpu_gdf_file.py -x -q \
    -k "zgemm or dgemm or unpack_stored or jk_lower or time_reversal_reduced or strided or many_n_tiles" 2>&1 | tail -8 | tee $O/r3l_memcheck.txt
timeout 900 python bench.py --steps 3 --warmup 3 --gdf-file 2>$O/r3l_bench.err | tail -1 | tee $O/r3l_bench_1gpu.json | cut -c1-400
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>$O/r3l_bench_ref.err | tail -1 | tee $O/r3l_bench_reference.json | cut -c1-400
echo done
