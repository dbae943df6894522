#!/bin/bash
# First gpurun call of the next round: the measurements round 1 had no GPU minutes left for.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Everything lands in gpurun_out/ (copy what is to be judged into profiles/ afterwards).
mkdir -p gpurun_out
# 1. whole GPU suite on the final code of round 1
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_gpu_tests.log
# 2. memcheck over the stored-entry unpack kernel and the file path (new at the end of round 1)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_zz_gpu_gdf_file.py -x -q \
    -k "unpack_stored or host_and_device or outcore" \
    2>&1 | tail -8 | tee gpurun_out/r2_sanitizer_gdf_file.txt
# 3. disk-backed build at a larger shape (2x2x2, nao 100, naux 400: 36 stored pairs of 64 MB, 2.3 GB file)
timeout 300 python tools/bench_gdf_file.py --kmesh 2 2 2 --nao 100 --naux 400 --neo 80 --dir /tmp --reps 2 \
    2>gpurun_out/r2_gdf_file.err | tee gpurun_out/r2_gdf_file_222.json
# 4. headline bench line with the final code
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/r2_bench.err | tail -1 | tee gpurun_out/r2_bench_1gpu.json
