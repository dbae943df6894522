#!/bin/bash
# Round 2, first GPU call (1 GPU):  gpurun --timeout 1700 -- 'bash tools/r2_call1.sh'
# whole GPU suite (incl. the real-size parity tests), HBM kernel timings, headline bench with the disk-backed leg,
# ncu launch list + per-kernel DRAM bytes, racecheck / synccheck on small shapes.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r2_box.txt 2>&1
nproc >> $O/r2_box.txt; free -g >> $O/r2_box.txt; numactl -H >> $O/r2_box.txt 2>&1
# 1. GPU test suite
timeout 1000 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 | tee $O/r2_gpu_tests.log
# 2. HBM-bound kernels, CUDA events
timeout 200 python tools/bench_hbm_kernels.py 2>&1 | tee $O/r2_hbm_kernels.txt
# 3. headline bench line (with the cderi-file leg)
timeout 900 python bench.py --steps 3 --warmup 3 --gdf-file 2>$O/r2_bench.err | tail -1 | tee $O/r2_bench_1gpu.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>$O/r2_bench_ref.err | tail -1 | tee $O/r2_bench_reference.json
# 4. ncu: launch list of a short bench run, per-kernel DRAM bytes of the HBM kernels
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-dmet --no-peak --no-parity > $O/r2_ncu_bench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'lattice_dft|mirror_lower|jk_rows|restore_s|ztranspose|phase_transform|unpack_stored' -c 60 --csv \
    --log-file $O/r2_hbm_ncu.csv python tools/bench_hbm_kernels.py > $O/r2_hbm_ncu.log 2>&1
# 5. sanitizers on small shapes
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q \
    -k "zgemm_tn or dgemm_tn_and_mirror or transpose or restore_and_jk or pipeline_many" 2>&1 | tail -12 | tee $O/r2_racecheck.txt
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q \
    -k "zgemm_tn or dgemm_tn_and_mirror or restore_and_jk or jk_streaming" 2>&1 | tail -12 | tee $O/r2_synccheck.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_zz_gpu_gdf_file.py \
    tests/test_gpu_fourier_basis.py -x -q -k "unpack_stored or host_and_device or outcore or fourier or r2k" 2>&1 | tail -12 | tee $O/r2_memcheck.txt
echo done
