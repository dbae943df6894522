#!/bin/bash
# Round 2, profiles of the final code: launch list of one build with per-launch DRAM bytes, full ncu capture of one
# stage-1a launch (pair-ordered main loop), small-shape launch lists.
mkdir -p gpurun_out
O=gpurun_out
timeout 700 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file $O/r3m_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-dmet --no-peak --no-parity > $O/r3m_ncu_bench.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:zgemm_tn_kernel -s 6 -c 1 -o $O/r3m_zgemm python tools/zcfg_bench.py > $O/r3m_ncu_zgemm.log 2>&1
ncu -i $O/r3m_zgemm.ncu-rep --page details --csv > $O/r3m_zgemm_details.csv 2>/dev/null
ncu -i $O/r3m_zgemm.ncu-rep --page raw --csv > $O/r3m_zgemm_raw.csv 2>/dev/null
ncu -i $O/r3m_zgemm.ncu-rep --page source --csv --print-source sass > $O/r3m_zgemm_source.csv 2>/dev/null
rm -f $O/r3m_zgemm.ncu-rep
ls -la $O/r3m*
echo done
