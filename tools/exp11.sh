#!/bin/bash
mkdir -p gpurun_out
for mb in 0 32 64 200; do
for K in 4 8; do
UMT_VERBOSE=1 UMT_L2_PERSIST_MB=$mb UMT_ANGLE_BATCH=$K timeout 300 python tools/perf_sweep.py 20 128 2>&1 | tail -2
done; done
UMT_L2_PERSIST_MB=200 UMT_ANGLE_BATCH=8 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sweep3d_plan -s 1 -c 1 --csv --log-file gpurun_out/exp11_ncu.csv python tools/perf_sweep.py 20 128 > gpurun_out/exp11_ncu.log 2>&1
grep -v "^==" gpurun_out/exp11_ncu.csv | cut -d, -f 13- 
