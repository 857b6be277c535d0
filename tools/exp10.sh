#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/tune_batch.py 20 128 > gpurun_out/exp10_tune.log 2>&1; cat gpurun_out/exp10_tune.log
for K in 2 8; do
UMT_ANGLE_BATCH=$K timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sweep3d_plan -s 1 -c 1 --csv --log-file gpurun_out/exp10_ncu_K$K.csv python tools/perf_sweep.py 20 128 > gpurun_out/exp10_ncu_K$K.log 2>&1
grep -v "^==" gpurun_out/exp10_ncu_K$K.csv | cut -d, -f 13- 
done
