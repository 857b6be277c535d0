#!/bin/bash
mkdir -p gpurun_out
for c in 1 2; do UMT_PLAN_CTAS_PER_SM=$c timeout 300 python tools/perf_sweep.py 20 128 2>&1 | tail -1; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sweep3d_plan -s 1 -c 1 -o gpurun_out/exp3_d12 -f python tools/perf_sweep.py 12 128 > gpurun_out/exp3_ncu.log 2>&1
tail -2 gpurun_out/exp3_ncu.log
UMT_LIB=$PWD/umt_b200/ab/libumtsweep_m4.so timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg --clock-control none -k regex:sweep3d_plan -s 1 -c 1 --csv --log-file gpurun_out/exp3_ncu_m4.csv python tools/perf_sweep.py 20 128 > gpurun_out/exp3_ncu_m4.log 2>&1
grep -v "^==" gpurun_out/exp3_ncu_m4.csv | cut -d, -f 12-
