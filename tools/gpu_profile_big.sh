#!/bin/bash
# ncu capture of the sweep kernel at a large domain (default -d 16,16,16 -G 128): dram bytes, throughput, stalls.
TAG=${1:-r1big}; D=${2:-16}
mkdir -p gpurun_out
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
  --clock-control none --import-source on -k regex:sweep3d_plan -s 1 -c 1 -o gpurun_out/${TAG}_sweep3d_d${D} -f \
  python bench.py --dims $D --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_d${D}.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_d${D}.log
ls -la gpurun_out | tail -5
