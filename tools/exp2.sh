#!/bin/bash
mkdir -p gpurun_out
for v in default m4 w8 w6; do
  if [ $v = default ]; then unset UMT_LIB; else export UMT_LIB=$PWD/umt_b200/ab/libumtsweep_$v.so; fi
  timeout 300 python tools/perf_sweep.py 20 128 2>&1 | tail -1
done
unset UMT_LIB
UMT_PLAN_STAGES=3 timeout 300 python tools/perf_sweep.py 20 128 2>&1 | tail -1
export UMT_LIB=$PWD/umt_b200/ab/libumtsweep_m4.so
UMT_ANGLE_BATCH=4 timeout 300 python tools/perf_sweep.py 20 128 2>&1 | tail -1
UMT_ANGLE_BATCH=2 timeout 300 python tools/perf_sweep.py 20 128 2>&1 | tail -1
