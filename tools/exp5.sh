#!/bin/bash
mkdir -p gpurun_out
for v in default c2; do
  if [ $v = default ]; then unset UMT_LIB; else export UMT_LIB=$PWD/umt_b200/ab/libumtsweep_$v.so; fi
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:sweep3d_plan -s 1 -c 1 -o gpurun_out/exp5_$v -f python tools/perf_sweep.py 12 128 > gpurun_out/exp5_ncu_$v.log 2>&1
  tail -1 gpurun_out/exp5_ncu_$v.log
done
