#!/bin/bash
mkdir -p gpurun_out
export UMT_LIB=$PWD/umt_b200/ab/libumtsweep_c2.so
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sweep3d_plan -s 1 -c 1 -o gpurun_out/exp7_c2 -f python tools/perf_sweep.py 16 128 > gpurun_out/exp7_ncu_c2.log 2>&1
tail -1 gpurun_out/exp7_ncu_c2.log
