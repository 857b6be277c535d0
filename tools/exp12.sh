#!/bin/bash
mkdir -p gpurun_out
for G in 128 64 32 16; do
for nh in 1 2; do
UMT_PLAN_NH=$nh timeout 300 python tools/perf_sweep.py 20 $G 2>&1 | tail -1
done; done
