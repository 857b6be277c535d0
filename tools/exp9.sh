#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/perf_sweep.py 6 128 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/exp9_pytest.log 2>&1; tail -3 gpurun_out/exp9_pytest.log
for v in default c1w8 c2w6; do
  if [ $v = default ]; then unset UMT_LIB; else export UMT_LIB=$PWD/umt_b200/ab/libumtsweep_$v.so; fi
  timeout 300 python tools/perf_sweep.py 20 128 2>&1 | tail -1
done
