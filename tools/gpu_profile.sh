#!/bin/bash
# Run under gpurun: bench line, ncu launch list, one ncu --set full capture of the sweep kernel.
# usage: tools/gpu_profile.sh <tag> [bench args...]
TAG=${1:-r1}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
kill $SMI
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu "$@" > gpurun_out/${TAG}_launches.log 2>&1
grep -v "^==" gpurun_out/${TAG}_launches.csv | tail -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep3d -s 1 -c 1 -o gpurun_out/${TAG}_sweep3d -f python bench.py --dims 6 --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out
