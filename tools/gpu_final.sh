#!/bin/bash
# Run under gpurun (one GPU): the final checks of a round -> gpurun_out/<tag>_*: full GPU test log, smoke, bench lines of the final
# build, launch list, a --set full capture of the sweep kernel at a small size (source page), compute-sanitizer memcheck.
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 700 python -u -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
Q="index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
nvidia-smi --query-gpu=$Q --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/${TAG}_bench_d20_G128.json 2> gpurun_out/${TAG}_bench.err
kill $SMI
tail -c 300 gpurun_out/${TAG}_bench_d20_G128.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --dims 16 --polar 4 --azimuthal 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_d16_P4A4.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_d20_G128.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
if [ -z "$QUICK" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweep3d_plan -s 1 -c 1 -o gpurun_out/${TAG}_sweep3d_d8_full -f python bench.py --dims 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_full.log
MEMTESTS="tests/test_gpu_psi_layout.py tests/test_gpu_sweep3d.py tests/test_gpu_sweeprz.py tests/test_gpu_gta.py tests/test_gpu_exchange.py tests/test_gpu_watchdog.py"
else   # QUICK=1: what changed since the last full run only
MEMTESTS="tests/test_gpu_gta.py tests/test_gpu_exchange.py tests/test_gpu_sweeprz.py::test_control_sweep_sets_rz_group_sets"
fi
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $MEMTESTS -m gpu -q -x --timeout=450 --timeout-method=thread -p no:cacheprovider -k "not strongly and not fullsize" > gpurun_out/${TAG}_memcheck.log 2>&1; tail -4 gpurun_out/${TAG}_memcheck.log
timeout 100 python -u tools/perf_rz_gta.py gta 20 > gpurun_out/${TAG}_gta_solve_d20.log 2>&1; tail -1 gpurun_out/${TAG}_gta_solve_d20.log
ls -la gpurun_out | grep ${TAG} | awk '{print $5, $9}'
