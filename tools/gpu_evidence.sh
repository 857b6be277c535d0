#!/bin/bash
# Run under gpurun (one GPU): the measured evidence of a round -> gpurun_out/<tag>_*.
#   bench lines (default workload; -P 4 -A 4 at -d 16), clock samples, ncu launch list of the bench command, ncu captures of the
#   3-D sweep kernel (--set full, -d 20) and of the secondary kernels (sections), read here with tools/ncu_summary.py.
# usage: tools/gpu_evidence.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
Q="index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
nvidia-smi --query-gpu=$Q --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/${TAG}_bench_d20_G128.json 2> gpurun_out/${TAG}_bench.err
kill $SMI
tail -c 600 gpurun_out/${TAG}_bench_d20_G128.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --dims 16 --polar 4 --azimuthal 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_d16_P4A4.json 2>> gpurun_out/${TAG}_bench.err
tail -c 400 gpurun_out/${TAG}_bench_d16_P4A4.json
timeout 600 python bench.py --ring 3 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_d20_G128_ring3.json 2>> gpurun_out/${TAG}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_d20_G128.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
grep -v "^==" gpurun_out/${TAG}_launches_d20_G128.csv | tail -12
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep3d_plan -s 1 -c 1 -o gpurun_out/${TAG}_sweep3d_d20 -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_sweep3d.log 2>&1
timeout 600 ncu $SEC --clock-control none -k regex:phi_reduce -s 1 -c 1 -o gpurun_out/${TAG}_phi_reduce_d20 -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_phi.log 2>&1
timeout 900 ncu $SEC --clock-control none -k regex:sweep3d_plan -s 1 -c 1 -o gpurun_out/${TAG}_sweep3d_ring_d16_P4A4 -f python bench.py --dims 16 --polar 4 --azimuthal 4 --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_ring.log 2>&1
timeout 600 ncu $SEC --clock-control none -k regex:sweeprz_rec -s 2 -c 1 -o gpurun_out/${TAG}_sweeprz_rec_d40_G64 -f python tools/perf_rz_gta.py rz 40 64 > gpurun_out/${TAG}_ncu_rz.log 2>&1
timeout 600 ncu $SEC --clock-control none -k regex:gta_sweep_flow -s 3 -c 1 -o gpurun_out/${TAG}_gta_sweep_flow_d20 -f python tools/perf_rz_gta.py gta 20 16 > gpurun_out/${TAG}_ncu_gta.log 2>&1
timeout 600 ncu $SEC --clock-control none -k regex:gta_sweep_rz_flow -s 3 -c 1 -o gpurun_out/${TAG}_gta_sweep_rz_flow_d40 -f python tools/perf_rz_gta.py gtarz 40 16 > gpurun_out/${TAG}_ncu_gtarz.log 2>&1
timeout 600 ncu $SEC --clock-control none -k regex:pack_tally -s 2 -c 1 -o gpurun_out/${TAG}_pack_tally_2dom_d12 -f python tools/perf_exchange.py 12 128 > gpurun_out/${TAG}_ncu_pack.log 2>&1
timeout 300 python tools/perf_rz_gta.py rz 40 64 > gpurun_out/${TAG}_rz_perf.txt 2>&1; timeout 300 python tools/perf_rz_gta.py rz 80 64 >> gpurun_out/${TAG}_rz_perf.txt 2>&1
timeout 300 python tools/perf_rz_gta.py gta 20 16 >> gpurun_out/${TAG}_rz_perf.txt 2>&1; timeout 300 python tools/perf_rz_gta.py gtarz 40 16 >> gpurun_out/${TAG}_rz_perf.txt 2>&1
timeout 300 python tools/perf_exchange.py 12 128 >> gpurun_out/${TAG}_rz_perf.txt 2>&1
tail -6 gpurun_out/${TAG}_rz_perf.txt
ls -la gpurun_out | grep ${TAG} | awk '{print $5, $9}'
