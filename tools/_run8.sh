mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_sweep3d.py tests/test_gpu_psi_layout.py tests/test_gpu_baseline_configs.py tests/test_gpu_exchange.py tests/test_gpu_reference_cuda.py tests/test_gpu_cycle.py tests/test_gpu_reflect.py -m gpu -q --timeout=200 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -5
for S in 2 4; do
timeout 300 python bench.py --no-cpu --group-sets $S > gpurun_out/g8_bench_sets$S.json 2> gpurun_out/g8_bench_sets$S.err; tail -3 gpurun_out/g8_bench_sets$S.err
python - <<PY
import json
d=json.loads(open("gpurun_out/g8_bench_sets$S.json").read().strip().splitlines()[-1])
print("S=$S", d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["what"][-60:], d["e2e"]["one_group_set"]["ms_per_step"])
PY
done
