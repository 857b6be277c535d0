mkdir -p gpurun_out
timeout 300 python -u -m pytest tests/test_gpu_sweep3d.py -m gpu -q --timeout=200 --timeout-method=thread -p no:cacheprovider -k "sets or control_sweep" > gpurun_out/g7_pytest.log 2>&1; tail -12 gpurun_out/g7_pytest.log
for S in 2 4; do
timeout 300 python bench.py --no-cpu --group-sets $S > gpurun_out/g7_bench_sets$S.json 2> gpurun_out/g7_bench_sets$S.err; tail -3 gpurun_out/g7_bench_sets$S.err
python - <<PY
import json
d=json.loads(open("gpurun_out/g7_bench_sets$S.json").read().strip().splitlines()[-1])
print("S=$S", d["ms_per_step"], d["e2e"])
PY
done
