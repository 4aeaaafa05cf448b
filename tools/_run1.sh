set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02ab_tests.log
tail -3 gpurun_out/r02ab_tests.log
timeout 400 python bench.py > gpurun_out/bench_r02ab.json 2> gpurun_out/bench_r02ab.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02ab.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages_ms"], d["bake"], d["config_a"]["views_per_s"], d["config_d"]["views_per_s"], d["bake_sharded"]["ms_per_bake"])
P
