set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02o_tests.log
timeout 300 python tools/prof_render.py --steps 3 --per-view --quick > gpurun_out/r02o_stages.log 2>&1
timeout 600 python bench.py --no-cpu > gpurun_out/bench_r02o.json 2> gpurun_out/bench_r02o.err
tail -3 gpurun_out/r02o_tests.log; cat gpurun_out/r02o_stages.log
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02o.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stages_ms"], d["roofline"]["kernel"], d["roofline"]["frac"], d["pipeline"]["frac"])
print(d["config_a"]["views_per_s"], d["config_d"]["views_per_s"], d["bake_sharded"]["ms_per_bake"], d["bake"]["ms_per_uv_bake_end_to_end"])
P
