set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02m_tests.log
timeout 600 python bench.py > gpurun_out/bench_r02m.json 2> gpurun_out/bench_r02m.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_r02m_ref.json 2> gpurun_out/bench_r02m_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02m.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02m_ncu_list.log 2>&1
tail -3 gpurun_out/r02m_tests.log
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02m.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stages_ms"], d["roofline"]["kernel"], d["roofline"]["frac"], d["pipeline"]["frac"], d["gpu_launches_per_step"])
print(d["bake"]); print(d["config_a"]); print(d["config_d"]["views_per_s"], d["bake_sharded"]["ms_per_bake"], d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"))
print(open("gpurun_out/bench_r02m_ref.json").read()[:600])
P
