set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02g_tests.log
timeout 300 python tools/exp_view_lanes.py > gpurun_out/r02g_view_lanes.log 2>&1
timeout 300 python tools/bake_graph_probe.py > gpurun_out/r02g_bake_graph.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
tail -3 gpurun_out/r02g_tests.log; cat gpurun_out/r02g_view_lanes.log gpurun_out/r02g_bake_graph.log
