set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02q_tests.log
timeout 300 python tools/bake_graph_probe.py > gpurun_out/r02q_bake_graph.log 2>&1
timeout 300 python tools/bake_phases.py 2>&1 | tail -4 > gpurun_out/r02q_phases.log
tail -4 gpurun_out/r02q_tests.log; cat gpurun_out/r02q_bake_graph.log gpurun_out/r02q_phases.log
