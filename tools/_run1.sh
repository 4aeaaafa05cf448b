set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02j_tests.log
timeout 300 python tools/bake_graph_probe.py > gpurun_out/r02j_bake_graph.log 2>&1
timeout 300 python tools/prof_render.py --steps 2 --mesh sphere --bake > gpurun_out/r02j_bake_stages.log 2>&1
timeout 300 python tools/bake_phases.py > gpurun_out/r02j_bake_phases.log 2>&1
tail -3 gpurun_out/r02j_tests.log; cat gpurun_out/r02j_bake_graph.log; tail -4 gpurun_out/r02j_bake_stages.log; tail -8 gpurun_out/r02j_bake_phases.log
