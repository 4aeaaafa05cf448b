set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02g_tests.log
timeout 300 python tools/variants.py run > gpurun_out/r02g_variants.log 2>&1
WR_B200_LIB=$PWD/worldrenderer_b200/lib/variants/lib_c8.so timeout 600 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_mv_path.py tests/test_gpu_raster_parity.py tests/test_gpu_render_graph.py -m gpu -x -q > gpurun_out/r02g_tests_c8.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02g_tests_c8.log
timeout 300 python tools/exp_view_lanes.py > gpurun_out/r02g_view_lanes.log 2>&1
timeout 300 python tools/bake_graph_probe.py > gpurun_out/r02g_bake_graph.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
tail -3 gpurun_out/r02g_tests.log; tail -3 gpurun_out/r02g_tests_c8.log; cat gpurun_out/r02g_variants.log gpurun_out/r02g_view_lanes.log gpurun_out/r02g_bake_graph.log
