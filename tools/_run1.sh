set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02h_tests.log
timeout 400 python tools/variants.py run > gpurun_out/r02h_variants.log 2>&1
for v in c8 split snap3; do
WR_B200_LIB=$PWD/worldrenderer_b200/lib/variants/lib_$v.so timeout 600 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_mv_path.py tests/test_gpu_raster_parity.py tests/test_gpu_render_graph.py -m gpu -x -q > gpurun_out/r02h_tests_$v.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02h_tests_$v.log
done
tail -3 gpurun_out/r02h_tests.log; tail -2 gpurun_out/r02h_tests_*.log; cat gpurun_out/r02h_variants.log
