set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mv_path.py tests/test_gpu_render_parity.py tests/test_gpu_raster_parity.py tests/test_gpu_operator_golden.py -m gpu -x -q > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02x_tests.log
tail -3 gpurun_out/r02x_tests.log
timeout 300 python tools/variants.py run > gpurun_out/r02x_variants.log 2>&1
timeout 300 python tools/variants.py run >> gpurun_out/r02x_variants.log 2>&1
cat gpurun_out/r02x_variants.log
