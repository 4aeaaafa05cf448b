set -x
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_bake_parity.py -m gpu -x -q > gpurun_out/r02ad_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02ad_tests.log
tail -15 gpurun_out/r02ad_tests.log
