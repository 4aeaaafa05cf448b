set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02z_tests.log
tail -3 gpurun_out/r02z_tests.log
timeout 300 python tools/variants.py run > gpurun_out/r02z_variants.log 2>&1
timeout 300 python tools/variants.py run >> gpurun_out/r02z_variants.log 2>&1
cat gpurun_out/r02z_variants.log
