set -x
mkdir -p gpurun_out
timeout 400 python tools/variants.py run > gpurun_out/r02w_variants.log 2>&1
cat gpurun_out/r02w_variants.log
