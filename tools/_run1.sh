set -x
mkdir -p gpurun_out
timeout 600 python tools/unproj_rec_probe.py > gpurun_out/r02r_rec.log 2>&1
cat gpurun_out/r02r_rec.log | tail -12
