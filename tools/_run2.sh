set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/bake_chunks_probe.py > gpurun_out/r02n_chunks.log 2> gpurun_out/r02n_chunks.err
tail -2 gpurun_out/r02n_chunks.log; tail -3 gpurun_out/r02n_chunks.err
