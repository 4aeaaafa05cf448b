set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_r02s_n8.json 2> gpurun_out/bench_r02s_n8.err
tail -2 gpurun_out/bench_r02s_n8.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02s_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["pcie_ceiling"]["e2e_frac_of_copy_ceiling"])
b=d["bake_sharded"]; print(d["config_d"]["views_per_s"], b["ms_per_bake"], b["strong_scaling_efficiency"], b["batched"]["ms_per_bake"], b["batched"]["scaling_efficiency"], b["ms_per_bake_1_rank"], b["mask_equal_to_1_rank"], b["max_abs_err_vs_1_rank"], b["ranks_identical"])
P
