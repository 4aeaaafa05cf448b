"""One config-B render (1M-face terrain, 6 views, 768^2) inside a cudaProfiler range:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02_b python tools/prof_b.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device('cuda', 0)
cam = wr.get_orthogonal_camera(device='cuda:0', **synth.CANONICAL_RIG)
v, f = bench.terrain_arrays(0)
m = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev))
m.set_stitched_mesh(m.v_pos, m.t_pos_idx); m.v_nrm
ctx = wr.NVDiffRastContextWrapper('cuda:0', 'cuda')
for _ in range(3):
    wr.render(ctx, m, cam, 768, 768, render_attr=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
wr.render(ctx, m, cam, 768, 768, render_attr=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
