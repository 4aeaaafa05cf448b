import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device("cuda", 0)
cam = wr.get_orthogonal_camera(device="cuda:0", **synth.CANONICAL_RIG)
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
v, f = synth.icosphere(int(sys.argv[1]) if len(sys.argv) > 1 else 8, 0.5)
m = wr.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32, device=dev), t_pos_idx=torch.tensor(f, device=dev)); m.set_stitched_mesh(m.v_pos, m.t_pos_idx); m.v_nrm
for _ in range(4):
    wr.render(ctx, m, cam, 768, 768, render_attr=False)
torch.cuda.synchronize()
print("ok")
