"""Config D (8 meshes x 6 views per step) through RenderGraph with 1..4 concurrent lanes: ms per step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device('cuda', 0)
cam = wr.get_orthogonal_camera(device='cuda:0', **synth.CANONICAL_RIG)
meshes = []
for j in range(8):
    v, f = bench.terrain_arrays(j)
    m = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev))
    m.set_stitched_mesh(m.v_pos, m.t_pos_idx); m.v_nrm
    meshes.append(m)
ctx = wr.NVDiffRastContextWrapper('cuda:0', 'cuda')
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ref = None
for lanes in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    g = wr.RenderGraph(ctx, [(m, cam) for m in meshes], 768, 768, lanes=lanes, render_attr=False)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    K = 20
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.fill_(k & 255); ev[k][0].record(); g.replay(); ev[k][1].record()
    torch.cuda.synchronize()
    ms = np.mean([a.elapsed_time(b) for a, b in ev])
    outs = g.replay(); torch.cuda.synchronize()
    sig = [(o.mask.sum().item(), float(o.pos.double().sum()), float(o.depth.double().sum()), float(o.normal.double().sum())) for o in outs]
    if ref is None: ref = sig
    print(f'lanes {lanes}: {ms*1e3:.1f} us per step, {ms*1e3/8:.1f} us per mesh, {48/ms*1e3:.0f} views/s, same outputs as lanes=1: {sig == ref}')
    del g
