"""Config B (1 mesh x 6 views) through RenderGraph with the views split over 1..3 concurrent lanes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device('cuda', 0)
cam = wr.get_orthogonal_camera(device='cuda:0', **synth.CANONICAL_RIG)
v, f = bench.terrain_arrays(0)
m = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev))
m.set_stitched_mesh(m.v_pos, m.t_pos_idx); m.v_nrm
ctx = wr.NVDiffRastContextWrapper('cuda:0', 'cuda')
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ref = wr.render(ctx, m, cam, 768, 768, render_attr=False)
for vl, st in ((1, False), (2, False), (2, True), (3, False), (3, True), (6, False), (6, True)):
    g = wr.RenderGraph(ctx, [(m, cam)], 768, 768, view_lanes=vl, stagger=st, render_attr=False)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    K = 40
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.fill_(k & 255); ev[k][0].record(); g.replay(); ev[k][1].record()
    torch.cuda.synchronize()
    ms = np.mean([a.elapsed_time(b) for a, b in ev])
    o = g.replay()[0]; torch.cuda.synchronize()
    same = all(torch.equal(getattr(o, n), getattr(ref, n)) for n in ("mask", "pos", "depth", "normal"))
    print(f'view_lanes {vl} stagger {st}: {ms*1e3:.1f} us per step, {6/ms*1e3:.0f} views/s, identical to eager: {same}')
    del g
