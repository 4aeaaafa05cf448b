import sys, time, cProfile, pstats, io
sys.path.insert(0, '.')
import numpy as np, torch
import bench, worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device('cuda', 0)
v, f = bench.terrain_arrays(0)
mesh = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev))
mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx); mesh.v_nrm
ctx = wr.NVDiffRastContextWrapper('cuda:0', 'cuda')
cam = wr.get_orthogonal_camera(device='cuda:0', **synth.CANONICAL_RIG)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def loop(n, tag):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    t0 = time.perf_counter()
    for k in range(n):
        flush.fill_(k & 255)
        ev[k][0].record(); wr.render(ctx, mesh, cam, 768, 768, render_attr=False); ev[k][1].record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(tag, 'gpu ms/step', np.mean([a.elapsed_time(b) for a, b in ev]), 'host ms/step', 1e3 * host / n)
for _ in range(5): wr.render(ctx, mesh, cam, 768, 768, render_attr=False)
torch.cuda.synchronize()
loop(30, 'before graph')
g = wr.RenderGraph(ctx, [(mesh, cam)], 768, 768, render_attr=False)
loop(30, 'after graph ')
pr = cProfile.Profile(); pr.enable(); loop(30, 'profiled    '); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(12); print(s.getvalue()[:2500])
