"""Host time of an eager render() call (config B shapes): cProfile of 200 calls, GPU idle between them."""
import cProfile, io, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device("cuda", 0)
v, f = bench.terrain_arrays(0)
mesh = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev))
mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx); mesh.v_nrm
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
cam = wr.get_orthogonal_camera(device="cuda:0", **synth.CANONICAL_RIG)
for _ in range(20): wr.render(ctx, mesh, cam, 768, 768, render_attr=False)
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N):
    wr.render(ctx, mesh, cam, 768, 768, render_attr=False)
host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host time per eager render(): {1e6 * host / N:.1f} us (GPU step ~105 us)")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): wr.render(ctx, mesh, cam, 768, 768, render_attr=False)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3000])
