"""torchrun: per-rank time of the view passes + unprojection of config E, contiguous vs interleaved view shards."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import worldrenderer_b200 as wr
from worldrenderer_b200 import parallel, synth
from worldrenderer_b200.uv import fused_view_maps, fused_unproject
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
NV, RES, UV = 32, 2048, 4096
v, f = synth.terrain(2500, 1000, 0)
v = v / np.abs(v).max() * 0.5
v = np.ascontiguousarray(np.stack([v[:, 0], -v[:, 2], v[:, 1]], -1), np.float32)
vt = synth.terrain_uv(2500, 1000).astype(np.float32)
mesh = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev), v_tex=torch.from_numpy(vt).to(dev),
                       t_tex_idx=torch.from_numpy(f).to(dev), texture=torch.zeros((UV, UV, 3), device=dev))
mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx); mesh.v_nrm
cam = wr.get_orthogonal_camera(elevation_deg=[20.0] * NV, distance=[1.0] * NV, left=-0.55, right=0.55, bottom=-0.55, top=0.55,
                               azimuth_deg=list(np.linspace(0, 360, NV + 1)[:-1]), device=str(dev))
ctx = wr.NVDiffRastContextWrapper(str(dev), "cuda")
pre = wr.uv_precompute(ctx, mesh, UV, UV)
accum = torch.empty((UV, UV, 5), device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for name, c in (("contiguous", cam[slice(*parallel.shard_bounds(NV, world)[rank])]), ("interleaved", cam[rank::world])):
    n = c.mvp_mtx.shape[0]
    img = torch.rand((n, RES, RES, 3), device=dev)
    def run():
        e = [ev() for _ in range(3)]
        e[0].record()
        _, geo, att = fused_view_maps(ctx, mesh, c, img, RES, RES, 5)
        e[1].record()
        fused_unproject(ctx, pre, c, RES, RES, geo, att, aoi_cos_thresh=0.2, depth_grad_thresh=0.1, alpha=3.0, accumulate_only=True, accum=accum, add_to_accum=False)
        e[2].record()
        return e
    for _ in range(3): run()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(6):
        e = run(); torch.cuda.synchronize(); ts.append([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])])
    t = torch.tensor(np.mean(ts, 0), dtype=torch.float64, device=dev)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        a = torch.stack(allt).cpu().numpy()
        print(name, "view maps per rank", np.round(a[:, 0], 3), "unproject per rank", np.round(a[:, 1], 3), "max total %.3f mean total %.3f" % (a.sum(1).max(), a.sum(1).mean()))
dist.destroy_process_group()
