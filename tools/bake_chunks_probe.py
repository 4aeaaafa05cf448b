"""torchrun: one config-E bake on N ranks by the number of atlas chunks (exchange of chunk k under the unprojection of k+1)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import worldrenderer_b200 as wr
from worldrenderer_b200 import parallel, synth
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
mine = parallel.shard_slice(NV, rank, world, interleave=True)
c = cam[mine]
img = torch.rand((c.mvp_mtx.shape[0], RES, RES, 3), device=dev)
ctx = wr.NVDiffRastContextWrapper(str(dev), "cuda")
kw = dict(aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1, uv_exp_blend_alpha=3.0)
def timed(fn, reps=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
res = {}
for shape in ("falling", "equal"):
    for ch in (1, 2, 3, 4, 6, 8, 12):
        if ch == 1 and shape == "equal":
            continue
        res[f"{shape}_{ch}"] = round(timed(lambda: parallel.sharded_bake(ctx, mesh, c, img, UV, chunks=ch, chunk_shape=shape, **kw)), 3)
for ch in (4,):   # repeat: run-to-run spread on this box
    res[f"falling_{ch}_again"] = round(timed(lambda: parallel.sharded_bake(ctx, mesh, c, img, UV, chunks=ch, **kw)), 3)
if rank == 0:
    print(res)
dist.destroy_process_group()
