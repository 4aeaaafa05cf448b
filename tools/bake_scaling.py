"""Strong-scaling measurement of the view-sharded bake (config E shape, scaled to fit a short run):
1M-face terrain with a planar atlas, 32 views of 1024^2 on a ring, 2048^2 atlas.  The 32 views are split
over the ranks; every rank ends up with the full atlas.  Run with torchrun at N = 1, 2, 4, 8.
"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, worldrenderer_b200 as wr
from worldrenderer_b200 import parallel, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
FULL = os.environ.get("WR_CONFIG_E", "0") == "1"  # the full config E: 5M faces, 32 x 2048^2 views, 4096^2 atlas
NV, RES, UV = (32, 2048, 4096) if FULL else (32, 1024, 2048)
NX, NY = (2500, 1000) if FULL else bench.TERRAIN
v, f = synth.terrain(NX, NY, 0)
v = v / np.abs(v).max() * 0.5
v = np.ascontiguousarray(np.stack([v[:, 0], -v[:, 2], v[:, 1]], -1), np.float32)
f = np.ascontiguousarray(f, np.int64)
vt = synth.terrain_uv(NX, NY).astype(np.float32)
mesh = wr.TexturedMesh(v_pos=torch.from_numpy(v), t_pos_idx=torch.from_numpy(f), v_tex=torch.from_numpy(vt),
                       t_tex_idx=torch.from_numpy(f).clone(), texture=torch.zeros((UV, UV, 3)))
mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx)
mesh.to(dev); mesh.v_nrm
cam = wr.get_orthogonal_camera(elevation_deg=[20.0] * NV, distance=[1.0] * NV, left=-0.55, right=0.55, bottom=-0.55,
                               top=0.55, azimuth_deg=list(np.linspace(0, 360, NV + 1)[:-1]), device=str(dev))
lo, hi = parallel.shard_bounds(NV, world)[rank]
images = torch.from_numpy(synth.view_images(hi - lo, RES, RES, seed=1 + rank)).to(dev)
ctx = wr.NVDiffRastContextWrapper(str(dev), "cuda")
kw = dict(aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1, uv_exp_blend_alpha=3.0)
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
for mode in (["auto", "nccl"] if world > 1 else ["auto"]):
    for _ in range(3):
        atlas, any_ = parallel.sharded_bake(ctx, mesh, cam[lo:hi], images, UV, exchange=mode, **kw)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        atlas, any_ = parallel.sharded_bake(ctx, mesh, cam[lo:hi], images, UV, exchange=mode, **kw)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"bake_scaling world={world} exchange={mode} views={NV}x{RES}^2 atlas={UV}^2 faces={f.shape[0]}: "
              f"{float(t):.3f} ms per bake (covered texels {int(any_.sum())})")
if world > 1:
    dist.destroy_process_group()
