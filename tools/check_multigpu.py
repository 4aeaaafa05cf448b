"""torchrun check of the multi-GPU paths: view-sharded bake with an NCCL all-reduce vs the single-GPU bake,
and by-mesh render sharding.  Rank 0 prints PASS / FAIL.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import worldrenderer_b200 as wr  # noqa: E402
from worldrenderer_b200 import parallel, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

v, f = synth.icosphere(50, 0.5)
vt, ft = synth.cell_atlas_uv(f.shape[0])
uv = 1024
mesh = wr.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32), t_pos_idx=torch.tensor(f),
                       v_tex=torch.tensor(vt, dtype=torch.float32), t_tex_idx=torch.tensor(ft),
                       texture=torch.full((uv, uv, 3), 0.25))
mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx)
mesh.to(dev)
nv = 8
cam = wr.get_orthogonal_camera(elevation_deg=[15.0] * nv, distance=[1.0] * nv, left=-0.55, right=0.55, bottom=-0.55,
                               top=0.55, azimuth_deg=list(np.linspace(0, 360, nv + 1)[:-1]), device=str(dev))
images = torch.from_numpy(synth.view_images(nv, 512, 512, seed=1)).to(dev)
ctx = wr.NVDiffRastContextWrapper(str(dev), "cuda")

lo, hi = parallel.my_shard(nv)
kw = dict(aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1, uv_exp_blend_alpha=3.0)
atlas, any_ = parallel.sharded_bake(ctx, mesh, cam[lo:hi], images[lo:hi], uv, **kw)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    atlas, any_ = parallel.sharded_bake(ctx, mesh, cam[lo:hi], images[lo:hi], uv, **kw)
torch.cuda.synchronize()
dist.barrier()
ms = (time.perf_counter() - t0) / 5 * 1e3

# every rank also computes the full single-GPU bake and compares
proj = wr.CameraProjection(None, None, str(dev), "cuda")
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    full = proj(images, mesh, cam, uv_size=uv, poisson_blending=False, uv_padding=False, iou_rejection_threshold=None,
                depth_grad_dilation=5, return_dict=True, **kw)
ok_mask = torch.equal(any_, full.uv_proj_mask)
err = float((atlas - full.uv_proj).abs().max())
ok = ok_mask and err < 1e-5
# all ranks must hold the same atlas bit for bit
gathered = [torch.empty_like(atlas) for _ in range(world)]
dist.all_gather(gathered, atlas)
same = all(torch.equal(g, gathered[0]) for g in gathered)
flag = torch.tensor([int(ok and same)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)

# by-mesh render sharding: rank r renders meshes [lo, hi) of 4 copies
first, outs = parallel.render_mesh_shard(ctx, [mesh] * 4, cam, 256, 256, render_attr=False)
n_local = torch.tensor([len(outs)], device=dev)
dist.all_reduce(n_local)
if rank == 0:
    print(f"world={world} views={nv} sharded_bake_ms={ms:.3f} mask_equal={ok_mask} max_abs_err={err:.2e} "
          f"ranks_identical={same} meshes_rendered={int(n_local)} covered_texels={int(any_.sum())}")
    print("PASS" if int(flag) == 1 and int(n_local) == 4 else "FAIL")
dist.destroy_process_group()
