"""torchrun: where does the N-rank bake differ from the 1-rank bake of the same views?"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import worldrenderer_b200 as wr
from worldrenderer_b200 import parallel, synth
from worldrenderer_b200.uv import fused_view_maps, fused_unproject, uv_finalize
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
g = torch.Generator(device="cpu").manual_seed(1)
img_all = torch.rand((NV, 64, 64, 3), generator=g)
img_all = torch.nn.functional.interpolate(img_all.permute(0, 3, 1, 2), size=(RES, RES), mode="bilinear").permute(0, 2, 3, 1).contiguous().to(dev)
ctx = wr.NVDiffRastContextWrapper(str(dev), "cuda")
mine = parallel.shard_slice(NV, rank, world, interleave=True)
kw = dict(aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1, uv_exp_blend_alpha=3.0)
for mode in ("auto", "nccl"):
    atlas, any_ = parallel.sharded_bake(ctx, mesh, cam[mine], img_all[mine].contiguous(), UV, exchange=mode, **kw)
    atlas, any_ = atlas.clone(), any_.clone()
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        pre = parallel._uv_precompute_cached(ctx, mesh, UV)
        _, geo, att = fused_view_maps(ctx, mesh, cam, img_all, RES, RES, 5)
        a1, m1, _, _, _ = fused_unproject(ctx, pre, cam, RES, RES, geo, att, aoi_cos_thresh=0.2, depth_grad_thresh=0.1, alpha=3.0)
        d = (a1 - atlas).abs().max(-1).values
        print(mode, "mask equal", bool(torch.equal(m1, any_)), "max abs diff", float(d.max()), "texels > 1e-5:", int((d > 1e-5).sum()), "> 1e-3:", int((d > 1e-3).sum()))
        idx = torch.nonzero(d > 1e-3)[:5].tolist()
        for y, x in idx:
            print("   ", (y, x), a1[y, x].tolist(), atlas[y, x].tolist())
    dist.barrier()
dist.destroy_process_group()
