"""torchrun: config E bake on N ranks -- compute only, exchange only, single bakes, pipelined (depth 2 / 3)."""
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
lo, hi = parallel.shard_bounds(NV, world)[rank]
c = cam[lo:hi]
img = torch.rand((hi - lo, RES, RES, 3), device=dev)
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
parallel.sharded_bake(ctx, mesh, c, img, UV, **kw)
ws = parallel._p2p_workspace(UV, UV, dev, None)
pre = parallel._uv_precompute_cached(ctx, mesh, UV)
def compute():
    _, geo, att = fused_view_maps(ctx, mesh, c, img, RES, RES, 5)
    fused_unproject(ctx, pre, c, RES, RES, geo, att, aoi_cos_thresh=0.2, depth_grad_thresh=0.1, alpha=3.0, accumulate_only=True,
                    accum=ws.accum, add_to_accum=False)
res = {"compute": timed(compute), "exchange": timed(lambda: ws.reduce_finalize(ctx, mesh.texture)),
       "single": timed(lambda: parallel.sharded_bake(ctx, mesh, c, img, UV, **kw))}
ws1 = parallel._p2p_workspace(UV, UV, dev, None, 1)
xs = torch.cuda.Stream(dev, priority=-1)
def independent():
    compute()
    with torch.cuda.stream(xs):
        ws1.reduce_finalize(ctx, mesh.texture)
def both():
    independent()
    torch.cuda.current_stream().wait_stream(xs)
res["independent_streams_compute_plus_exchange"] = timed(both)
import time
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter(); compute(); t1 = time.perf_counter()
with torch.cuda.stream(xs):
    ws1.reduce_finalize(ctx, mesh.texture)
t2 = time.perf_counter(); torch.cuda.synchronize()
res["host_ms_compute_launch"] = (t1 - t0) * 1e3; res["host_ms_exchange_launch"] = (t2 - t1) * 1e3
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
import ctypes
from worldrenderer_b200 import _native
def with_side(side_fn):
    def f():
        compute()
        with torch.cuda.stream(xs):
            side_fn()
        torch.cuda.current_stream().wait_stream(xs)
    return f
def only_kernel():
    a = _native.P2PReduceArgs()
    for r in range(ws1.world):
        a.accum[r] = ws1.base_ptrs[r] + ws1._off_accum; a.out_attr[r] = ws1.base_ptrs[r] + ws1._off_attr; a.out_valid[r] = ws1.base_ptrs[r] + ws1._off_valid
    a.old_attr = _native.ptr(mesh.texture); a.world, a.rank, a.Hu, a.Wu = ws1.world, ws1.rank, UV, UV
    c_ = ctx.ctx
    c_.check(_native.lib().wr_uv_reduce_finalize_p2p(c_.handle, ctypes.byref(a), c_.stream()), "p2p")
res["side_fill_1GiB_alone"] = timed(lambda: big.fill_(1))
res["compute+side_fill"] = timed(with_side(lambda: big.fill_(1)))
res["side_barriers_alone"] = timed(lambda: (ws1.barrier(0), ws1.barrier(1)))
res["compute+side_barriers"] = timed(with_side(lambda: (ws1.barrier(0), ws1.barrier(1))))
res["side_p2p_kernel_alone"] = timed(only_kernel)
res["compute+side_p2p_kernel"] = timed(with_side(only_kernel))
for nb in (1, 2, 4):
    res[f"exchange_alone_{nb}_blocks_per_sm"] = timed(lambda: ws.reduce_finalize(ctx, mesh.texture, max_blocks=148 * nb))
    res[f"compute+side_exchange_{nb}_blocks_per_sm"] = timed(with_side(lambda: ws1.reduce_finalize(ctx, mesh.texture, max_blocks=148 * nb)))
for depth, nb in ((2, 1), (2, 2), (3, 1)):
    pipe = parallel.BakePipeline(ctx, UV, depth=depth, exchange_blocks_per_sm=nb)
    def batch():
        tk = [pipe.submit(mesh, c, img, **kw) for _ in range(8)]
        for t in tk: t.result()
    res[f"pipelined_depth{depth}_{nb}blk"] = timed(batch, reps=2) / 8
if rank == 0:
    print({k: round(x, 3) for k, x in res.items()})
dist.destroy_process_group()
