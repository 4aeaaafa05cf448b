"""GPU tests of the multi-GPU bake's exchange kernels on ONE device (csrc/bake.cu k_uv_reduce_finalize_p2p / _mc).

wr_uv_reduce_finalize_p2p only takes N pointers to the ranks' accumulators / atlases / masks: here they are N local
buffers standing in for the peer mappings, and the entry point is called once per emulated rank (rank r sums and
finalises its 1/N of the texels and stores them into every "rank's" atlas).  After all N calls every atlas must be
complete, identical, and equal to wr_uv_finalize of the summed accumulators: mask bit-exact, colours 1e-5.

The multicast variant (multimem.ld_reduce / multimem.st) needs an NVSwitch multicast window; a single-rank NCCL
group with torch symmetric memory provides one on NVSwitch boxes -- the kernel then runs with world = 1 (sum over one
rank).  Skipped where no window can be made."""
import ctypes
import os

import numpy as np
import pytest
import torch

import worldrenderer_b200 as wr
from worldrenderer_b200 import _native
from worldrenderer_b200.uv import uv_finalize

pytestmark = pytest.mark.gpu


def _accumulators(world, Hu, Wu, dev, seed):
    """Sparse, view-like accumulators: every rank sees a different part of the atlas; valid counts are small ints."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    acc = []
    for r in range(world):
        seen = torch.rand((Hu, Wu), generator=g) < 0.45
        w = torch.rand((Hu, Wu), generator=g) ** 3 * seen
        rgb = torch.rand((Hu, Wu, 3), generator=g)
        nvalid = torch.randint(0, 4, (Hu, Wu), generator=g).float() * seen
        w = torch.where(nvalid > 0, w, torch.zeros_like(w))           # a valid view may still have weight 0 ...
        w[:: 7, :: 5] = 0.0                                            # ... (aoi threshold below 0, uv.py:335-340)
        a = torch.cat([rgb * w[..., None], w[..., None], nvalid[..., None]], -1).contiguous()
        acc.append(a.to(dev))
    return acc


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("Hu,Wu,with_old", [(256, 256, True), (96, 172, False), (64, 1030, True)])
def test_p2p_exchange_emulated_ranks(wr_ctx, world, Hu, Wu, with_old):
    dev = wr_ctx.device
    assert (Hu * Wu) % 4 == 0
    acc = _accumulators(world, Hu, Wu, dev, seed=world * 1000 + Hu)
    old = torch.rand((Hu, Wu, 3), device=dev) if with_old else None
    attr = [torch.full((Hu, Wu, 3), -7.0, device=dev) for _ in range(world)]
    valid = [torch.full((Hu, Wu), 9, dtype=torch.uint8, device=dev) for _ in range(world)]
    c = wr_ctx.ctx
    for rank in range(world):
        a = _native.P2PReduceArgs()
        for r in range(world):
            a.accum[r], a.out_attr[r], a.out_valid[r] = _native.ptr(acc[r]), _native.ptr(attr[r]), _native.ptr(valid[r])
        a.old_attr = _native.ptr(old) if old is not None else None
        a.world, a.rank, a.Hu, a.Wu = world, rank, Hu, Wu
        a.max_blocks = 0 if rank % 2 == 0 else 3   # a bounded grid (BakePipeline) must cover the same texels
        c.check(_native.lib().wr_uv_reduce_finalize_p2p(c.handle, ctypes.byref(a), c.stream()), "wr_uv_reduce_finalize_p2p")
    torch.cuda.synchronize()
    total = torch.stack(acc).sum(0)
    want_attr, want_any = uv_finalize(wr_ctx, total.contiguous(), old)
    for r in range(world):
        assert torch.equal(valid[r], valid[0]) and torch.equal(attr[r], attr[0]), "ranks hold different atlases"
    np.testing.assert_array_equal(valid[0].cpu().numpy().astype(bool), want_any.cpu().numpy())
    np.testing.assert_allclose(attr[0].cpu().numpy(), want_attr.cpu().numpy(), rtol=1e-5, atol=1e-6)
    # and against float64 arithmetic on the host (sum order of the kernel is rank+1, rank+2, ...: 1e-5 covers it)
    t64 = torch.stack([a.double().cpu() for a in acc]).sum(0)
    any64 = t64[..., 4] > 0.5
    ref = torch.where(any64[..., None], t64[..., :3] / t64[..., 3:4].clamp(min=1e-5),
                      old.double().cpu() if old is not None else torch.zeros(Hu, Wu, 3, dtype=torch.float64))
    np.testing.assert_array_equal(valid[0].cpu().numpy().astype(bool), any64.numpy())
    got = attr[0].double().cpu()
    stable = (t64[..., 3] > 1e-4) | ~any64        # tiny weight sums amplify the last bit of the sum
    np.testing.assert_allclose(got[stable].numpy(), ref[stable].numpy(), rtol=1e-5, atol=1e-6)


def test_p2p_exchange_rejects_bad_arguments(wr_ctx):
    c = wr_ctx.ctx
    a = _native.P2PReduceArgs()
    a.world, a.rank, a.Hu, a.Wu = 2, 0, 3, 3   # 9 texels: not a multiple of 4
    buf = torch.zeros(64, device=wr_ctx.device)
    for r in range(2):
        a.accum[r] = a.out_attr[r] = a.out_valid[r] = _native.ptr(buf)
    assert _native.lib().wr_uv_reduce_finalize_p2p(c.handle, ctypes.byref(a), c.stream()) != 0
    a.Hu, a.Wu, a.rank = 4, 4, 5               # rank outside the world
    assert _native.lib().wr_uv_reduce_finalize_p2p(c.handle, ctypes.byref(a), c.stream()) != 0


@pytest.mark.parametrize("world", [2, 8])
def test_p2p_exchange_in_chunks(wr_ctx, world):
    """The atlas exchanged chunk by chunk (tex_lo / tex_hi, what sharded_bake(chunks=...) issues): chunks on 1024-texel
    boundaries, the last one ragged, every rank owning 1/N of each chunk -- same atlas as one exchange."""
    dev = wr_ctx.device
    Hu, Wu = 120, 172   # 20640 texels: chunks of 7168 texels -> 7168, 7168, 6304
    acc = _accumulators(world, Hu, Wu, dev, seed=77)
    old = torch.rand((Hu, Wu, 3), device=dev)
    attr = [torch.full((Hu, Wu, 3), -7.0, device=dev) for _ in range(world)]
    valid = [torch.full((Hu, Wu), 9, dtype=torch.uint8, device=dev) for _ in range(world)]
    c = wr_ctx.ctx
    ntex, step = Hu * Wu, 7168
    for lo in range(0, ntex, step):
        for rank in range(world):
            a = _native.P2PReduceArgs()
            for r in range(world):
                a.accum[r], a.out_attr[r], a.out_valid[r] = _native.ptr(acc[r]), _native.ptr(attr[r]), _native.ptr(valid[r])
            a.old_attr = _native.ptr(old)
            a.world, a.rank, a.Hu, a.Wu = world, rank, Hu, Wu
            a.tex_lo, a.tex_hi, a.max_blocks = lo, min(lo + step, ntex), 2
            c.check(_native.lib().wr_uv_reduce_finalize_p2p(c.handle, ctypes.byref(a), c.stream()), "wr_uv_reduce_finalize_p2p")
    torch.cuda.synchronize()
    want_attr, want_any = uv_finalize(wr_ctx, torch.stack(acc).sum(0).contiguous(), old)
    for r in range(world):
        np.testing.assert_array_equal(valid[r].cpu().numpy().astype(bool), want_any.cpu().numpy())
        np.testing.assert_allclose(attr[r].cpu().numpy(), want_attr.cpu().numpy(), rtol=1e-5, atol=1e-6)
    a.tex_lo = 512   # not on a block boundary
    assert _native.lib().wr_uv_reduce_finalize_p2p(c.handle, ctypes.byref(a), c.stream()) != 0


def test_multicast_exchange_single_rank_window(wr_ctx):
    """k_uv_reduce_finalize_mc through a real multicast window (world = 1)."""
    import torch.distributed as dist
    dev = wr_ctx.device
    created = False
    try:
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29731")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
            created = True
        import torch.distributed._symmetric_memory as symm_mem
        Hu = Wu = 128
        T = Hu * Wu
        buf = symm_mem.empty(33 * T, dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
    except Exception as e:  # no symmetric memory / no NVLS on this box
        if created:
            dist.destroy_process_group()
        pytest.skip(f"no multicast window available: {type(e).__name__}: {e}")
    try:
        if mc == 0:
            pytest.skip("symmetric memory has no multicast pointer on this box (NVLS unavailable)")
        accum = buf[0:20 * T].view(torch.float32).view(Hu, Wu, 5)
        attr = buf[20 * T:32 * T].view(torch.float32).view(Hu, Wu, 3)
        valid = buf[32 * T:33 * T].view(Hu, Wu)
        accum.copy_(_accumulators(1, Hu, Wu, dev, seed=5)[0])
        attr.fill_(-3.0); valid.fill_(7)
        old = torch.rand((Hu, Wu, 3), device=dev)
        a = _native.P2PReduceArgs()
        base = int(hdl.buffer_ptrs[0])
        a.accum[0], a.out_attr[0], a.out_valid[0] = base, base + 20 * T, base + 32 * T
        a.mc_accum, a.mc_attr, a.mc_valid = mc, mc + 20 * T, mc + 32 * T
        a.old_attr = _native.ptr(old)
        a.world, a.rank, a.Hu, a.Wu = 1, 0, Hu, Wu
        c = wr_ctx.ctx
        want_attr, want_any = uv_finalize(wr_ctx, accum.clone(), old)
        for max_blocks in (0, 37):   # the stand-alone kernel, then the light bounded one BakePipeline launches
            attr.fill_(-3.0); valid.fill_(7)
            a.max_blocks = max_blocks
            torch.cuda.synchronize()
            c.check(_native.lib().wr_uv_reduce_finalize_p2p(c.handle, ctypes.byref(a), c.stream()), "wr_uv_reduce_finalize_p2p(mc)")
            torch.cuda.synchronize()
            np.testing.assert_array_equal(valid.cpu().numpy().astype(bool), want_any.cpu().numpy())
            np.testing.assert_allclose(attr.cpu().numpy(), want_attr.cpu().numpy(), rtol=1e-5, atol=1e-6)
    finally:
        if created:
            dist.destroy_process_group()
