"""CUDA-graph replay of render() calls (no reference counterpart: host-side launch path).

One render() is five short kernels (~110 us of GPU time for a 1M-face mesh and six 768^2 views) issued from
Python through ctypes: with several processes sharing the host (one per GPU) the launch path can fall behind
the GPU and the gaps land inside the step.  `RenderGraph` captures the render() calls of a fixed job list --
config D is eight meshes per GPU per step -- into ONE CUDA graph (the kernels keep their programmatic
dependent launch edges), so that a step is a single `cudaGraphLaunch`.

`lanes > 1` captures the jobs round-robin on that many streams, each with a raster context (scratch) of its own,
so the graph holds independent chains: the raster set-up of one mesh is bound by instruction issue, the shading
pass of another by the latency of its gathers and by its stores, and the two overlap when they run side by side.

`view_lanes > 1` does the same inside ONE job: its views are split into that many groups, each rendered by its own
chain into a slice of the job's output tensors (the vertex pass runs once per group).  With `stagger=True` group
k + 1 starts when the raster passes of group k are done (an event the library records between the raster passes and
the shading pass), so that the chains run out of phase: the set-up pass of one group next to the shading pass of the
previous one instead of two set-up passes competing for the integer pipe.

The output tensors are static: every replay overwrites them.  Vertex positions, faces and cameras are read
from the tensors the job list held at capture time -- update those in place (copy_) to render new data of the
same shape.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .camera import Camera
from .mesh import TexturedMesh
from .render import NVDiffRastContextWrapper, RenderOutput, render


class RenderGraph:
    def __init__(self, ctx: NVDiffRastContextWrapper, jobs: Sequence[Tuple[TexturedMesh, Camera]], height: int,
                 width: int, warmup: int = 2, lanes: int = 1, view_lanes: int = 1, stagger: bool = False,
                 **render_kwargs):
        if not jobs:
            raise ValueError("RenderGraph needs at least one (mesh, camera) job")
        self.jobs, self.height, self.width = list(jobs), int(height), int(width)
        self.kwargs = dict(render_kwargs)
        self.view_lanes = max(1, int(view_lanes))
        self.stagger = bool(stagger) and self.view_lanes > 1
        self.lanes = max(1, min(int(lanes), len(self.jobs))) if self.view_lanes == 1 else self.view_lanes
        # contexts of its own: the captured kernels hold pointers into a context's scratch, which an eager call
        # of a larger shape on a shared context would reallocate; one per lane, because concurrent chains cannot
        # share the packed visibility buffer
        self.ctxs = [NVDiffRastContextWrapper(str(ctx.device), ctx.context_type) for _ in range(self.lanes)]
        self.ctx = self.ctxs[0]
        dev = self.ctx.device
        for mesh, _ in self.jobs:
            if self.kwargs.get("render_normal", True):
                mesh.v_nrm  # lazily computed once, outside the capture
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self._static: List[Optional[RenderOutput]] = [None] * len(self.jobs)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):   # scratch growth (cudaMalloc) and index caches happen here
                outs = self._run_serial()
            if self.view_lanes > 1:           # the whole-batch outputs of a job: the groups render into their slices
                self._static = outs
                for _ in range(max(1, warmup)):
                    for j in range(len(self.jobs)):
                        for g in range(self.view_lanes):
                            self._render_group(g, j)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._lane_streams = [torch.cuda.Stream(dev) for _ in range(self.lanes - 1)]
        self._raster_events = []
        if self.stagger:
            for _ in range(self.view_lanes):
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))   # torch creates the CUDA event on the first record
                self._raster_events.append(ev)
            torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs: List[RenderOutput] = self._run_lanes()

    def _render(self, lane: int, job: int) -> RenderOutput:
        m, c = self.jobs[job]
        return render(self.ctxs[lane], m, c, self.height, self.width, **self.kwargs)

    def _groups(self, job: int):
        n = self.jobs[job][1].mvp_mtx.shape[0]
        k = min(self.view_lanes, max(n, 1))
        return [slice(n * g // k, n * (g + 1) // k) for g in range(k)]

    def _render_group(self, group: int, job: int, raster_done_event=None) -> None:
        """Views `group` of job `job` into the matching slices of the job's static outputs."""
        groups = self._groups(job)
        if group >= len(groups) or groups[group].start == groups[group].stop:
            return
        sl = groups[group]
        m, c = self.jobs[job]
        full = self._static[job]
        bufs = {name: getattr(full, name)[sl] for name in ("mask", "pos", "depth", "normal", "attr", "tangent")
                if getattr(full, name, None) is not None}
        out = render(self.ctxs[group], m, c[sl], self.height, self.width, _out_buffers=bufs,
                     _raster_done_event=raster_done_event, **self.kwargs)
        for name, buf in bufs.items():   # a normaliser / background that allocates its own result: copy it in
            got = getattr(out, name)
            if got.data_ptr() != buf.data_ptr():
                buf.copy_(got)

    def _run_serial(self) -> List[RenderOutput]:
        return [self._render(j % self.lanes, j) for j in range(len(self.jobs))]

    def _run_lanes(self) -> List[RenderOutput]:
        if self.lanes == 1:
            return self._run_serial()
        if self.view_lanes > 1:
            main = torch.cuda.current_stream(self.ctx.device)
            for s in self._lane_streams:
                s.wait_stream(main)
            for g in range(self.view_lanes):
                stream = main if g == 0 else self._lane_streams[g - 1]
                with torch.cuda.stream(stream):
                    if self.stagger and g > 0:
                        stream.wait_event(self._raster_events[g - 1])   # raster passes of the previous group are done
                    last = len(self.jobs) - 1
                    for j in range(len(self.jobs)):
                        self._render_group(g, j, self._raster_events[g] if self.stagger and j == last else None)
            for s in self._lane_streams:
                main.wait_stream(s)
            return list(self._static)
        main = torch.cuda.current_stream(self.ctx.device)
        outs: List[RenderOutput] = [None] * len(self.jobs)  # type: ignore[list-item]
        for s in self._lane_streams:       # fork
            s.wait_stream(main)
        for lane in range(self.lanes):
            stream = main if lane == 0 else self._lane_streams[lane - 1]
            with torch.cuda.stream(stream):
                for j in range(lane, len(self.jobs), self.lanes):
                    outs[j] = self._render(lane, j)
        for s in self._lane_streams:       # join
            main.wait_stream(s)
        return outs

    def replay(self) -> List[RenderOutput]:
        self.graph.replay()
        return self.outputs


class BakeGraph:
    """CUDA-graph replay of ONE `CameraProjection` call (projection.py:66-204 of the reference) on fixed shapes.

    A bake is a dozen short kernels (view raster + shading, view prep, unprojection, optional padding / Poisson tail)
    issued from Python; on config C a quarter of the 0.32 ms is launch gaps.  The capture holds the call as it is --
    same kernels, same arithmetic -- so `replay()` returns what the eager call returns.

    Static inputs: `images` (and `masks`) must be float tensors on the device; update them in place (`copy_`) between
    replays.  The mesh texture (the "old" atlas the result is stitched with) is read at replay time from
    `mesh.texture`'s storage.  Not capturable, refused here: IoU rejection with masks (it reads a value back to
    the host, projection.py:125-138), background removal, cameras built inside the call."""

    def __init__(self, proj, images: torch.Tensor, mesh: TexturedMesh, cam: Camera, warmup: int = 2, masks=None,
                 **kwargs):
        from .projection import CameraProjection
        if not isinstance(images, torch.Tensor) or not images.is_cuda or images.dtype != torch.float32:
            raise ValueError("BakeGraph: images must be a float32 tensor on the device")
        if masks is not None and kwargs.get("iou_rejection_threshold", 0.8) is not None:
            raise ValueError("BakeGraph: pass iou_rejection_threshold=None with masks (the rejection test reads the IoU "
                             "back to the host)")
        if masks is not None and (not isinstance(masks, torch.Tensor) or not masks.is_cuda):
            raise ValueError("BakeGraph: masks must be a tensor on the device")
        if kwargs.get("remove_bg") or kwargs.get("warp_images") or cam is None:
            raise ValueError("BakeGraph: remove_bg, warp_images and cam=None are not capturable")
        dev = images.device
        # a projection object (raster context, scratch, uv_precompute cache) of its own: the captured kernels hold
        # pointers into them
        self.proj = CameraProjection(proj.pb_backend, None, str(dev), proj.ctx.context_type)
        self.images, self.masks, self.mesh, self.cam = images, masks, mesh, cam
        vw = kwargs.get("uv_exp_blend_view_weight")
        if vw is not None:   # a host tensor would be uploaded inside the capture
            kwargs["uv_exp_blend_view_weight"] = torch.as_tensor(vw).to(device=dev, dtype=torch.float32).contiguous()
        self.kwargs = kwargs
        mesh.v_nrm  # lazily computed once, outside the capture
        import contextlib
        import io
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), contextlib.redirect_stdout(io.StringIO()):
            for _ in range(max(1, warmup)):   # scratch growth, uv_precompute and index caches happen here
                self._call()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), contextlib.redirect_stdout(io.StringIO()):
            self.output = self._call()

    def _call(self):
        return self.proj(self.images, self.mesh, self.cam, masks=self.masks, **self.kwargs)

    def replay(self):
        """Runs the captured bake; returns the (static) result of the call -- overwritten by the next replay."""
        self.graph.replay()
        return self.output
