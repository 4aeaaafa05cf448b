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
chain into a slice of the job's output tensors (the vertex pass runs once per group).

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
                 width: int, warmup: int = 2, lanes: int = 1, view_lanes: int = 1, **render_kwargs):
        if not jobs:
            raise ValueError("RenderGraph needs at least one (mesh, camera) job")
        self.jobs, self.height, self.width = list(jobs), int(height), int(width)
        self.kwargs = dict(render_kwargs)
        self.view_lanes = max(1, int(view_lanes))
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

    def _render_group(self, group: int, job: int) -> None:
        """Views `group` of job `job` into the matching slices of the job's static outputs."""
        groups = self._groups(job)
        if group >= len(groups) or groups[group].start == groups[group].stop:
            return
        sl = groups[group]
        m, c = self.jobs[job]
        full = self._static[job]
        bufs = {name: getattr(full, name)[sl] for name in ("mask", "pos", "depth", "normal", "attr", "tangent")
                if getattr(full, name, None) is not None}
        out = render(self.ctxs[group], m, c[sl], self.height, self.width, _out_buffers=bufs, **self.kwargs)
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
                    for j in range(len(self.jobs)):
                        self._render_group(g, j)
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
