"""CUDA-graph replay of render() calls (no reference counterpart: host-side launch path).

One render() is five short kernels (~110 us of GPU time for a 1M-face mesh and six 768^2 views) issued from
Python through ctypes: with several processes sharing the host (one per GPU) the launch path can fall behind
the GPU and the gaps land inside the step.  `RenderGraph` captures the render() calls of a fixed job list --
config D is eight meshes per GPU per step -- into ONE CUDA graph (the kernels keep their programmatic
dependent launch edges), so that a step is a single `cudaGraphLaunch`.

The output tensors are static: every replay overwrites them.  Vertex positions, faces and cameras are read
from the tensors the job list held at capture time -- update those in place (copy_) to render new data of the
same shape.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .camera import Camera
from .mesh import TexturedMesh
from .render import NVDiffRastContextWrapper, RenderOutput, render


class RenderGraph:
    def __init__(self, ctx: NVDiffRastContextWrapper, jobs: Sequence[Tuple[TexturedMesh, Camera]], height: int,
                 width: int, warmup: int = 2, **render_kwargs):
        if not jobs:
            raise ValueError("RenderGraph needs at least one (mesh, camera) job")
        # a context of its own: the captured kernels hold pointers into the context's scratch, which an eager call
        # of a larger shape on a shared context would reallocate
        self.ctx = NVDiffRastContextWrapper(str(ctx.device), ctx.context_type)
        ctx = self.ctx
        self.jobs, self.height, self.width = list(jobs), int(height), int(width)
        self.kwargs = dict(render_kwargs)
        for mesh, _ in self.jobs:
            if self.kwargs.get("render_normal", True):
                mesh.v_nrm  # lazily computed once, outside the capture
        side = torch.cuda.Stream(ctx.device)
        side.wait_stream(torch.cuda.current_stream(ctx.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):   # scratch growth (cudaMalloc) and index caches happen here
                self._run()
        torch.cuda.current_stream(ctx.device).wait_stream(side)
        torch.cuda.synchronize(ctx.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs: List[RenderOutput] = self._run()

    def _run(self) -> List[RenderOutput]:
        return [render(self.ctx, m, c, self.height, self.width, **self.kwargs) for m, c in self.jobs]

    def replay(self) -> List[RenderOutput]:
        self.graph.replay()
        return self.outputs
