"""worldrenderer_b200 -- the geometry path of Tengpaz/WorldRenderer on B200 (sm_100a).

Exports the names of the reference's `mvadapter.utils.mesh_utils` package (its __init__.py:1-20)
plus the uv.* functions, so `import worldrenderer_b200 as mesh_utils` is a drop-in for the
render / bake path.  Everything that touches pixels or texels runs in hand-written CUDA kernels
behind the C ABI of include/wr_b200.h (worldrenderer_b200/lib/libwr_b200.so); importing the package
works without a GPU, calling into it does not.
"""
from .blend import PoissonBlendingSolver
from .camera import (
    Camera,
    get_c2w,
    get_camera,
    get_orthogonal_camera,
    get_orthogonal_projection_matrix,
    get_projection_matrix,
)
from .mesh import TexturedMesh, load_mesh, mesh_use_texture, replace_mesh_texture_and_save
from .projection import CameraProjection, CameraProjectionOutput
from .render import (
    DepthControlNetNormalization,
    DepthNormalizationStrategy,
    NVDiffRastContextWrapper,
    RenderOutput,
    SimpleNormalization,
    Zero123PlusPlusNormalization,
    render,
)
from .graph import BakeGraph, RenderGraph
from .smart_paint import SmartPainter
from .tangent import view_normals_to_tangent_space
from .utils import (
    get_clip_space_position,
    image_to_tensor,
    make_image_grid,
    tensor_to_image,
    transform_points_homo,
)
from .uv import (
    ExponentialBlend,
    RandomChoiceBlend,
    SimpleUVValidityStrategy,
    UVBlendOutput,
    UVPrecomputeOutput,
    UVRenderAttrOutput,
    UVRenderGeometryOutput,
    uv_blend,
    uv_precompute,
    uv_render_attr,
    uv_render_geometry,
)


__all__ = [
    "Camera", "get_c2w", "get_camera", "get_orthogonal_camera", "get_orthogonal_projection_matrix",
    "get_projection_matrix", "TexturedMesh", "load_mesh", "mesh_use_texture", "replace_mesh_texture_and_save",
    "CameraProjection", "CameraProjectionOutput", "DepthControlNetNormalization", "DepthNormalizationStrategy",
    "NVDiffRastContextWrapper", "RenderOutput", "SimpleNormalization", "Zero123PlusPlusNormalization", "render",
    "SmartPainter", "get_clip_space_position", "image_to_tensor", "make_image_grid", "tensor_to_image",
    "transform_points_homo", "ExponentialBlend", "RandomChoiceBlend", "SimpleUVValidityStrategy", "UVBlendOutput",
    "UVPrecomputeOutput", "UVRenderAttrOutput", "UVRenderGeometryOutput", "uv_blend", "uv_precompute",
    "uv_render_attr", "uv_render_geometry", "PoissonBlendingSolver", "RenderGraph", "BakeGraph", "view_normals_to_tangent_space",
]
