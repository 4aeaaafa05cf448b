"""The C ABI used from a plain C++ host program (examples/c_abi_render.cpp): no Python, no PyTorch in the loop.
The program renders a scene file with wr_vertex_normals + wr_render; the result must equal what the Python
package returns for the same scene (same kernels), and the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest
import torch

import cases
import worldrenderer_b200 as wr
from oracle import render_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_standalone_c_abi_program_matches_python_and_oracle(tmp_path, wr_ctx):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import build as example_build
    exe = example_build.build()
    v, f = cases.terrain_mesh(96, 64)
    cam = cases.canonical_cameras()
    B, H, W = 6, 120, 168
    scene, result = str(tmp_path / "scene.bin"), str(tmp_path / "result.bin")
    with open(scene, "wb") as fh:
        np.array([v.shape[0], f.shape[0], B, H, W], np.int32).tofile(fh)
        v.astype(np.float32).tofile(fh)
        f.astype(np.int32).tofile(fh)
        cam.mvp_mtx.numpy().astype(np.float32).tofile(fh)
        cam.w2c.numpy().astype(np.float32).tofile(fh)
    out = subprocess.run([exe, scene, result], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "covered pixels" in out.stdout
    raw = np.fromfile(result, np.uint8)
    npix = B * H * W
    assert raw.size == npix * (1 + 12 + 4 + 12)
    mask = raw[:npix].reshape(B, H, W).astype(bool)
    pos = raw[npix:npix * 13].view(np.float32).reshape(B, H, W, 3)
    depth = raw[npix * 13:npix * 17].view(np.float32).reshape(B, H, W)
    normal = raw[npix * 17:].view(np.float32).reshape(B, H, W, 3)

    mesh = wr.TexturedMesh(v_pos=torch.tensor(v), t_pos_idx=torch.tensor(f, dtype=torch.int64))
    mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx)
    mesh.to(wr_ctx.device)
    cam.to(wr_ctx.device)
    py = wr.render(wr_ctx, mesh, cam, H, W, render_attr=False)
    np.testing.assert_array_equal(mask, py.mask.cpu().numpy())
    np.testing.assert_array_equal(pos, py.pos.cpu().numpy())       # same kernels: identical
    np.testing.assert_array_equal(depth, py.depth.cpu().numpy())
    # vertex normals are summed with float atomics (order varies between runs): tolerance, not bits
    np.testing.assert_allclose(normal, py.normal.cpu().numpy(), rtol=1e-5, atol=1e-6)
    ref = render_oracle.render(v, f, cam.mvp_mtx.cpu().numpy(), cam.w2c.cpu().numpy(), H, W,
                               v_nrm=render_oracle.vertex_normals(v, f))
    np.testing.assert_array_equal(mask, ref["mask"])
    np.testing.assert_allclose(pos, ref["pos"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(depth, ref["depth"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(normal, ref["normal"], rtol=1e-5, atol=1e-6)
