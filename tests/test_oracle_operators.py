"""The dr.* operator boundary and the tangent helpers of the oracle against recordings of the UNMODIFIED reference
(tests/golden/operators.npz, tangent.npz; oracle/gen_golden.py operator_cases / tangent_cases).  Runs on CPU.

operators.npz holds the clip-space positions the reference itself built (its torch.matmul, its UV -> clip
construction) and handed to dr.rasterize / dr.interpolate, with what came back -- at generation time the C oracle
answered those calls, so these tests pin the oracle's operators against their own recorded behaviour on the
reference's inputs; tests/test_gpu_operator_golden.py then holds the CUDA operators to the same recordings bit for
bit, which closes the reference -> GPU chain at the operator boundary without a restated clip transform."""
import os

import numpy as np

from oracle import render_oracle, shim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_rasterize_recordings_reproduce():
    g = load("operators.npz")
    assert int(g["n_rasterize"]) == 4
    for k in range(int(g["n_rasterize"])):
        rast, ids = shim.rasterize(g[f"r{k}_pos"], g[f"r{k}_tri"], tuple(int(x) for x in g[f"r{k}_res"]))
        np.testing.assert_array_equal(bits(rast), bits(g[f"r{k}_rast"]), err_msg=f"rasterize call {k}")
        np.testing.assert_array_equal(ids, g[f"r{k}_rast"][..., 3].astype(np.int64) - 1)
        assert (ids >= 0).any()


def test_interpolate_recordings_reproduce():
    g = load("operators.npz")
    assert int(g["n_interpolate"]) == 8
    for k in range(int(g["n_interpolate"])):
        out = shim.interpolate(g[f"i{k}_attr"], g[f"i{k}_rast"], g[f"i{k}_tri"])
        np.testing.assert_array_equal(bits(out), bits(g[f"i{k}_out"]), err_msg=f"interpolate call {k}")


def test_vertex_tangents_match_reference():
    g = load("tangent.npz")
    t = render_oracle.vertex_tangents(g["v_pos"], g["t_pos_idx"], g["v_tex"], g["t_tex_idx"], g["v_nrm"])
    assert np.isfinite(g["v_tang"]).all()
    np.testing.assert_allclose(t, g["v_tang"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(t, axis=1), 1.0, atol=1e-5)   # unit length
    assert np.median(np.abs((t * g["v_nrm"]).sum(1))) < 1e-6                 # perpendicular to the normal


def test_rendered_tangent_matches_reference():
    """render.py:280-284 = interpolate(v_tang) + normalize on the recorded rast of the same render call."""
    g, o = load("tangent.npz"), load("operators.npz")
    rast = o["r0_rast"]  # first recorded call: the 6-view sphere render of the same mesh and cameras
    np.testing.assert_array_equal(g["mask"], rast[..., 3] > 0)
    t = shim.interpolate(g["v_tang"][None], rast, g["t_pos_idx"])
    t = t / np.maximum(np.sqrt((t * t).sum(-1, keepdims=True)), 1e-12)
    m = g["mask"]
    np.testing.assert_allclose(t[m], g["tangent"][m], rtol=1e-5, atol=1e-6)


def test_tangent_space_normals_match_reference():
    g = load("tangent.npz")
    from worldrenderer_b200.tangent import CANONICAL_VIEW_TANGENTS
    out = render_oracle.tangent_space_normals(g["normal"], g["tangent"], g["normal_images"], np.asarray(CANONICAL_VIEW_TANGENTS, np.float32))
    np.testing.assert_allclose(out, g["tangent_space"], rtol=1e-5, atol=2e-6)
    assert g["tangent_space"][g["mask"]].std() > 0.05
