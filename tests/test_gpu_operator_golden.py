"""CUDA operators against recordings of the UNMODIFIED reference at the dr.* boundary (tests/golden/operators.npz):
`ctx.rasterize` on the reference's own clip-space positions gives the recorded rast tensor BIT FOR BIT (triangle
ids, coverage, u, v, z/w), `ctx.interpolate` the recorded attribute maps bit for bit.  Tangents: v_tang, the
rendered tangent map and the tangent-space rotation of the normal modality against tests/golden/tangent.npz."""
import os

import numpy as np
import pytest
import torch

import worldrenderer_b200 as wr

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_rasterize_bit_exact_on_reference_clip_positions(wr_ctx):
    g = load("operators.npz")
    dev = wr_ctx.device
    for k in range(int(g["n_rasterize"])):
        res = tuple(int(x) for x in g[f"r{k}_res"])
        rast, _ = wr_ctx.rasterize(torch.from_numpy(g[f"r{k}_pos"]).to(dev), torch.from_numpy(g[f"r{k}_tri"]).to(dev), res)
        got, want = rast.cpu().numpy(), g[f"r{k}_rast"]
        np.testing.assert_array_equal(got[..., 3], want[..., 3], err_msg=f"triangle ids / coverage, call {k}")
        np.testing.assert_array_equal(bits(got), bits(want), err_msg=f"u, v, z/w, call {k}")


def test_interpolate_bit_exact_on_reference_inputs(wr_ctx):
    g = load("operators.npz")
    dev = wr_ctx.device
    for k in range(int(g["n_interpolate"])):
        out, _ = wr_ctx.interpolate(torch.from_numpy(g[f"i{k}_attr"]).to(dev), torch.from_numpy(g[f"i{k}_rast"]).to(dev),
                                    torch.from_numpy(g[f"i{k}_tri"]).to(dev))
        np.testing.assert_array_equal(bits(out.cpu().numpy()), bits(g[f"i{k}_out"]), err_msg=f"interpolate call {k}")


def _tangent_mesh(g, dev):
    m = wr.TexturedMesh(v_pos=torch.from_numpy(g["v_pos"]), t_pos_idx=torch.from_numpy(g["t_pos_idx"]).long(),
                        v_tex=torch.from_numpy(g["v_tex"]), t_tex_idx=torch.from_numpy(g["t_tex_idx"]).long(),
                        texture=torch.from_numpy(g["texture"]))
    m.set_stitched_mesh(m.v_pos, m.t_pos_idx)
    m.to(dev)
    return m


def test_vertex_tangents_match_reference(wr_ctx):
    g = load("tangent.npz")
    m = _tangent_mesh(g, wr_ctx.device)
    np.testing.assert_allclose(m.v_nrm.cpu().numpy(), g["v_nrm"], rtol=1e-5, atol=1e-6)
    # At a dozen vertices of this mesh (the five-fold corners of the icosphere with its per-face atlas cells) the mean
    # of the face tangents either nearly cancels or is nearly parallel to the normal, so that one of the two
    # normalisations amplifies the rounding of the SUM -- whose order is unspecified in the reference too
    # (scatter_add_, mesh.py:149-155) and differs here (float atomics).  Conditioning: |sum of the face tangents| /
    # sum of their lengths, and the length of the mean's component perpendicular to the normal.
    v, vt, t, tt = g["v_pos"], g["v_tex"], g["t_pos_idx"].astype(np.int64), g["t_tex_idx"].astype(np.int64)
    u1, u2 = vt[tt[:, 1]] - vt[tt[:, 0]], vt[tt[:, 2]] - vt[tt[:, 0]]
    e1, e2 = v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]]
    den = u1[:, 0:1] * u2[:, 1:2] - u1[:, 1:2] * u2[:, 0:1]
    ft = ((e1 * u2[:, 1:2] - e2 * u1[:, 1:2]) / den).astype(np.float64)
    acc, mag = np.zeros((v.shape[0], 3)), np.zeros(v.shape[0])
    for k in range(3):
        np.add.at(acc, t[:, k], ft)
        np.add.at(mag, t[:, k], np.linalg.norm(ft, axis=1))
    well = np.linalg.norm(acc, axis=1) > 1e-3 * mag
    mean = acc / np.maximum(np.linalg.norm(acc, axis=1, keepdims=True), 1e-30)
    nrm = g["v_nrm"].astype(np.float64)
    well &= np.linalg.norm(mean - (mean * nrm).sum(1, keepdims=True) * nrm, axis=1) > 1e-3
    assert well.mean() > 0.95
    got = m.v_tang.cpu().numpy()
    np.testing.assert_allclose(got[well], g["v_tang"][well], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    # a vertex no face refers to: NaN, as in the reference (0 / 0)
    m2 = _tangent_mesh(g, wr_ctx.device)
    m2.v_pos = torch.cat([m2.v_pos, torch.zeros(1, 3, device=m2.v_pos.device)])
    m2.set_stitched_mesh(m2.v_pos, m2.t_pos_idx)
    t = m2.v_tang
    assert torch.isnan(t[-1]).all() and torch.isfinite(t[:-1]).all()


def test_rendered_tangent_and_tangent_space_match_reference(wr_ctx):
    g = load("tangent.npz")
    dev = wr_ctx.device
    m = _tangent_mesh(g, dev)
    cam = wr.Camera(c2w=torch.linalg.inv(torch.from_numpy(g["w2c"])).to(dev), w2c=torch.from_numpy(g["w2c"]).to(dev),
                    proj_mtx=torch.eye(4)[None].repeat(6, 1, 1).to(dev), mvp_mtx=torch.from_numpy(g["mvp"]).to(dev),
                    cam_pos=torch.zeros(6, 3, device=dev))
    out = wr.render(wr_ctx, m, cam, 64, 64, render_attr=False, render_depth=False, render_normal=True, render_tangent=True)
    mask = out.mask.cpu().numpy()
    same = mask == g["mask"]
    assert (~same).mean() <= 2e-4   # the reference's torch.matmul clip transform (tests/test_oracle_golden.py)
    sel = same & g["mask"]
    # pixels whose triangle touches one of the ill-conditioned vertices of test_vertex_tangents_match_reference
    # follow their own (equally valid) tangent there: compare where the GPU's v_tang agrees with the recording
    okv = np.abs(m.v_tang.cpu().numpy() - g["v_tang"]).max(1) < 1e-5
    raw = wr.render.__globals__["render_geometry_raw"](wr_ctx, m, cam, 64, 64, want_tri_id=True)
    tid = raw["tri_id"].cpu().numpy()
    okp = okv[g["t_pos_idx"].astype(np.int64)].all(1)[np.maximum(tid, 0)] & (tid >= 0)
    sel_t = sel & okp
    assert sel_t.sum() > 0.8 * sel.sum()
    np.testing.assert_allclose(out.tangent.cpu().numpy()[sel_t], g["tangent"][sel_t], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out.normal.cpu().numpy()[sel], g["normal"][sel], rtol=1e-5, atol=2e-6)
    # the rotation kernel alone, on the reference's own maps
    ro = wr.RenderOutput(normal=torch.from_numpy(g["normal"]).to(dev), tangent=torch.from_numpy(g["tangent"]).to(dev))
    ts = wr.view_normals_to_tangent_space(torch.from_numpy(g["normal_images"]), ro)
    np.testing.assert_allclose(ts.cpu().numpy(), g["tangent_space"], rtol=1e-5, atol=2e-6)
    # and end to end: rendered maps -> tangent space
    ts2 = wr.view_normals_to_tangent_space(torch.from_numpy(g["normal_images"]), out)
    np.testing.assert_allclose(ts2.cpu().numpy()[sel_t], g["tangent_space"][sel_t], rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        wr.view_normals_to_tangent_space(torch.from_numpy(g["normal_images"]), wr.RenderOutput(normal=ro.normal))
