"""Property tests (hypothesis) of the oracle: the C implementation and the independent NumPy twin must agree
bit for bit on arbitrary small scenes, and basic raster invariants must hold."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import raster_numpy, shim

coord = st.floats(min_value=-1.625, max_value=1.625, allow_nan=False, width=32)
depth = st.floats(min_value=-1.25, max_value=1.25, allow_nan=False, width=32)
wval = st.one_of(st.just(1.0), st.floats(min_value=-0.5, max_value=3.0, allow_nan=False, width=32))
vertex = st.tuples(coord, coord, depth, wval)
scene = st.lists(st.tuples(vertex, vertex, vertex), min_size=1, max_size=10)
size = st.integers(min_value=1, max_value=20)


def _arrays(tris):
    pos = np.array([v for t in tris for v in t], np.float32)[None]
    pos[0, :, :3] *= np.where(pos[0, :, 3:4] == 1.0, 1.0, np.abs(pos[0, :, 3:4]) + 0.25)  # spread clip coordinates
    tri = np.arange(pos.shape[1], dtype=np.int32).reshape(-1, 3)
    return pos, tri


@settings(max_examples=60, deadline=None)
@given(scene, size, size)
def test_c_oracle_equals_numpy_twin_on_random_scenes(tris, H, W):
    pos, tri = _arrays(tris)
    rast, ids = shim.rasterize(pos, tri, (H, W))
    rast2, ids2 = raster_numpy.rasterize(pos, tri, (H, W))
    np.testing.assert_array_equal(ids, ids2)
    np.testing.assert_array_equal(rast, rast2)


@settings(max_examples=40, deadline=None)
@given(scene, size, size)
def test_raster_invariants(tris, H, W):
    pos, tri = _arrays(tris)
    rast, ids = shim.rasterize(pos, tri, (H, W))
    assert ids.min() >= -1 and ids.max() < tri.shape[0]
    bg = ids < 0
    assert (rast[bg] == 0).all()
    fg = rast[~bg]
    assert ((fg[:, 0] >= 0) & (fg[:, 0] <= 1) & (fg[:, 1] >= 0) & (fg[:, 1] <= 1)).all()
    assert ((fg[:, 2] >= -1) & (fg[:, 2] <= 1)).all()
    np.testing.assert_array_equal(rast[..., 3], (ids + 1).astype(np.float32))
    # permuting the faces permutes the ids of non-tied winners; coverage is unchanged
    perm = np.random.default_rng(0).permutation(tri.shape[0])
    _, ids_p = shim.rasterize(pos, tri[perm], (H, W))
    np.testing.assert_array_equal(ids_p >= 0, ids >= 0)


@settings(max_examples=25, deadline=None)
@given(scene, st.integers(min_value=2, max_value=12))
def test_views_are_independent(tris, H):
    pos, tri = _arrays(tris)
    both = np.concatenate([pos, pos[:, ::-1].copy()], 0)  # second view: same triangles, reversed vertex order
    tri2 = tri
    r, i = shim.rasterize(both, tri2, (H, H))
    r0, i0 = shim.rasterize(both[:1], tri2, (H, H))
    r1, i1 = shim.rasterize(both[1:], tri2, (H, H))
    np.testing.assert_array_equal(i[0], i0[0])
    np.testing.assert_array_equal(i[1], i1[0])
    np.testing.assert_array_equal(r[1], r1[0])


# ------------------------------------------------------------------------------------------------------------
# atlas post-processing: the C statement (wr_oracle_blend.c) against its independent NumPy / Python twin
# ------------------------------------------------------------------------------------------------------------
from oracle import blend_numpy  # noqa: E402

_img_shape = st.tuples(st.integers(min_value=1, max_value=14), st.integers(min_value=1, max_value=14))


@settings(max_examples=40, deadline=None)
@given(_img_shape, st.integers(min_value=0, max_value=2 ** 31 - 1), st.integers(min_value=0, max_value=12),
       st.sampled_from(["src", "max", "avg"]))
def test_poisson_c_oracle_equals_numpy_twin(shape, seed, iters, mode):
    H, W = shape
    rng = np.random.default_rng(seed)
    src = (rng.random((H, W, 3)) * 1.5 - 0.25).astype(np.float32)
    tgt = rng.random((H, W, 3)).astype(np.float32)
    mask = rng.random((H, W)) < 0.6
    np.testing.assert_array_equal(shim.poisson_blend(src, mask, tgt, iters, mode),
                                  blend_numpy.poisson_blend(src, mask, tgt, iters, mode))


@settings(max_examples=40, deadline=None)
@given(_img_shape, st.integers(min_value=0, max_value=2 ** 31 - 1), st.integers(min_value=0, max_value=4),
       st.floats(min_value=0.0, max_value=1.0))
def test_inpaint_c_oracle_equals_python_twin(shape, seed, radius, known):
    H, W = shape
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    mask = rng.random((H, W)) >= known
    np.testing.assert_array_equal(shim.inpaint_u8(img, mask, radius), blend_numpy.inpaint_u8(img, mask, radius))
