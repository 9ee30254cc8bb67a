"""GPU parity: depth_map_fusion's matching-score preprocessing (MatchingScoreCb1/2) against the cv2 fixtures and
the oracle, and the fusion entry that consumes the preprocessed caches."""
import numpy as np
import pytest

import oracle
from conftest import assert_same_bits, golden
from disparity_to_point_cloud_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import disparity_to_point_cloud_b200 as d2pc
    with d2pc.Context(offset_x=-7, offset_y=15) as c:
        yield c


def test_score_chain_cv2_fixtures(ctx):
    g = golden("score_chain_golden.npz")
    for i in range(3):
        ox, oy = (int(v) for v in g[f"off_{i}"])
        ctx.set_tuning("offset_x", ox)
        ctx.set_tuning("offset_y", oy)
        assert_same_bits(ctx.preprocess_score(g[f"s1_{i}"], 1), g[f"pre1_{i}"], f"MatchingScoreCb1 #{i}")
        assert_same_bits(ctx.preprocess_score(g[f"s2_{i}"], 2), g[f"pre2_{i}"], f"MatchingScoreCb2 #{i}")
    ctx.set_tuning("offset_x", -7)
    ctx.set_tuning("offset_y", 15)


@pytest.mark.parametrize("w,h,ox,oy", [(96, 96, 0, 0), (129, 160, 0, 0), (200, 65, 3, 0), (90, 40, -2, 1), (705, 64, 0, 0),
                                       (300, 424, -7, 15), (64, 64, 0, 0), (128, 128, 1, 0)])
def test_score_chain_tile_edges_and_blur_paths(ctx, w, h, ox, oy):
    """Partial last tiles (n = 65, 129), a single tile, n smaller than the halos, and the two paths of the first
    blur: a crop that is the whole frame (square frame, zero offsets) is not a submatrix and takes OpenCV's
    fixed-point Gaussian, any smaller crop runs as sepFilter2D (float32)."""
    ctx.set_tuning("offset_x", ox)
    ctx.set_tuning("offset_y", oy)
    try:
        s1, s2 = _score_frame(h, w, 5), _score_frame(h, w, 6)
        rc1, r1 = oracle.crop_to_square(w, h, ox, oy, oy)
        rc2, r2 = oracle.crop_to_square(h, w, -ox, -oy, oy)
        assert rc1 == 0 and rc2 == 0
        assert_same_bits(ctx.preprocess_score(s1, 1), oracle.score_preprocess(s1, r1, False), "score 1")
        assert_same_bits(ctx.preprocess_score(s2, 2), oracle.score_preprocess(oracle.rotate_cw(s2), r2, True), "score 2")
    finally:
        ctx.set_tuning("offset_x", -7)
        ctx.set_tuning("offset_y", 15)


def _score_frame(h, w, seed):
    rng = np.random.default_rng(seed)
    a = synth.s2_scene(h, w, seed)
    a[rng.integers(0, h, 12), :] = 250
    a[:, rng.integers(0, w, 12)] = 3
    return a


@pytest.mark.parametrize("w,h", [(1280, 720), (640, 480), (752, 480)])
def test_score_chain_vs_oracle_and_fusion(ctx, w, h):
    s1, s2 = _score_frame(h, w, 1), _score_frame(h, w, 2)
    _, r1 = oracle.crop_to_square(w, h, -7, 15, 15)
    _, r2 = oracle.crop_to_square(h, w, 7, -15, 15)
    p1, p2 = ctx.preprocess_score(s1, 1), ctx.preprocess_score(s2, 2)
    assert_same_bits(p1, oracle.score_preprocess(s1, r1, False), "score 1")
    assert_same_bits(p2, oracle.score_preprocess(oracle.rotate_cw(s2), r2, True), "score 2")
    # fusion from the preprocessed caches == oracle fusion on frames that carry those caches in their crops
    d1, d2 = synth.s2_scene(h, w, 3), synth.s2_scene(h, w, 4)
    c1 = np.zeros((h, w), np.uint8)
    c1[r1[1]:r1[1] + r1[2], r1[0]:r1[0] + r1[2]] = p1
    rot = np.zeros((w, h), np.uint8)
    rot[r2[1]:r2[1] + r2[2], r2[0]:r2[0] + r2[2]] = p2
    c2 = np.ascontiguousarray(np.rot90(rot, 1))       # undo the clockwise rotation
    fused, combined = ctx.fuse_preprocessed(d1, d2, p1, p2)
    of, oc = oracle.fuse(d1, d2, c1, c2, -7, 15)
    assert_same_bits(fused, of, "fused")
    assert_same_bits(combined, oc, "combined")
