"""Size-independent properties, hypothesis-driven: on the oracle here on CPU, and on the GPU
path with random shapes under -m gpu."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from conftest import golden


def _q():
    return golden("q_golden.npz")["q"][0]


def _check_cloud(cloud, d, border=40):
    h, w = d.shape
    n = max(0, w - 2 * border) * max(0, h - 2 * border)
    assert cloud.size == n * 16                                        # count = (W-80)(H-80), no data-dependent filter
    if n == 0:
        return
    pts = cloud.view(np.float32).reshape(h - 2 * border, w - 2 * border, 4)
    assert np.all(pts[..., 3].view(np.uint32) == 0x3F800000)           # pcl::PointXYZ pad word
    crop = d[border:h - border, border:w - border]
    fin = np.isfinite(pts[..., :3]).all(axis=2)
    assert np.array_equal(fin, crop != 0)                              # d == 0 <=> the point is not finite (default Q)
    # row-major order: where the disparity is positive, z > 0 and x/z increases with u, y/z with v
    pos = crop > 0
    if pos.any():
        zs = pts[..., 2][pos]
        assert np.all(zs > 0)
        rx = np.where(pos, pts[..., 0] / pts[..., 2], np.nan)
        with np.errstate(invalid="ignore"):
            cols = np.nanmean(rx, axis=0)
        cols = cols[~np.isnan(cols)]
        assert np.all(np.diff(cols) > 0)


@settings(max_examples=25, deadline=None)
@given(w=st.integers(1, 180), h=st.integers(1, 150), seed=st.integers(0, 2**16))
def test_oracle_cloud_properties(w, h, seed):
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, size=(h, w), dtype=np.uint8).astype(np.float32) * np.float32(0.125)
    _check_cloud(oracle.disparity_cb_f32(d, _q()), d)


@settings(max_examples=20, deadline=None)
@given(w=st.integers(1, 120), h=st.integers(1, 120), seed=st.integers(0, 2**16), k=st.sampled_from([3, 5, 11]))
def test_oracle_median_properties(w, h, seed, k):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    m = oracle.median_blur(img, k)
    assert m.min() >= img.min() and m.max() <= img.max()               # an order statistic stays inside the range
    assert np.array_equal(oracle.median_blur(np.full((h, w), 7, np.uint8), k), np.full((h, w), 7, np.uint8))
    assert np.array_equal(oracle.median_blur(255 - img, k), 255 - m)   # median commutes with order reversal


@pytest.mark.gpu
@settings(max_examples=25, deadline=None)
@given(w=st.integers(1, 700), h=st.integers(1, 400), seed=st.integers(0, 2**16), mono=st.booleans(),
       finite=st.booleans())
def test_gpu_matches_oracle_random_shapes(w, h, seed, mono, finite):
    import disparity_to_point_cloud_b200 as d2pc
    ctx = _gpu_ctx()
    ctx.set_q(_q())
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    img[rng.random((h, w)) < 0.05] = 0
    ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE if finite else d2pc.FILTER_CROP)
    try:
        if mono:
            got, want = ctx.process_mono8(img), oracle.disparity_cb_mono8(img, _q())
        else:
            d = img.astype(np.float32) * np.float32(0.125)
            got, want = ctx.process_f32(d), oracle.disparity_cb_f32(d, _q())
    finally:
        ctx.set_filter_mode(d2pc.FILTER_CROP)
    if finite:
        want = oracle.filter_finite(want)
    assert got.tobytes() == want.tobytes()


_CTX = []


def _gpu_ctx():
    import disparity_to_point_cloud_b200 as d2pc
    if not _CTX:
        _CTX.append(d2pc.Context())
    return _CTX[0]
