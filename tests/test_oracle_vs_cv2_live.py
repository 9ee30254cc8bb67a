"""Oracle against cv2 run live, at sizes the fixtures are too small for.

Skipped where cv2 is not importable.  cv2 is only ever the checker of the
oracle; neither the product nor the GPU parity tests touch it.
"""
import numpy as np
import pytest

import oracle

cv2 = pytest.importorskip("cv2")


def test_reproject_4k_width_bit_exact(q_default):
    rng = np.random.default_rng(1004)
    d = rng.uniform(0.1, 32.0, size=(96, 3840)).astype(np.float32)
    d[::7, ::13] = 0.0
    want = cv2.reprojectImageTo3D(d, q_default)
    got = oracle.reproject_image_to_3d(d, q_default)
    assert got.tobytes() == want.tobytes()


def test_median11_752x480_uniform():
    rng = np.random.default_rng(1000)
    img = rng.integers(0, 256, size=(480, 752), dtype=np.uint8)
    assert np.array_equal(oracle.median_blur(img, 11), cv2.medianBlur(img, 11))


def test_q_random_intrinsics():
    rng = np.random.default_rng(5)
    for _ in range(200):
        fx, fy = rng.uniform(200, 2000, 2)
        cx, cy = rng.uniform(100, 700), rng.uniform(50, 450)
        b = rng.uniform(0.01, 1.0)
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
        out = cv2.stereoRectify(K, np.zeros((5, 1)), K, np.zeros((5, 1)), (752, 480), np.eye(3),
                                np.array([[-b], [0.0], [0.0]]))
        assert oracle.q_from_intrinsics(fx, fy, cx, cy, b).tobytes() == out[4].tobytes()


@pytest.mark.parametrize("n", [705, 450, 465, 96])
def test_sepfilter_gaussian_live_at_the_reference_sizes(n):
    """cv::GaussianBlur on a submatrix = sepFilter2D with float32 kernels (what the reference's first blur runs as,
    src/depth_map_fusion.cpp:70-71): the oracle against cv2.sepFilter2D at the crop sizes of the BASELINE frames
    (705 at 1280x720, 465 at 752x480 / 640x480 with the launch offsets).  A different cv2 build (other SIMD width,
    no FMA) fails here loudly rather than silently."""
    rng = np.random.default_rng(n)
    k = cv2.getGaussianKernel(13, 3.0, cv2.CV_32F)
    for _ in range(3):
        img = rng.integers(0, 256, size=(n, n), dtype=np.uint8)
        want = cv2.sepFilter2D(img, cv2.CV_8U, k, k)
        assert np.array_equal(oracle.gaussian_blur_u8(img, (0, 0, n, n), 13, 3.0, True), want)
    # non-isolated border: a ROI whose tails coincide with the frame's (frame width % 32 == 0, ROI flush right)
    frame = rng.integers(0, 256, size=(n + 20, 32 * ((n + 64) // 32)), dtype=np.uint8)
    full = cv2.sepFilter2D(frame, cv2.CV_8U, k, k)
    w = 32 * (n // 32)
    x = frame.shape[1] - w
    got = oracle.gaussian_blur_u8(frame, (x, 7, w, n), 13, 3.0, True)
    assert np.array_equal(got, full[7:7 + n, x:x + w])
