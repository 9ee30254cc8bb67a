"""The CPU oracle against the committed cv2 fixtures (tests/golden/make_golden.py).

This is the pin of SURVEY.md 8(c): the reference has no tests of its own, so
the oracle is held to the OpenCV entry points the reference calls.
"""
import numpy as np
import pytest

import oracle
from conftest import assert_same_bits, golden


def test_q_matches_stereo_rectify():
    g = golden("q_golden.npz")
    for p, q in zip(g["params"], g["q"]):
        mine = oracle.q_from_intrinsics(*p)
        assert_same_bits(mine, q, f"Q for {p}")


def test_q_default_known_answer():
    # SURVEY.md A.1, hex-exact
    q = oracle.q_from_intrinsics()
    assert q[0, 3].hex() == "-0x1.77ffde0000000p+8"
    assert q[1, 3] == -240.0 and q[2, 3] == 713.5
    assert q[3, 2].hex() == "0x1.638e38e38e38ep+3"
    assert q[3, 3] == 0.0 and np.signbit(q[3, 3])
    assert q[2, 2] == 0.0 and q[0, 0] == 1.0 and q[1, 1] == 1.0


@pytest.mark.parametrize("disp,qname", [("d_s3", "ref"), ("d_s3", "gen"), ("d_s4", "ref"), ("d_s4", "gen")])
def test_reproject_bit_exact(disp, qname):
    g = golden("reproject_golden.npz")
    q = g["q_ref"] if qname == "ref" else g["q_generic"]
    want = g[f"xyz_{disp[2:]}_{qname}"]
    got = oracle.reproject_image_to_3d(g[disp], q)
    assert_same_bits(got, want, f"reprojectImageTo3D {disp} {qname}")


def test_reproject_zero_disparity_known_answers(q_default):
    # d == 0 -> +-inf, NaN on row 240
    d = np.zeros((482, 752), dtype=np.float32)
    xyz = oracle.reproject_image_to_3d(d, q_default)
    assert np.all(np.isneginf(xyz[:, :376, 0])) and np.all(np.isposinf(xyz[:, 376:, 0]))
    assert np.all(np.isneginf(xyz[:240, :, 1])) and np.all(np.isposinf(xyz[241:, :, 1]))
    assert np.all(np.isnan(xyz[240, :, 1]))
    assert np.all(np.isposinf(xyz[:, :, 2]))
    # x86 default NaN, the byte pattern the reference node publishes
    assert np.all(xyz[240, :, 1].view(np.uint32) == 0xFFC00000)


def test_median_blur():
    g = golden("median_golden.npz")
    for i in range(4):
        img = g[f"img{i}"]
        assert_same_bits(oracle.median_blur(img, 11), g[f"m11_{i}"], f"median 11 #{i}")
        assert_same_bits(oracle.median_blur(img, 3), g[f"m3_{i}"], f"median 3 #{i}")


def test_convert_is_exact_eighths():
    img = np.arange(256, dtype=np.uint8).reshape(16, 16)
    f = oracle.convert_u8_f32(img)
    assert np.array_equal(f, img.astype(np.float32) / 8)


def test_full_callback_mono8():
    g = golden("callback_golden.npz")
    for i in range(3):
        img = g[f"img{i}"]
        cloud = oracle.disparity_cb_mono8(img, g["q"])
        assert cloud.size == oracle.n_points(img.shape[1], img.shape[0]) * 16
        assert_same_bits(cloud, g[f"cloud{i}"], f"callback #{i}")
        pts = cloud.view(np.float32).reshape(-1, 4)
        assert np.all(pts[:, 3].view(np.uint32) == 0x3F800000)


def test_callback_stages_compose():
    g = golden("callback_golden.npz")
    img = g["img0"]
    med = oracle.median_blur(img, 11)
    real = oracle.convert_u8_f32(med)
    xyz = oracle.reproject_image_to_3d(real, g["q"])
    assert_same_bits(oracle.crop_pack(xyz), g["cloud0"], "staged callback")
    assert_same_bits(oracle.disparity_cb_f32(real, g["q"]), g["cloud0"], "float entry")


def test_point_count_small_frames(q_default):
    for w, h in [(80, 80), (81, 81), (79, 200), (200, 40), (1, 1)]:
        img = np.zeros((h, w), dtype=np.uint8)
        assert oracle.disparity_cb_mono8(img, q_default).size == max(0, w - 80) * max(0, h - 80) * 16


def test_fusion_golden():
    g = golden("fusion_golden.npz")
    for i in range(3):
        ox, oy = (int(v) for v in g[f"off_{i}"])
        fused, combined = oracle.fuse(g[f"d1_{i}"], g[f"d2_{i}"], g[f"s1_{i}"], g[f"s2_{i}"], ox, oy)
        assert_same_bits(fused, g[f"fused_{i}"], f"fused #{i}")
        assert_same_bits(combined, g[f"combined_{i}"], f"combined #{i}")


def test_grad_filter_ratio_table_and_quirk():
    g = golden("fusion_golden.npz")
    table = g["ratio_table"]
    mine = np.array([[oracle.grad_filter(a, b, 110, 110) for b in range(256)] for a in range(256)], dtype=np.uint8)
    assert_same_bits(mine, table, "ratio branch table")
    # (4k,5k) passes 0.8 < rd because float32(0.8) > 0.8
    for k in range(1, 52):
        assert oracle.grad_filter(4 * k, 5 * k, 110, 110) == (9 * k) // 2
    assert oracle.grad_filter(5, 4, 110, 110) == 0  # ratio exactly 1.25 fails the strict <
    assert oracle.grad_filter(10, 0, 110, 110) == 0  # inf
    assert oracle.grad_filter(0, 0, 110, 110) == 0  # NaN


def test_grad_filter_branches():
    assert oracle.grad_filter(50, 60, 10, 20) == 50
    assert oracle.grad_filter(50, 60, 20, 10) == 60
    assert oracle.grad_filter(230, 60, 10, 20) == 0  # too close, ratio out of range
    assert oracle.grad_filter(230, 229, 10, 20) == 229  # falls to the average branch: (230+229)/2
    assert oracle.grad_filter(100, 101, 124, 124) == 100
    assert oracle.grad_filter(100, 101, 125, 124) == 0


def test_fusion_geometry_table():
    # SURVEY.md A.6 at launch offsets (-7, 15)
    for (w, h), r1, r2, rc in [((640, 480), (77, 15, 465, 465), (7, 61, 465, 465), (80, 0, 465, 465)),
                               ((752, 480), (133, 15, 465, 465), (7, 117, 465, 465), (136, 0, 465, 465)),
                               ((1280, 720), (277, 15, 705, 705), (7, 261, 705, 705), (280, 0, 705, 705))]:
        assert oracle.crop_to_square(w, h, -7, 15, 15) == (0, r1)
        assert oracle.crop_to_square(h, w, 7, -15, 15) == (0, r2)
        assert oracle.crop_to_square(w, h, 0, 0, 15) == (0, rc)


def test_rotate_cw():
    img = np.arange(12, dtype=np.uint8).reshape(3, 4)
    assert np.array_equal(oracle.rotate_cw(img), np.rot90(img, -1))


def test_pointcloud2_wire_layout():
    import struct
    pts = np.arange(32, dtype=np.uint8)
    msg = oracle.serialize_pointcloud2(pts, seq=7, sec=11, nsec=13)
    fid = b"/camera_optical_frame"
    want = struct.pack("<III", 7, 11, 13) + struct.pack("<I", len(fid)) + fid + struct.pack("<III", 1, 2, 3)
    for i, nm in enumerate((b"x", b"y", b"z")):
        want += struct.pack("<I", 1) + nm + struct.pack("<IBI", 4 * i, 7, 1)
    want += struct.pack("<BIII", 0, 16, 32, 32) + pts.tobytes() + b"\x00"
    assert msg == want


def test_score_preprocess_chain_cv2():
    g = golden("score_chain_golden.npz")
    for i in range(3):
        s1, s2 = g[f"s1_{i}"], g[f"s2_{i}"]
        ox, oy = (int(v) for v in g[f"off_{i}"])
        h, w = s1.shape
        _, r1 = oracle.crop_to_square(w, h, ox, oy, oy)
        _, r2 = oracle.crop_to_square(h, w, -ox, -oy, oy)
        assert_same_bits(oracle.score_preprocess(s1, r1, False), g[f"pre1_{i}"], f"MatchingScoreCb1 #{i}")
        assert_same_bits(oracle.score_preprocess(oracle.rotate_cw(s2), r2, True), g[f"pre2_{i}"], f"MatchingScoreCb2 #{i}")


def test_sepfilter_gaussian_cv2():
    """The first blur of MatchingScoreCb1/2 runs on a submatrix, i.e. as sepFilter2D with float32 kernels
    (oracle/d2pc_oracle.c gauss13_sepfilter): stand-alone images straight from cv2.sepFilter2D (vector / tail
    columns at four widths), ROIs of a frame from the generator's cv2-pinned restatement."""
    g = golden("sepfilter_golden.npz")
    for i in range(4):
        img = g[f"img{i}"]
        h, w = img.shape
        assert_same_bits(oracle.gaussian_blur_u8(img, (0, 0, w, h), 13, 3.0, True), g[f"sep{i}"], f"stand-alone #{i}")
    for i, r in enumerate(g["rects"]):
        assert_same_bits(oracle.gaussian_blur_u8(g["frame"], tuple(int(v) for v in r), 13, 3.0, True), g[f"roi{i}"],
                         f"ROI #{i}")
    # and the fixed-point path is what a non-submatrix source still gets (score_golden.npz g13_* are cv2.GaussianBlur)
    sg = golden("score_golden.npz")
    for i in range(2):
        base = sg[f"score{i}"]
        h, w = base.shape
        assert_same_bits(oracle.gaussian_blur_u8(base, (0, 0, w, h), 13, 3.0, False), sg[f"g13_{i}"], f"fixed point #{i}")


def test_colorize_depth_known_answers():
    # src/depth_map_fusion.cpp:304-358, bytes in the order the reference stores them
    ramp = np.arange(256, dtype=np.uint8).reshape(1, 256)
    c = oracle.colorize_depth(ramp)[0]
    assert tuple(c[0]) == (0, 0, 0) and tuple(c[1]) == (0, 0, 0)     # d = uchar(40 + 0.8 g) == 40 -> black
    assert c[2:].max(axis=1).min() == 255                            # S = V = 1: one channel is always saturated
    assert tuple(c[5]) == (0, 101, 255) and tuple(c[128]) == (46, 255, 0) and tuple(c[255]) == (255, 0, 12)
