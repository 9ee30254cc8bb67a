#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ with Python cv2.

Run here (CPU container, cv2 4.13.0):  python tests/golden/make_golden.py

The reference (PX4/disparity_to_point_cloud) has no tests or fixtures and can
not be built in this image, and its arithmetic lives in OpenCV.  So the
fixtures are produced by driving the SAME OpenCV entry points the reference
calls, with the reference's arguments (file:line relative to /root/reference):

  stereoRectify        include/disparity_to_point_cloud/disparity_to_point_cloud.hpp:90-104
  medianBlur(.., 11)   src/disparity_to_point_cloud.cpp:55-57
  convertTo(32F, 1/8)  src/disparity_to_point_cloud.cpp:60-61
  reprojectImageTo3D   src/disparity_to_point_cloud.cpp:63-64
  crop 40 + PointXYZ   src/disparity_to_point_cloud.cpp:69-85 (numpy restatement)
  transpose + flip     src/depth_map_fusion.cpp:268-273
  Mat(Rect) ROI        src/depth_map_fusion.cpp:237-265 (numpy slicing)
  medianBlur(.., 3)    src/depth_map_fusion.cpp:124
  GaussianBlur/Sobel/threshold  src/depth_map_fusion.cpp:64-99

Nothing here is imported by the product or by the tests at run time; the
tests only read the .npz files this script wrote.
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cv_q(fx, fy, cx, cy, b, size=(752, 480)):
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)
    D = np.zeros((5, 1))
    R = np.eye(3)
    t = np.array([[-b], [0], [0]], dtype=np.float64)
    Q = np.eye(4)
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K, D, K, D, size, R, t)
    return Q


def scene_u8(rng, h, w):
    """S2 'scene' distribution of SURVEY.md 8(d): ramp + blobs + holes + salt."""
    v = np.arange(h, dtype=np.float64)[:, None]
    img = 8.0 * (2.0 + 20.0 * v / h) + np.zeros((1, w))
    for _ in range(6):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        ry, rx = rng.integers(4, max(5, h // 3)), rng.integers(4, max(5, w // 3))
        img[max(0, cy - ry):cy + ry, max(0, cx - rx):cx + rx] += 16.0 * rng.uniform(0.5, 3.0)
    img = np.clip(img, 0, 255).astype(np.uint8)
    for _ in range(4):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        img[cy:cy + rng.integers(6, 30), cx:cx + rng.integers(6, 30)] = 0
    salt = rng.random((h, w)) < 0.01
    img[salt] = rng.integers(0, 256, size=int(salt.sum()), dtype=np.uint8)
    return img


def crop_pack(xyz, border=40):
    h, w, _ = xyz.shape
    c = xyz[border:h - border, border:w - border, :]
    pts = np.empty((c.shape[0] * c.shape[1], 4), dtype=np.float32)
    pts[:, :3] = c.reshape(-1, 3)
    pts[:, 3] = 1.0
    return pts.view(np.uint8).reshape(-1)


def rotate_cw(m):
    r = cv2.transpose(m)
    return cv2.flip(r, 1)


def crop_to_square(cols, rows, ox, oy, member_oy):
    num_cols = cols - abs(ox)
    num_rows = rows - abs(oy)
    n = min(cols, rows) - max(abs(ox), abs(member_oy))
    if num_cols < num_rows:
        sc = max(0, ox)
        sr = max(0, oy + int((num_rows - num_cols) / 2))
    else:
        sc = max(0, ox + int((num_cols - num_rows) / 2))
        sr = max(0, oy)
    return sc, sr, n


def grad_filter_np(d1, d2, s1, s2):
    d1i, d2i, s1i, s2i = (a.astype(np.int32) for a in (d1, d2, s1, s2))
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = (d1.astype(np.float32) / d2.astype(np.float32)).astype(np.float64)
    c1 = (s1i < s2i) & (s1i < 100) & (d1i < 230)
    c2 = (s2i < s1i) & (s2i < 100) & (d2i < 230)
    c3 = (0.8 < rel) & (rel < 1.25) & (s1i < 125) & (s2i < 125)
    avg = ((d1i + d2i).astype(np.float32).astype(np.float64) / 2.0).astype(np.int32)
    out = np.where(c1, d1i, np.where(c2, d2i, np.where(c3, avg, 0)))
    return out.astype(np.uint8)


def fuse_cv(d1, d2, s1, s2, ox, oy):
    h, w = d1.shape
    x1, y1, n = crop_to_square(w, h, ox, oy, oy)
    x2, y2, n2 = crop_to_square(h, w, -ox, -oy, oy)
    xc, yc, nc = crop_to_square(w, h, 0, 0, oy)
    assert n == n2 and n <= nc
    cd1 = d1[y1:y1 + n, x1:x1 + n]
    cs1 = s1[y1:y1 + n, x1:x1 + n].copy()
    cd2 = rotate_cw(d2)[y2:y2 + n, x2:x2 + n]
    cs2 = rotate_cw(s2)[y2:y2 + n, x2:x2 + n]
    cont = d2[yc:yc + nc, xc:xc + nc].copy()
    cont[:n, :n] = grad_filter_np(cd1, cd2, cs1, cs2)
    combined = np.minimum(cs1, cs2)
    cont = cv2.medianBlur(cont, 3)
    fused = cont[30:nc - 10, 0:nc - 40].copy()
    return fused, combined


def score_preprocess_cv(score, vertical):
    """src/depth_map_fusion.cpp:64-80 (vertical=False) / :82-99 (True) on the cropped score."""
    g = cv2.GaussianBlur(score, (13, 13), 3.0)
    if vertical:
        g = cv2.Sobel(g, -1, 2, 0, ksize=7, scale=0.03)
    else:
        g = cv2.Sobel(g, -1, 0, 2, ksize=7, scale=0.03)
    _, g = cv2.threshold(g, 30, 255, 0)
    g = cv2.GaussianBlur(g, (21, 21), 10.0)
    return cv2.add(score, cv2.add(g, g))  # score + 2*grad, saturating u8 MatExpr


def _f32(a):
    return np.asarray(a, dtype=np.float32)


def _fma32(k, v, acc):
    """float32 fused multiply-add, elementwise: k scalar, v / acc float32 arrays.  k * v is exact in float64 (24 x 24
    bits), so the only way float64 arithmetic can differ from a true FMA is double rounding: the float64 sum landing
    exactly on a float32 rounding boundary.  That case is refused rather than guessed."""
    a, b = np.float64(k) * v.astype(np.float64), acc.astype(np.float64)
    t = a + b
    bb = t - a
    err = (a - (t - bb)) + (b - bb)  # exact error of the float64 addition (two-sum)
    low = t.view(np.uint64) & np.uint64((1 << 29) - 1)
    assert not np.any((err != 0) & (low == np.uint64(1 << 28))), "double-rounding hazard: regenerate with another seed"
    return t.astype(np.float32)


def gauss13_submatrix(frame, x, y, w, h):
    """cv::GaussianBlur(frame(Rect(x, y, w, h)), dst, Size(13, 13), 3.0) for a CV_8U SUBMATRIX source, default
    border -- the reference's first blur (src/depth_map_fusion.cpp:70-71 / :89-90).  OpenCV 4.x enters its
    fixed-point Gaussian only if `(borderType & BORDER_ISOLATED) || !src.isSubmatrix()`; a submatrix falls through
    to sepFilter2D(src, dst, CV_8U, kx, ky) with kx = ky = getGaussianKernel(13, 3, CV_32F).  Python cannot hand
    cv2 a Mat with the SUBMATRIX flag (a numpy view arrives as a stand-alone Mat with a step), so that call is
    restated here in numpy and pinned against cv2.sepFilter2D where cv2 can express it (check_sepfilter_model):
      * float32 row pass, taps in order 0..12; float32 symmetric column pass; saturate_cast<uchar> (rint);
      * the AVX2 vector loops fuse multiply-add, the scalar tails do not: row pass columns >= w - w % 32, column
        pass columns >= w - w % 4, w being the width of the filtered Mat, i.e. of the ROI;
      * non-isolated border: the pixels around the ROI, reflect-101 only at the edge of the whole frame."""
    k = cv2.getGaussianKernel(13, 3.0, cv2.CV_32F).ravel()
    pad = cv2.copyMakeBorder(frame, 6, 6, 6, 6, cv2.BORDER_REFLECT_101)
    p = pad[y:y + h + 12, x:x + w + 12].astype(np.float32)
    rv, cv_ = w - w % 32, w - w % 4
    acc_f = _f32(k[0] * p[:, 0:w])
    acc_m = acc_f.copy()
    for i in range(1, 13):
        t = p[:, i:i + w]
        acc_f = _fma32(k[i], t, acc_f)
        acc_m = _f32(acc_m + _f32(k[i] * t))
    rows = acc_f
    rows[:, rv:] = acc_m[:, rv:]
    acc_f = _f32(k[6] * rows[6:6 + h])
    acc_m = acc_f.copy()
    for j in range(1, 7):
        t = _f32(rows[6 + j:6 + j + h] + rows[6 - j:6 - j + h])
        acc_f = _fma32(k[6 + j], t, acc_f)
        acc_m = _f32(acc_m + _f32(k[6 + j] * t))
    cols = acc_f
    cols[:, cv_:] = acc_m[:, cv_:]
    return np.clip(np.rint(cols), 0, 255).astype(np.uint8)


def check_sepfilter_model(rng):
    """gauss13_submatrix against cv2.sepFilter2D in the two situations cv2 can be driven into from Python:
    (a) the ROI is a whole stand-alone image (every tail rule, border = the image's own edge);
    (b) a ROI inside a frame, compared with sepFilter2D of the whole frame on the columns whose vector / tail
        status is the same in both calls (the border pixels then come from around the ROI)."""
    k = cv2.getGaussianKernel(13, 3.0, cv2.CV_32F)
    for (h, w) in [(64, 95), (50, 63), (90, 130), (40, 31), (33, 705), (20, 4), (20, 3)]:
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        assert np.array_equal(gauss13_submatrix(img, 0, 0, w, h), cv2.sepFilter2D(img, cv2.CV_8U, k, k)), (h, w)
    frame = rng.integers(0, 256, size=(96, 160), dtype=np.uint8)   # 160 % 32 == 0: the frame call has no tails
    full = cv2.sepFilter2D(frame, cv2.CV_8U, k, k)
    for (x, y, w, h) in [(0, 0, 64, 64), (32, 5, 96, 80), (96, 16, 64, 64), (3, 2, 128, 90)]:
        assert np.array_equal(gauss13_submatrix(frame, x, y, w, h), full[y:y + h, x:x + w]), (x, y, w, h)


def score_chain(frame, x, y, n, vertical):
    """MatchingScoreCb1 (vertical=False, src/depth_map_fusion.cpp:64-80) / Cb2 (True, :82-99) for a score frame
    (already rotated for Cb2) and its cropToSquare rectangle."""
    score = frame[y:y + n, x:x + n].copy()
    if n < frame.shape[0] or n < frame.shape[1]:      # mat(region) is a submatrix: sepFilter2D path
        g = gauss13_submatrix(frame, x, y, n, n)
    else:                                             # the region is the whole frame: fixed-point path
        g = cv2.GaussianBlur(frame, (13, 13), 3.0)
    if vertical:
        g = cv2.Sobel(g, -1, 2, 0, ksize=7, scale=0.03)
    else:
        g = cv2.Sobel(g, -1, 0, 2, ksize=7, scale=0.03)
    _, g = cv2.threshold(g, 30, 255, 0)
    g = cv2.GaussianBlur(g, (21, 21), 10.0)
    return cv2.add(score, cv2.add(g, g))


def main():
    rng = np.random.default_rng(20261018)
    assert cv2.__version__.startswith("4."), cv2.__version__

    # ---- Q ----------------------------------------------------------------
    params = [(714.24, 713.5, 376.0, 240.0, 0.09), (714.24, 713.5, 376.0, 240.0, 0.043)]
    for _ in range(14):
        params.append((rng.uniform(300, 1500), rng.uniform(300, 1500), rng.uniform(200, 600),
                       rng.uniform(100, 400), rng.uniform(0.02, 0.5)))
    params = np.array(params, dtype=np.float64)
    qs = np.stack([cv_q(*p) for p in params])
    np.savez_compressed(os.path.join(HERE, "q_golden.npz"), params=params, q=qs, cv2_version=cv2.__version__)
    q_ref = qs[0]

    # ---- reprojectImageTo3D -------------------------------------------------
    q_generic = rng.uniform(-2, 2, size=(4, 4))
    q_generic[3, 2] = 9.5
    d_s3 = (rng.integers(0, 256, size=(100, 128), dtype=np.uint8).astype(np.float32) * np.float32(0.125))
    d_s4 = rng.uniform(0.1, 32.0, size=(3, 3840)).astype(np.float32)
    d_s4[1, ::97] = 0.0
    np.savez_compressed(
        os.path.join(HERE, "reproject_golden.npz"),
        q_ref=q_ref, q_generic=q_generic, d_s3=d_s3, d_s4=d_s4,
        xyz_s3_ref=cv2.reprojectImageTo3D(d_s3, q_ref), xyz_s3_gen=cv2.reprojectImageTo3D(d_s3, q_generic),
        xyz_s4_ref=cv2.reprojectImageTo3D(d_s4, q_ref), xyz_s4_gen=cv2.reprojectImageTo3D(d_s4, q_generic))

    # ---- medianBlur ----------------------------------------------------------
    med = {}
    for i, (h, w) in enumerate([(5, 7), (11, 11), (37, 53), (64, 96)]):
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        med[f"img{i}"] = img
        med[f"m11_{i}"] = cv2.medianBlur(img, 11)
        med[f"m3_{i}"] = cv2.medianBlur(img, 3)
    np.savez_compressed(os.path.join(HERE, "median_golden.npz"), **med)

    # ---- full DisparityCb on mono8 -------------------------------------------
    cb = {}
    for i, (h, w, kind) in enumerate([(120, 136, "scene"), (104, 150, "uniform"), (97, 101, "scene")]):
        img = scene_u8(rng, h, w) if kind == "scene" else rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        medf = cv2.medianBlur(img, 11)
        # convertTo(CV_32FC1, 1/8) has no Python binding; u8 * 0.125 is exact in float32, so numpy states it
        real = medf.astype(np.float32) * np.float32(1.0 / 8.0)
        xyz = cv2.reprojectImageTo3D(real, q_ref)
        cb[f"img{i}"] = img
        cb[f"cloud{i}"] = crop_pack(xyz)
    cb["q"] = q_ref
    np.savez_compressed(os.path.join(HERE, "callback_golden.npz"), **cb)

    # ---- fusion ----------------------------------------------------------------
    fu = {}
    for i, (h, w, ox, oy) in enumerate([(150, 200, -7, 15), (160, 120, 5, -9), (140, 180, 20, 3)]):
        d1 = scene_u8(rng, h, w)
        d2 = np.clip(d1.astype(np.int32) + rng.integers(-20, 21, size=(h, w)), 0, 255).astype(np.uint8)
        d2 = np.ascontiguousarray(np.rot90(d2, 1))  # so rotating it back roughly aligns with d1
        d2 = cv2.resize(d2, (w, h), interpolation=cv2.INTER_NEAREST)
        s1 = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        s2 = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        fused, combined = fuse_cv(d1, d2, s1, s2, ox, oy)
        fu.update({f"d1_{i}": d1, f"d2_{i}": d2, f"s1_{i}": s1, f"s2_{i}": s2, f"off_{i}": np.array([ox, oy]),
                   f"fused_{i}": fused, f"combined_{i}": combined})
    # every (d1,d2) pair at fixed scores that force the ratio branch (quirk A.5)
    dd1, dd2 = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    sc = np.full((256, 256), 110, dtype=np.uint8)
    fu["ratio_table"] = grad_filter_np(dd1, dd2, sc, sc)
    np.savez_compressed(os.path.join(HERE, "fusion_golden.npz"), **fu)

    # ---- matching-score preprocessing (next-row) ---------
    sp = {}
    for i, (h, w) in enumerate([(96, 96), (130, 130)]):
        base = cv2.GaussianBlur(rng.integers(0, 256, size=(h, w), dtype=np.uint8), (9, 9), 2.5)
        base[h // 3:h // 3 + 3, :] = 250
        base[:, w // 2:w // 2 + 2] = 5
        sp[f"score{i}"] = base
        sp[f"pre_h{i}"] = score_preprocess_cv(base, vertical=False)
        sp[f"pre_v{i}"] = score_preprocess_cv(base, vertical=True)
        sp[f"g13_{i}"] = cv2.GaussianBlur(base, (13, 13), 3.0)
        sp[f"sob_h{i}"] = cv2.Sobel(sp[f"g13_{i}"], -1, 0, 2, ksize=7, scale=0.03)
        sp[f"sob_v{i}"] = cv2.Sobel(sp[f"g13_{i}"], -1, 2, 0, ksize=7, scale=0.03)
    np.savez_compressed(os.path.join(HERE, "score_golden.npz"), **sp)
    # ---- the whole MatchingScoreCb chain on full frames, ROI semantics included ------------------------------
    # cropped_score_k_ is a non-isolated ROI (a submatrix) of the (rotated) frame: the first GaussianBlur sees the
    # pixels around the ROI and runs as sepFilter2D, not as the fixed-point Gaussian (gauss13_submatrix).
    # Everything after that runs on a stand-alone n x n Mat (reflect-101 at its own edge).
    check_sepfilter_model(np.random.default_rng(77))
    ch = {}
    for i, (h, w, ox, oy) in enumerate([(150, 200, -7, 15), (160, 120, 5, -9), (300, 424, -7, 15)]):
        s1 = cv2.GaussianBlur(rng.integers(0, 256, size=(h, w), dtype=np.uint8), (7, 7), 2.0)
        s2 = cv2.GaussianBlur(rng.integers(0, 256, size=(h, w), dtype=np.uint8), (7, 7), 2.0)
        for a in (s1, s2):
            for _ in range(6):
                y, x = int(rng.integers(0, h - 4)), int(rng.integers(0, w - 4))
                a[y:y + 3, :] = int(rng.integers(0, 256))
                a[:, x:x + 2] = int(rng.integers(0, 256))
        x1, y1, n = crop_to_square(w, h, ox, oy, oy)
        x2, y2, n2 = crop_to_square(h, w, -ox, -oy, oy)
        pre1 = score_chain(s1, x1, y1, n, vertical=False)
        pre2 = score_chain(rotate_cw(s2), x2, y2, n, vertical=True)
        if i == 2:  # keep the big case small on disk: store the inputs' seed-free digest rows only
            ch.update({f"s1_{i}": s1, f"s2_{i}": s2, f"off_{i}": np.array([ox, oy]), f"pre1_{i}": pre1, f"pre2_{i}": pre2})
        else:
            ch.update({f"s1_{i}": s1, f"s2_{i}": s2, f"off_{i}": np.array([ox, oy]), f"pre1_{i}": pre1, f"pre2_{i}": pre2})
    np.savez_compressed(os.path.join(HERE, "score_chain_golden.npz"), **ch)
    # ---- the sepFilter2D Gaussian on its own: stand-alone images straight from cv2 (tail rules at several widths)
    # and ROIs of a frame (non-isolated border) from the numpy restatement that check_sepfilter_model pins ---------
    sf = {}
    k13 = cv2.getGaussianKernel(13, 3.0, cv2.CV_32F)
    sf["kernel13"] = k13.ravel()
    for i, (h, w) in enumerate([(40, 95), (33, 130), (50, 31), (24, 64)]):
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        sf[f"img{i}"] = img
        sf[f"sep{i}"] = cv2.sepFilter2D(img, cv2.CV_8U, k13, k13)
    frame = rng.integers(0, 256, size=(90, 150), dtype=np.uint8)
    sf["frame"] = frame
    rects = np.array([[0, 0, 70, 70], [10, 5, 101, 77], [80, 20, 70, 70], [3, 2, 140, 86]])
    sf["rects"] = rects
    for i, (x, y, w, h) in enumerate(rects):
        sf[f"roi{i}"] = gauss13_submatrix(frame, int(x), int(y), int(w), int(h))
    np.savez_compressed(os.path.join(HERE, "sepfilter_golden.npz"), **sf)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
