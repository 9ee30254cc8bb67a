"""Sequence oracle: the STATE the reference's DepthMapFusion keeps between callbacks, restated over the pure
functions of oracle/__init__.py (TEST INFRASTRUCTURE; pinned against the reference's own compiled class in
tests/test_ref_compiled.py through oracle/ref.py).

What it models (all file:line relative to /root/reference):
  * four caches, filled by four independent callbacks; only DisparityCb2 publishes a fused map, and only once all
    four are non-empty (src/depth_map_fusion.cpp:108-111);
  * cropped_score_combined_ = cropped_score_1_ is a shallow cv::Mat copy (:113) and cropped_score_1_ is
    cropped_score_1_grad_ (:77): the merge loop writes min(grad1, grad2) into that one buffer (:118-121), so after
    a fused publish the cached score 1 IS the combined score until the next MatchingScoreCb1 replaces it;
  * every publishWithColor / fused publish carries the header of the message whose callback is running.
"""
from __future__ import annotations

import numpy as np

import oracle


def grad_filter_np(d1, d2, s1, s2):
    """gradFilter (src/depth_map_fusion.cpp:219-235) over arrays: numpy's float32 division is the same IEEE
    operation as the C float division (tests/test_ref_compiled.py checks this function against the compiled
    reference through d2pc_oracle_grad_filter_table)."""
    d1, d2, s1, s2 = (np.asarray(v, dtype=np.int64) for v in (d1, d2, s1, s2))
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = d1.astype(np.float32) / d2.astype(np.float32)
    relative = rel.astype(np.float64)  # the comparisons promote the float to double (0.8, 1.25 are double literals)
    take1 = (s1 < s2) & (s1 < 100) & (d1 < 230)
    take2 = ~take1 & (s2 < s1) & (s2 < 100) & (d2 < 230)
    avg = ~take1 & ~take2 & (0.8 < relative) & (relative < 1.25) & (s1 < 125.0) & (s2 < 125.0)
    out = np.zeros(d1.shape, dtype=np.int64)
    out[take1] = d1[take1]
    out[take2] = d2[take2]
    out[avg] = ((d1[avg] + d2[avg]).astype(np.float32) / np.float32(2.0)).astype(np.int64)  # float(d1 + d2) / 2.0, truncated
    return out.astype(np.uint8)


class FusionNodeOracle:
    def __init__(self, offset_x=0, offset_y=0, mode=0):
        self.ox, self.oy, self.mode = offset_x, offset_y, mode
        self.d1 = self.d2 = self.s1 = self.s2 = None  # cropped_depth_1_, cropped_depth_2_ (rotated), cropped_score_1_, _2_

    def _crop(self, img, which):
        h, w = img.shape
        if which == 1:
            rc, r = oracle.crop_to_square(w, h, self.ox, self.oy, self.oy)           # :48 / :66
            src = img
        else:
            rc, r = oracle.crop_to_square(h, w, -self.ox, -self.oy, self.oy)         # :56-57 / :84-85
            src = oracle.rotate_cw(np.ascontiguousarray(img))
        if rc != 0:
            raise ValueError("cropToSquare leaves the image")
        return src, r

    def callback(self, which, img, hdr=(0, 0, 0)):
        """which: 1 DisparityCb1, 2 DisparityCb2, 3 MatchingScoreCb1, 4 MatchingScoreCb2.
        Returns [(topic, encoding, array, hdr)] in publish order."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        out = []
        if which in (1, 2):
            src, r = self._crop(img, which)
            crop = src[r[1]:r[1] + r[3], r[0]:r[0] + r[2]].copy()
            if which == 1:
                self.d1 = crop
                out.append(("/cropped_depth_1", "rgb8", oracle.colorize_depth(crop), hdr))                # :50-51
            else:
                self.d2 = crop
                out.append(("/cropped_depth_2", "rgb8", oracle.colorize_depth(crop), hdr))                # :59-60
                out += self._publish_fused(img, hdr)                                                       # :61
        else:
            src, r = self._crop(img, 1 if which == 3 else 2)
            pre = oracle.score_preprocess(src, r, vertical=(which == 4))                                   # :70-77 / :89-96
            if which == 3:
                self.s1 = pre
                out.append(("/cropped_score_1", "mono8", pre.copy(), hdr))                                 # :79
            else:
                self.s2 = pre
                out.append(("/cropped_score_2", "mono8", pre.copy(), hdr))                                 # :98
        return out

    def _publish_fused(self, img2, hdr):
        if self.d1 is None or self.d2 is None or self.s1 is None or self.s2 is None:                        # :108-111
            return []
        h, w = img2.shape
        rc, rcont = oracle.crop_to_square(w, h, 0, 0, self.oy)                                              # :106
        if rc != 0:
            raise ValueError("cropToSquare leaves the image")
        n = self.d2.shape[0]
        if self.d1.shape != (n, n) or self.s1.shape != (n, n) or self.s2.shape != (n, n) or n > rcont[2]:
            raise ValueError("cached images differ in size (cv::Mat::at would read out of bounds)")
        cont = img2[rcont[1]:rcont[1] + rcont[3], rcont[0]:rcont[0] + rcont[2]].copy()
        a1, a2 = self.d1.astype(np.int64), self.d2.astype(np.int64)
        c1, c2 = self.s1.astype(np.int64), self.s2.astype(np.int64)
        fused = np.empty((n, n), dtype=np.uint8)
        if self.mode == 0:
            fused = grad_filter_np(a1, a2, c1, c2)
        else:
            flat = [oracle.fuse_rule(self.mode, int(x), int(y), int(u), int(v))
                    for x, y, u, v in zip(a1.ravel(), a2.ravel(), c1.ravel(), c2.ravel())]
            fused = (np.array(flat, dtype=np.int64) & 255).astype(np.uint8).reshape(n, n)
        cont[:n, :n] = fused                                                                                # :117
        self.s1 = np.minimum(self.s1, self.s2)                                                              # :113, :118-121 (aliased)
        med = oracle.median_blur(np.ascontiguousarray(cont), 3)                                             # :124
        nc = cont.shape[0]
        out_img = med[30:nc - 10, 0:nc - 40].copy()                                                         # :130
        return [("/combined_score", "mono8", self.s1.copy(), hdr),                                          # :126-127
                ("/gradient", "rgb8", oracle.colorize_depth(out_img), hdr),                                 # :132
                ("/fused_depth_map", "mono8", out_img, hdr)]                                                # :134-136
