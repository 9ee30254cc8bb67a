"""ctypes view of the CPU ORACLE (oracle/d2pc_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs may import this module,
and only as the checker or the CPU baseline.  The product package
(disparity_to_point_cloud_b200) never imports it.

Parity pin: see oracle/d2pc_oracle.h -- OpenCV's arithmetic against cv2 4.13.0
fixtures in tests/golden/, the reference's own logic against the reference's
own sources compiled unmodified against stand-in headers (oracle/_ref); the
reference ships no tests and its build system cannot run here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libd2pc_oracle.so")

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "d2pc_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []),
                       check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.d2pc_oracle_q_from_intrinsics.argtypes = [C.c_double] * 5 + [C.c_int, C.c_int, _f64p]
        L.d2pc_oracle_q_from_intrinsics.restype = C.c_int
        L.d2pc_oracle_median_blur_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _u8p, C.c_size_t, C.c_int]
        L.d2pc_oracle_median_blur_u8.restype = None
        L.d2pc_oracle_convert_u8_f32.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _f32p, C.c_size_t, C.c_double]
        L.d2pc_oracle_convert_u8_f32.restype = None
        L.d2pc_oracle_reproject_image_to_3d.argtypes = [_f32p, C.c_int, C.c_int, C.c_size_t, _f64p, _f32p]
        L.d2pc_oracle_reproject_image_to_3d.restype = None
        L.d2pc_oracle_crop_pack.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _u8p]
        L.d2pc_oracle_crop_pack.restype = C.c_size_t
        L.d2pc_oracle_disparity_cb_mono8.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _f64p, _u8p]
        L.d2pc_oracle_disparity_cb_mono8.restype = C.c_size_t
        L.d2pc_oracle_disparity_cb_f32.argtypes = [_f32p, C.c_int, C.c_int, C.c_size_t, _f64p, _u8p]
        L.d2pc_oracle_disparity_cb_f32.restype = C.c_size_t
        L.d2pc_oracle_filter_finite.argtypes = [_u8p, C.c_size_t, _u8p]
        L.d2pc_oracle_filter_finite.restype = C.c_size_t
        L.d2pc_oracle_serialize_pointcloud2.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, _u8p,
                                                        C.c_uint32, C.c_uint8, _u8p, C.c_size_t]
        L.d2pc_oracle_serialize_pointcloud2.restype = C.c_size_t
        L.d2pc_oracle_crop_to_square.argtypes = [C.c_int] * 5 + [_i32p]
        L.d2pc_oracle_crop_to_square.restype = C.c_int
        L.d2pc_oracle_rotate_cw.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _u8p]
        L.d2pc_oracle_rotate_cw.restype = None
        L.d2pc_oracle_grad_filter.argtypes = [C.c_int] * 6
        L.d2pc_oracle_grad_filter.restype = C.c_int
        L.d2pc_oracle_grad_filter_table.argtypes = [C.c_int, C.c_int, _u8p]
        L.d2pc_oracle_grad_filter_table.restype = None
        L.d2pc_oracle_fuse_rule.argtypes = [C.c_int] * 5
        L.d2pc_oracle_fuse_rule.restype = C.c_int
        L.d2pc_oracle_fuse.argtypes = [_u8p, _u8p, _u8p, _u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                       C.c_int, _u8p, _u8p, _i32p]
        L.d2pc_oracle_fuse.restype = C.c_int
        L.d2pc_oracle_score_preprocess.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _i32p, C.c_int, _u8p]
        L.d2pc_oracle_score_preprocess.restype = C.c_int
        L.d2pc_oracle_gaussian_blur_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _i32p, C.c_int, C.c_double, C.c_int,
                                                   _u8p, C.c_size_t]
        L.d2pc_oracle_gaussian_blur_u8.restype = C.c_int
        L.d2pc_oracle_sobel7_second_u8.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.d2pc_oracle_sobel7_second_u8.restype = None
        L.d2pc_oracle_colorize_depth.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _u8p]
        L.d2pc_oracle_colorize_depth.restype = None
        L.d2pc_oracle_run_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, _f64p,
                                             _u8p, C.c_size_t, C.c_int]
        L.d2pc_oracle_run_frames.restype = C.c_size_t
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _q(q) -> np.ndarray:
    q = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(16))
    return q


def q_from_intrinsics(fx=714.24, fy=713.5, cx=376.0, cy=240.0, baseline=0.09, rect_w=752, rect_h=480) -> np.ndarray:
    q = np.zeros(16, dtype=np.float64)
    rc = lib().d2pc_oracle_q_from_intrinsics(fx, fy, cx, cy, baseline, rect_w, rect_h, _p(q, _f64p))
    if rc != 0:
        raise ValueError("bad intrinsics")
    return q.reshape(4, 4)


def median_blur(img: np.ndarray, ksize: int) -> np.ndarray:
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    h, w = img.shape
    out = np.empty((h, w), dtype=np.uint8)
    lib().d2pc_oracle_median_blur_u8(_p(img, _u8p), w, h, img.strides[0], _p(out, _u8p), w, ksize)
    return out


def convert_u8_f32(img: np.ndarray, alpha: float = 1.0 / 8.0) -> np.ndarray:
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    h, w = img.shape
    out = np.empty((h, w), dtype=np.float32)
    lib().d2pc_oracle_convert_u8_f32(_p(img, _u8p), w, h, img.strides[0], _p(out, _f32p), w * 4, alpha)
    return out


def reproject_image_to_3d(disp: np.ndarray, q) -> np.ndarray:
    assert disp.dtype == np.float32 and disp.ndim == 2 and disp.strides[1] == 4
    h, w = disp.shape
    out = np.empty((h, w, 3), dtype=np.float32)
    qq = _q(q)
    lib().d2pc_oracle_reproject_image_to_3d(_p(disp, _f32p), w, h, disp.strides[0], _p(qq, _f64p), _p(out, _f32p))
    return out


def n_points(w: int, h: int, border: int = 40) -> int:
    return max(0, w - 2 * border) * max(0, h - 2 * border)


def crop_pack(xyz: np.ndarray, border: int = 40) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    h, w, _ = xyz.shape
    n = n_points(w, h, border)
    out = np.empty(n * 16, dtype=np.uint8)
    got = lib().d2pc_oracle_crop_pack(_p(xyz, _f32p), w, h, border, _p(out, _u8p))
    assert got == n
    return out


def disparity_cb_mono8(img: np.ndarray, q) -> np.ndarray:
    """Full DisparityCb on a mono8 frame -> PointCloud2.data bytes (N*16)."""
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    h, w = img.shape
    out = np.empty(n_points(w, h) * 16, dtype=np.uint8)
    qq = _q(q)
    got = lib().d2pc_oracle_disparity_cb_mono8(_p(img, _u8p), w, h, img.strides[0], _p(qq, _f64p), _p(out, _u8p))
    assert got * 16 == out.size
    return out


def disparity_cb_f32(disp: np.ndarray, q) -> np.ndarray:
    """DisparityCb entered after convertTo (float disparity) -> N*16 bytes."""
    assert disp.dtype == np.float32 and disp.ndim == 2 and disp.strides[1] == 4
    h, w = disp.shape
    out = np.empty(n_points(w, h) * 16, dtype=np.uint8)
    qq = _q(q)
    got = lib().d2pc_oracle_disparity_cb_f32(_p(disp, _f32p), w, h, disp.strides[0], _p(qq, _f64p), _p(out, _u8p))
    assert got * 16 == out.size
    return out


def filter_finite(cloud: np.ndarray) -> np.ndarray:
    cloud = np.ascontiguousarray(cloud, dtype=np.uint8).reshape(-1)
    n = cloud.size // 16
    out = np.empty(n * 16, dtype=np.uint8)
    k = lib().d2pc_oracle_filter_finite(_p(cloud, _u8p), n, _p(out, _u8p))
    return out[: k * 16].copy()


def serialize_pointcloud2(points: np.ndarray, seq=0, sec=0, nsec=0, frame_id="/camera_optical_frame",
                          is_dense=0) -> bytes:
    points = np.ascontiguousarray(points, dtype=np.uint8).reshape(-1)
    n = points.size // 16
    need = lib().d2pc_oracle_serialize_pointcloud2(seq, sec, nsec, frame_id.encode(), _p(points, _u8p), n, is_dense,
                                                   None, 0)
    out = np.empty(need, dtype=np.uint8)
    lib().d2pc_oracle_serialize_pointcloud2(seq, sec, nsec, frame_id.encode(), _p(points, _u8p), n, is_dense,
                                            _p(out, _u8p), need)
    return out.tobytes()


def crop_to_square(cols, rows, offset_x, offset_y, member_offset_y):
    r = np.zeros(4, dtype=np.int32)
    rc = lib().d2pc_oracle_crop_to_square(cols, rows, offset_x, offset_y, member_offset_y, _p(r, _i32p))
    return rc, tuple(int(v) for v in r)


def rotate_cw(img: np.ndarray) -> np.ndarray:
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    h, w = img.shape
    out = np.empty((w, h), dtype=np.uint8)
    lib().d2pc_oracle_rotate_cw(_p(img, _u8p), w, h, img.strides[0], _p(out, _u8p))
    return out


def grad_filter(d1, d2, s1, s2) -> int:
    return lib().d2pc_oracle_grad_filter(int(d1), int(d2), int(s1), int(s2), int(s1), int(s2))


def grad_filter_table(s1, s2) -> np.ndarray:
    """gradFilter for all (d1, d2) at fixed scores -> (256, 256) uint8, indexed [d1, d2]."""
    out = np.empty((256, 256), dtype=np.uint8)
    lib().d2pc_oracle_grad_filter_table(int(s1), int(s2), _p(out, _u8p))
    return out


def fuse_rule(mode, d1, d2, s1, s2) -> int:
    return lib().d2pc_oracle_fuse_rule(int(mode), int(d1), int(d2), int(s1), int(s2))


def fuse(d1, d2, s1, s2, offset_x=-7, offset_y=15, mode=0):
    """-> (fused (out_h,out_w) u8, combined (n,n) u8).  Raises on bad geometry."""
    arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in (d1, d2, s1, s2)]
    h, w = arrs[0].shape
    assert all(a.shape == (h, w) for a in arrs)
    side = min(w, h)
    fused = np.zeros(side * side, dtype=np.uint8)
    combined = np.zeros(side * side, dtype=np.uint8)
    dims = np.zeros(3, dtype=np.int32)
    rc = lib().d2pc_oracle_fuse(*[_p(a, _u8p) for a in arrs], w, h, w, offset_x, offset_y, mode, _p(fused, _u8p),
                                _p(combined, _u8p), _p(dims, _i32p))
    if rc != 0:
        raise ValueError("fusion geometry leaves the image")
    n, ow, oh = (int(v) for v in dims)
    return fused[: ow * oh].reshape(oh, ow).copy(), combined[: n * n].reshape(n, n).copy()


def score_preprocess(frame: np.ndarray, rect, vertical: bool) -> np.ndarray:
    """MatchingScoreCb{1,2} on a (rotated, for 2) score frame and its cropToSquare rect (x, y, n, n)."""
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    h, w = frame.shape
    r = np.array(rect, dtype=np.int32)
    out = np.empty((int(r[2]), int(r[2])), dtype=np.uint8)
    rc = lib().d2pc_oracle_score_preprocess(_p(frame, _u8p), w, h, w, _p(r, _i32p), 1 if vertical else 0, _p(out, _u8p))
    if rc != 0:
        raise ValueError("bad rectangle")
    return out


def gaussian_blur_u8(parent: np.ndarray, rect, ksize: int, sigma: float, submatrix: bool) -> np.ndarray:
    """cv::GaussianBlur(parent(rect), dst, Size(ksize, ksize), sigma) on CV_8U; rect = (x, y, w, h).  submatrix is
    Mat::isSubmatrix() of the source: True -> sepFilter2D with float32 kernels, False -> the fixed-point path."""
    parent = np.ascontiguousarray(parent, dtype=np.uint8)
    h, w = parent.shape
    r = np.array(rect, dtype=np.int32)
    out = np.empty((int(r[3]), int(r[2])), dtype=np.uint8)
    rc = lib().d2pc_oracle_gaussian_blur_u8(_p(parent, _u8p), w, h, w, _p(r, _i32p), ksize, sigma, 1 if submatrix else 0,
                                            _p(out, _u8p), out.strides[0])
    if rc != 0:
        raise ValueError(f"gaussian_blur_u8: {rc}")
    return out


def colorize_depth(gray: np.ndarray) -> np.ndarray:
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    h, w = gray.shape
    out = np.empty((h, w, 3), dtype=np.uint8)
    lib().d2pc_oracle_colorize_depth(_p(gray, _u8p), w, h, w, _p(out, _u8p))
    return out


def run_frames(frames: np.ndarray, q, mono8: bool, n_threads: int = 1, cloud: np.ndarray | None = None):
    """Batch driver for the CPU baseline; frames is (F,H,W) u8 or f32, dense."""
    frames = np.ascontiguousarray(frames)
    f, h, w = frames.shape
    n = n_points(w, h)
    if cloud is None:
        cloud = np.empty((f, n * 16), dtype=np.uint8)
    qq = _q(q)
    total = lib().d2pc_oracle_run_frames(frames.ctypes.data_as(C.c_void_p), f, w, h, frames.strides[1],
                                         1 if mono8 else 0, _p(qq, _f64p), _p(cloud, _u8p), cloud.strides[0],
                                         n_threads)
    assert total == f * n
    return cloud
