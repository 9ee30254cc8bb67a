"""ctypes view of oracle/_ref/libd2pc_ref.so: the reference's OWN translation units
(/root/reference/src/depth_map_fusion.cpp, src/disparity_to_point_cloud.cpp) compiled unmodified against the
stand-in headers of oracle/ref_stubs/ (see ref_stubs.hpp for what is reference code and what is a stand-in).

TEST INFRASTRUCTURE.  Built only where /root/reference exists (this container); the built library travels to the
GPU box with the snapshot, the reference sources do not.  Only tests/ import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libd2pc_ref.so")
REFERENCE = "/root/reference"

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int)


def build(force: bool = False) -> str | None:
    """make -C oracle ref; returns the library path, or None when the reference sources are not present and no
    prebuilt library exists."""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        import oracle
        oracle.build()
        subprocess.run(["make", "-C", _HERE, "ref"] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH if os.path.exists(LIB_PATH) else None


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not available() and build() is None:
            raise FileNotFoundError("oracle/_ref/libd2pc_ref.so is not built and /root/reference is absent")
        import oracle
        oracle.lib()  # libd2pc_oracle.so (the stand-ins' third-party arithmetic) must be loadable first
        L = C.CDLL(LIB_PATH)
        L.ref_reset.restype = None
        L.ref_set_param.argtypes = [C.c_char_p, C.c_double]
        L.ref_set_param.restype = None
        L.ref_published_clear.restype = None
        L.ref_published_count.restype = C.c_int
        L.ref_warnings.restype = C.c_int
        L.ref_published_info.argtypes = [C.c_int, C.c_char_p, C.c_char_p, _u32p]
        L.ref_published_data.argtypes = [C.c_int, _u8p, C.c_size_t]
        L.ref_published_data.restype = C.c_size_t
        L.ref_published_field.argtypes = [C.c_int, C.c_int, C.c_char_p, _u32p]
        L.ref_topics.argtypes = [C.c_int, C.c_int, C.c_char_p, _u32p]
        L.ref_fusion_create.restype = C.c_void_p
        L.ref_fusion_destroy.argtypes = [C.c_void_p]
        L.ref_fusion_destroy.restype = None
        L.ref_fusion_offsets.argtypes = [C.c_void_p, _i32p]
        L.ref_fusion_offsets.restype = None
        L.ref_fusion_callback.argtypes = [C.c_void_p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_uint32,
                                          C.c_uint32, C.c_uint32]
        L.ref_grad_filter.argtypes = [C.c_void_p] + [C.c_int] * 6
        L.ref_fuse_rule.argtypes = [C.c_void_p] + [C.c_int] * 5
        L.ref_crop_to_square.argtypes = [C.c_void_p] + [C.c_int] * 4 + [_i32p]
        L.ref_crop_mat.argtypes = [C.c_void_p] + [C.c_int] * 6 + [_i32p]
        L.ref_rotate.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _u8p]
        L.ref_rotate.restype = None
        L.ref_colorize.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _u8p]
        L.ref_colorize.restype = None
        L.ref_d2pc_create.restype = C.c_void_p
        L.ref_d2pc_destroy.argtypes = [C.c_void_p]
        L.ref_d2pc_destroy.restype = None
        L.ref_d2pc_callback.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_uint32, C.c_uint32,
                                        C.c_uint32]
        _lib = L
    return _lib


class Published:
    """One message a reference node handed to ros::Publisher::publish."""

    def __init__(self, idx: int):
        L = lib()
        topic, extra = C.create_string_buffer(64), C.create_string_buffer(64)
        info = (C.c_uint32 * 12)()
        assert L.ref_published_info(idx, topic, extra, info) == 0
        self.topic = topic.value.decode()
        self.is_cloud = bool(info[0])
        self.width, self.height = int(info[1]), int(info[2])
        self.latched, self.queue = bool(info[4]), int(info[5])
        self.seq, self.sec, self.nsec = int(info[6]), int(info[7]), int(info[8])
        n = L.ref_published_data(idx, None, 0)
        self.data = np.zeros(n, dtype=np.uint8)
        if n:
            L.ref_published_data(idx, self.data.ctypes.data_as(_u8p), n)
        if self.is_cloud:
            self.frame_id, self.point_step = extra.value.decode(), int(info[3])
            self.is_dense, self.row_step = int(info[9]), int(info[10])
            self.fields = []
            for k in range(int(info[11])):
                name, odc = C.create_string_buffer(16), (C.c_uint32 * 3)()
                L.ref_published_field(idx, k, name, odc)
                self.fields.append((name.value.decode(), int(odc[0]), int(odc[1]), int(odc[2])))
        else:
            self.encoding, self.step = extra.value.decode(), int(info[3])

    def image(self) -> np.ndarray:
        ch = self.step // self.width
        a = self.data.reshape(self.height, self.step)
        return a[:, : self.width * ch].reshape(self.height, self.width, ch)[:, :, 0] if ch == 1 else \
            a[:, : self.width * ch].reshape(self.height, self.width, ch)


def take_published() -> list[Published]:
    L = lib()
    out = [Published(i) for i in range(L.ref_published_count())]
    L.ref_published_clear()
    return out


def topics(advertised: bool):
    L = lib()
    n = L.ref_topics(1 if advertised else 0, -1, None, None)
    out = []
    for i in range(n):
        t, ql = C.create_string_buffer(64), (C.c_uint32 * 2)()
        L.ref_topics(1 if advertised else 0, i, t, ql)
        out.append((t.value.decode(), int(ql[0]), bool(ql[1])))
    return out


def _img(a):
    a = np.asarray(a)
    assert a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1
    return a


class FusionNode:
    """depth_map_fusion::DepthMapFusion, the reference's class (depth_map_fusion.hpp:63-155)."""

    def __init__(self, offset_x=None, offset_y=None):
        L = lib()
        L.ref_reset()
        if offset_x is not None:
            L.ref_set_param(b"offset_x", float(offset_x))
        if offset_y is not None:
            L.ref_set_param(b"offset_y", float(offset_y))
        self._h = C.c_void_p(L.ref_fusion_create())
        self.warnings = L.ref_warnings()
        self.advertised, self.subscribed = topics(True), topics(False)

    def close(self):
        if self._h:
            lib().ref_fusion_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def offsets(self):
        o = (C.c_int * 2)()
        lib().ref_fusion_offsets(self._h, o)
        return int(o[0]), int(o[1])

    def callback(self, which: int, img, seq=0, sec=0, nsec=0, encoding="mono8") -> tuple[int, list[Published]]:
        """which: 1 DisparityCb1, 2 DisparityCb2, 3 MatchingScoreCb1, 4 MatchingScoreCb2.
        Returns (status, messages published by this callback); status -2 / -3 = the uncaught cv_bridge / cv
        exception that would terminate the reference node."""
        a = _img(img)
        rc = lib().ref_fusion_callback(self._h, which, a.ctypes.data_as(_u8p), a.shape[1], a.shape[0], a.strides[0],
                                       encoding.encode(), seq, sec, nsec)
        return rc, take_published()

    def grad_filter(self, d1, d2, s1, s2, g1=None, g2=None):
        return lib().ref_grad_filter(self._h, d1, d2, s1, s2, s1 if g1 is None else g1, s2 if g2 is None else g2)

    def fuse_rule(self, mode, d1, d2, s1, s2):
        return lib().ref_fuse_rule(self._h, mode, d1, d2, s1, s2)

    def crop_to_square(self, cols, rows, offset_x, offset_y):
        r = (C.c_int * 4)()
        rc = lib().ref_crop_to_square(self._h, cols, rows, offset_x, offset_y, r)
        return rc, tuple(int(v) for v in r)

    def crop_mat(self, cols, rows, left, right, top, bottom):
        r = (C.c_int * 4)()
        rc = lib().ref_crop_mat(self._h, cols, rows, left, right, top, bottom, r)
        return rc, tuple(int(v) for v in r)

    def rotate(self, img):
        a = np.ascontiguousarray(_img(img))
        out = np.empty((a.shape[1], a.shape[0]), dtype=np.uint8)
        lib().ref_rotate(self._h, a.ctypes.data_as(_u8p), a.shape[1], a.shape[0], out.ctypes.data_as(_u8p))
        return out

    def colorize(self, gray):
        a = np.ascontiguousarray(_img(gray))
        out = np.empty(a.shape + (3,), dtype=np.uint8)
        lib().ref_colorize(self._h, a.ctypes.data_as(_u8p), a.shape[1], a.shape[0], out.ctypes.data_as(_u8p))
        return out


class D2pcNode:
    """d2pc::Disparity2PCloud, the reference's class (disparity_to_point_cloud.hpp:60-109)."""

    def __init__(self, **params):
        L = lib()
        L.ref_reset()
        for k, v in params.items():  # fx_, fy_, cx_, cy_, base_line_
            L.ref_set_param(k.encode(), float(v))
        self._h = C.c_void_p(L.ref_d2pc_create())
        self.advertised, self.subscribed = topics(True), topics(False)

    def close(self):
        if self._h:
            lib().ref_d2pc_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def callback(self, img, seq=0, sec=0, nsec=0, encoding="mono8"):
        a = _img(img)
        rc = lib().ref_d2pc_callback(self._h, a.ctypes.data_as(_u8p), a.shape[1], a.shape[0], a.strides[0],
                                     encoding.encode(), seq, sec, nsec)
        return rc, take_published()
