"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
its host-only entry points work, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import disparity_to_point_cloud_b200 as d2pc
from conftest import ROOT, assert_same_bits, golden


@pytest.fixture(scope="module")
def lib():
    from disparity_to_point_cloud_b200 import build
    build.build()
    return d2pc.lib()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "d2pc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(d2pc_[a-z0-9_]+)\s*\(", text))
    names -= {"d2pc_cloud_sink"}
    return sorted(names)


def test_every_declared_symbol_is_exported(lib):
    raw = C.CDLL(d2pc.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/d2pc_b200.h but not exported"
    bound = {s[0] for s in d2pc.SYMBOLS}
    assert set(names) == bound, set(names) ^ bound


def test_library_is_sm100a_cuda(lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", d2pc.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_config_defaults_are_the_reference_constants(lib):
    c = d2pc.default_config()
    assert (c.fx, c.fy, c.cx, c.cy, c.baseline) == (714.24, 713.5, 376.0, 240.0, 0.09)
    assert (c.rect_width, c.rect_height, c.border, c.median_ksize) == (752, 480, 40, 11)
    assert c.disparity_scale == 0.125
    assert c.frame_id == b"/camera_optical_frame"
    assert (c.fuse_crop_left, c.fuse_crop_right, c.fuse_crop_top, c.fuse_crop_bottom) == (0, 40, 30, 10)
    assert c.fuse_median_ksize == 3 and c.filter_mode == 0 and c.arith_mode == 0
    assert c.struct_size == C.sizeof(d2pc.Config)


def test_q_from_intrinsics_matches_stereo_rectify(lib):
    g = golden("q_golden.npz")
    for p, q in zip(g["params"], g["q"]):
        assert_same_bits(d2pc.q_from_intrinsics(*p), q, f"Q {p}")
    with pytest.raises(d2pc.D2pcError):
        d2pc.q_from_intrinsics(baseline=0.0)


def test_strerror_and_abi(lib):
    assert lib.d2pc_abi_version() == 1
    assert lib.d2pc_strerror(0) == b"ok"
    assert b"B200" in lib.d2pc_strerror(-5)


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert d2pc.device_count() == 0
    with pytest.raises(d2pc.D2pcError) as e:
        d2pc.Context()
    assert e.value.status == -5


def test_product_never_imports_the_oracle():
    import subprocess
    import sys
    code = ("import sys; import disparity_to_point_cloud_b200 as m; m.lib(); "
            "assert not any(k == 'oracle' or k.startswith('oracle.') for k in sys.modules), 'oracle imported'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    pkg = os.path.join(ROOT, "disparity_to_point_cloud_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|d2pc_oracle|libd2pc_oracle|oracle/")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_tuning_keys_documented_in_the_header_match_the_library():
    """Every integer knob d2pc_set_tuning accepts is listed in include/d2pc_b200.h, and nothing is listed that the
    library would refuse (no GPU needed: both sides are read from the sources)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "disparity_to_point_cloud_b200", "csrc", "capi.cu")).read()
    body = src[src.index("int d2pc_set_tuning("):]
    body = body[: body.index("\n}\n")]
    accepted = set(re.findall(r'k == "([a-z0-9_]+)"', body))
    hdr = open(os.path.join(root, "include", "d2pc_b200.h")).read()
    doc = hdr[hdr.index("Tuning / test hook"): hdr.index("int d2pc_set_tuning(")]
    documented = set()
    for m in re.findall(r'"([a-z0-9_|]+)"', doc):
        if "|" in m:   # "fuse_crop_left|right|top|bottom"
            head, *rest = m.split("|")
            stem = head[: head.rindex("_") + 1]
            documented.add(head)
            documented.update(stem + r for r in rest)
        else:
            documented.add(m)
    accepted.discard("force_park")  # an alias of compact_variant kept for old scripts
    assert accepted == documented, (sorted(accepted - documented), sorted(documented - accepted))
