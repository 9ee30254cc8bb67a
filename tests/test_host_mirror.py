"""The C++ host mirror (include/d2pc_b200/nodes.hpp) driven through tools/d2pc_offline:
Disparity2PCloud::DisparityCb and DepthMapFusion's callbacks behind the reference's topic names and the
launch-file remaps, ROS1 wire bytes in and out."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT, golden
from disparity_to_point_cloud_b200 import synth


@pytest.fixture(scope="module")
def harness():
    from disparity_to_point_cloud_b200 import build
    build.build()
    return build.build_harness()


def image_msg(img, sec, nsec, seq=0, frame_id="camera"):
    """sensor_msgs/Image on the ROS1 wire; (H, W) is mono8, (H, W, 3) is what the reference labels rgb8."""
    h, w = img.shape[:2]
    fid = frame_id.encode()
    enc, step = (b"mono8", w) if img.ndim == 2 else (b"rgb8", 3 * w)
    return (struct.pack("<III", seq, sec, nsec) + struct.pack("<I", len(fid)) + fid + struct.pack("<II", h, w) +
            struct.pack("<I", len(enc)) + enc + struct.pack("<BI", 0, step) + struct.pack("<I", step * h) + img.tobytes())


def test_wire_selftest(harness, tmp_path):
    out = tmp_path / "wire.bin"
    subprocess.run([harness, "wire", str(out)], check=True)
    assert out.read_bytes() == oracle.serialize_pointcloud2(np.arange(32, dtype=np.uint8), seq=7, sec=11, nsec=13)


def test_launch_files_keep_the_reference_remaps():
    d2p = open(os.path.join(ROOT, "launch", "d2pcloud.launch")).read()
    assert 'from="/disparity" to="/throttled_depth_map"' in d2p and 'from="/point_cloud" to="/omi_cam/point_cloud"' in d2p
    assert 'type="disparity_to_point_cloud_node"' in d2p
    fus = open(os.path.join(ROOT, "launch", "depth_map_fusion.launch")).read()
    for a, b in [("/matching_score_1", "cam_0"), ("/disparity_1", "cam_1"), ("/matching_score_2", "cam_2"),
                 ("/disparity_2", "cam_3")]:
        assert f'from="{a}" to="/uvc_camera/{b}/image_raw"' in fus
    assert 'name="offset_x" value="-7"' in fus and 'name="offset_y" value="15"' in fus


@pytest.mark.gpu
def test_node1_disparity_cb_through_launch_remaps(harness, tmp_path):
    q = golden("q_golden.npz")["q"][0]
    img = synth.s2_scene(480, 752, 21)
    raw, out = tmp_path / "in.raw", tmp_path / "cloud.bin"
    raw.write_bytes(img.tobytes())
    subprocess.run([harness, "node1", os.path.join(ROOT, "launch", "d2pcloud.launch"), "752", "480", str(raw), str(out),
                    "1700000000", "250"], check=True)
    want = oracle.serialize_pointcloud2(oracle.disparity_cb_mono8(img, q), seq=0, sec=1700000000, nsec=250)
    assert out.read_bytes() == want


@pytest.mark.gpu
def test_fusion_node_then_node1_config5(harness, tmp_path):
    q = golden("q_golden.npz")["q"][0]
    h, w = 720, 1280
    rng = np.random.default_rng(9)
    d1, d2 = synth.s2_scene(h, w, 31), synth.s2_scene(h, w, 32)
    s1 = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    s2 = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    paths = []
    for name, a in (("d1", d1), ("d2", d2), ("s1", s1), ("s2", s2)):
        p = tmp_path / f"{name}.raw"
        p.write_bytes(a.tobytes())
        paths.append(str(p))
    fused_bin, cloud_bin, dbg = tmp_path / "fused.bin", tmp_path / "cloud.bin", tmp_path / "debug"
    dbg.mkdir()
    subprocess.run([harness, "fusion", os.path.join(ROOT, "launch", "depth_map_fusion.launch"), str(w), str(h)] + paths +
                   [str(fused_bin), str(cloud_bin), str(dbg)], check=True)
    # MatchingScoreCb1/2 preprocess the scores (src/depth_map_fusion.cpp:64-99) before they are cached
    _, r1 = oracle.crop_to_square(w, h, -7, 15, 15)
    _, r2 = oracle.crop_to_square(h, w, 7, -15, 15)
    p1 = oracle.score_preprocess(s1, r1, False)
    p2 = oracle.score_preprocess(oracle.rotate_cw(s2), r2, True)
    c1 = np.zeros((h, w), np.uint8)
    c1[r1[1]:r1[1] + r1[2], r1[0]:r1[0] + r1[2]] = p1
    rot = np.zeros((w, h), np.uint8)
    rot[r2[1]:r2[1] + r2[2], r2[0]:r2[0] + r2[2]] = p2
    fused, combined = oracle.fuse(d1, d2, c1, np.ascontiguousarray(np.rot90(rot, 1)), -7, 15)
    assert fused.shape == (665, 665)
    # the six debug topics (depth_map_fusion.hpp:106-117; publishWithColor, src/depth_map_fusion.cpp:275-302), each
    # with the header of the message whose callback published it
    crop1 = d1[r1[1]:r1[1] + r1[2], r1[0]:r1[0] + r1[2]]
    crop2 = oracle.rotate_cw(d2)[r2[1]:r2[1] + r2[2], r2[0]:r2[0] + r2[2]]
    want_dbg = {"cropped_depth_1": image_msg(oracle.colorize_depth(crop1), 1, 0),
                "cropped_depth_2": image_msg(oracle.colorize_depth(crop2), 2, 500),
                "cropped_score_1": image_msg(p1, 1, 0), "cropped_score_2": image_msg(p2, 1, 0),
                "combined_score": image_msg(np.ascontiguousarray(combined), 2, 500),
                "gradient": image_msg(oracle.colorize_depth(fused), 2, 500)}
    for name, want_bytes in want_dbg.items():
        assert (dbg / f"{name}.bin").read_bytes() == want_bytes, name
    assert fused_bin.read_bytes() == image_msg(fused, 2, 500)          # header of message 2 (:134-135)
    want = oracle.serialize_pointcloud2(oracle.disparity_cb_mono8(fused, q), seq=0, sec=2, nsec=500)
    assert cloud_bin.read_bytes() == want


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(300, 424), (480, 752)])
def test_fusion_node_state_between_callbacks(harness, tmp_path, h, w):
    """The state DepthMapFusion keeps between callbacks (src/depth_map_fusion.cpp:77, :113, :118-121):
    DisparityCb2 before the caches are full publishes no fused map; two DisparityCb2 in a row -- the second merges
    with score 1 = min(score 1, score 2) left behind by the first; a fresh MatchingScoreCb1 replaces it.  Every
    message on all seven topics, in order, against oracle/nodes.py, which tests/test_ref_compiled.py pins to the
    reference's own compiled class."""
    from oracle import nodes
    rng = np.random.default_rng(h + w)
    seq = [2, 1, 3, 2, 4, 2, 2, 2, 3, 2, 4, 4, 2, 1, 2, 3]
    o = nodes.FusionNodeOracle(-7, 15)   # launch/depth_map_fusion.launch
    lines, want = [], []
    for step, which in enumerate(seq):
        img = _score_like(rng, h, w) if which in (3, 4) else synth.s2_scene(h, w, 100 + step)
        p = tmp_path / f"m{step}.raw"
        p.write_bytes(img.tobytes())
        lines.append(f"{which} {p} {1000 + step} {17 * step}")
        for topic, enc, arr, hdr in o.callback(which, img, (step, 1000 + step, 17 * step)):
            want.append((topic, image_msg(np.ascontiguousarray(arr), hdr[1], hdr[2], seq=hdr[0])))
    script, out = tmp_path / "script.txt", tmp_path / "log.bin"
    script.write_text("\n".join(lines) + "\n")
    subprocess.run([harness, "fusion-seq", os.path.join(ROOT, "launch", "depth_map_fusion.launch"), str(w), str(h),
                    str(script), str(out)], check=True)
    blob, got, i = out.read_bytes(), [], 0
    while i < len(blob):
        tl = struct.unpack_from("<I", blob, i)[0]
        topic = blob[i + 4:i + 4 + tl].decode()
        ml = struct.unpack_from("<I", blob, i + 4 + tl)[0]
        got.append((topic, blob[i + 8 + tl:i + 8 + tl + ml]))
        i += 8 + tl + ml
    assert [t for t, _ in got] == [t for t, _ in want]
    assert sum(t == "/fused_depth_map" for t, _ in got) == 6
    for k, ((t, g_bytes), (_, w_bytes)) in enumerate(zip(got, want)):
        assert g_bytes == w_bytes, (k, t)


def _score_like(rng, h, w):
    a = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    a[rng.integers(0, h, 6), :] = 250
    a[:, rng.integers(0, w, 6)] = 3
    return a


# ---- the host side of the drop-in under AddressSanitizer / UBSan --------------------------
ASAN_ENV = dict(os.environ, ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0:abort_on_error=1",
                UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")


@pytest.fixture(scope="module")
def harness_asan():
    from disparity_to_point_cloud_b200 import build
    build.build()
    return build.build_harness(sanitize=True)


def test_wire_selftest_under_sanitizers(harness_asan, tmp_path):
    out = tmp_path / "wire.bin"
    r = subprocess.run([harness_asan, "wire", str(out)], env=ASAN_ENV, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert out.read_bytes() == oracle.serialize_pointcloud2(np.arange(32, dtype=np.uint8), seq=7, sec=11, nsec=13)


@pytest.mark.gpu
def test_node_classes_under_sanitizers(harness, harness_asan, tmp_path):
    """Both node classes through their whole life (launch-file parsing, subscriptions, callbacks, the registered
    message buffer the kernel writes into, publication, serialisation, destruction) under ASan + UBSan: clean exit
    and the same bytes as the plain harness -- node 1 on a 752x480 frame, the fusion node on a 16-callback
    sequence."""
    img = synth.s2_scene(480, 752, 44)
    raw = tmp_path / "in.raw"
    raw.write_bytes(img.tobytes())
    outs = {}
    for name, exe in (("plain", harness), ("asan", harness_asan)):
        out = tmp_path / f"cloud_{name}.bin"
        r = subprocess.run([exe, "node1", os.path.join(ROOT, "launch", "d2pcloud.launch"), "752", "480", str(raw), str(out),
                            "1700000001", "7"], env=ASAN_ENV, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr[-3000:])
        outs[name] = out.read_bytes()
    assert outs["plain"] == outs["asan"]
    q = golden("q_golden.npz")["q"][0]
    assert outs["asan"] == oracle.serialize_pointcloud2(oracle.disparity_cb_mono8(img, q), seq=0, sec=1700000001, nsec=7)

    rng = np.random.default_rng(5)
    h, w, lines = 300, 424, []
    for step, which in enumerate([3, 4, 1, 2, 2, 3, 2, 1, 4, 2, 2, 1, 3, 4, 2, 2]):
        a = _score_like(rng, h, w) if which in (3, 4) else synth.s2_scene(h, w, 300 + step)
        p = tmp_path / f"s{step}.raw"
        p.write_bytes(a.tobytes())
        lines.append(f"{which} {p} {2000 + step} {step}")
    script = tmp_path / "script.txt"
    script.write_text("\n".join(lines) + "\n")
    logs = {}
    for name, exe in (("plain", harness), ("asan", harness_asan)):
        out = tmp_path / f"log_{name}.bin"
        r = subprocess.run([exe, "fusion-seq", os.path.join(ROOT, "launch", "depth_map_fusion.launch"), str(w), str(h),
                            str(script), str(out)], env=ASAN_ENV, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr[-3000:])
        logs[name] = out.read_bytes()
    assert len(logs["asan"]) > 0 and logs["plain"] == logs["asan"]
