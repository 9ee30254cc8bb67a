"""GPU: seeded differential fuzz of the DisparityCb entry points against the oracle.

Every case draws a frame size, an entry (mono8 / float), a Q (the reference's default, stereoRectify of random
intrinsics incl. an integral principal point, q33 != 0, or a generic non-rectified matrix), a data class (uniform,
scene, constant, zero-heavy; float frames also carry inf / NaN / -0 / denormals / negatives / huge values), a row
layout (dense or padded rows, pageable / pinned / registered), an output destination (library-owned, caller-owned
page-locked or pageable, a pipeline slot), the filter mode and a set of tuning knobs (fused or two-launch callback,
direct output, zero-numerator variant, Markstein quotients, scalar loads, median strip, compaction kernel) -- and
must reproduce the oracle's bytes.  D2PC_FUZZ_CASES sets the number of cases (default 300; a 2000-case run is
recorded in profiles/r2_fuzz.txt)."""
import os

import numpy as np
import pytest

import oracle
from disparity_to_point_cloud_b200 import synth

pytestmark = pytest.mark.gpu

N_CASES = int(os.environ.get("D2PC_FUZZ_CASES", "300"))
SEED0 = int(os.environ.get("D2PC_FUZZ_SEED", "20260"))


def _draw_q(rng, d2pc):
    kind = rng.choice(["default", "rectify", "rectify_integral", "q33", "generic"], p=[0.35, 0.25, 0.2, 0.1, 0.1])
    if kind == "default":
        return kind, oracle.q_from_intrinsics()
    if kind in ("rectify", "rectify_integral"):
        fx, fy = rng.uniform(300, 1500, 2)
        cx, cy = rng.uniform(100, 600), rng.uniform(100, 400)
        b = rng.choice([-1, 1]) * rng.uniform(0.02, 0.5)
        q = oracle.q_from_intrinsics(fx, fy, cx, cy, b).reshape(4, 4).copy()
        if kind == "rectify_integral":  # X (and sometimes Y) is exactly zero on an image column (row)
            q[0, 3] = -float(rng.integers(60, 200))
            if rng.random() < 0.5:
                q[1, 3] = -float(rng.integers(60, 200))
        return kind, q
    if kind == "q33":
        q = oracle.q_from_intrinsics().reshape(4, 4).copy()
        q[3, 3] = rng.choice([-1, 1]) * rng.uniform(0.01, 30.0)
        return kind, q
    q = rng.normal(0, 1, (4, 4))
    q[rng.random((4, 4)) < 0.3] = 0.0
    q[3, 2] = rng.uniform(0.5, 20.0)
    return kind, q


def _draw_u8(rng, h, w):
    kind = rng.choice(["uniform", "scene", "constant", "zeros", "steps"])
    if kind == "uniform":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "scene":
        return synth.s2_scene(h, w, int(rng.integers(0, 1 << 16)))
    if kind == "constant":
        return np.full((h, w), int(rng.integers(0, 256)), np.uint8)
    if kind == "zeros":
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img[rng.random((h, w)) < rng.uniform(0.2, 0.98)] = 0
        return img
    img = np.repeat(np.repeat(rng.integers(0, 256, ((h + 15) // 16, (w + 15) // 16), dtype=np.uint8), 16, 0), 16, 1)
    return np.ascontiguousarray(img[:h, :w])


def _draw_f32(rng, h, w):
    d = _draw_u8(rng, h, w).astype(np.float32) * np.float32(0.125)
    if rng.random() < 0.4:
        d = rng.uniform(-4.0, 64.0, (h, w)).astype(np.float32) if rng.random() < 0.5 else synth.s4_stress(h, w, int(rng.integers(1 << 16)))
    if rng.random() < 0.5:  # specials
        specials = np.concatenate([
            np.array([np.inf, -np.inf, np.nan, -0.0, 0.0, 1e-45, -1e-40, 1.17549435e-38, 3.4e38, -3.4e38, 1e-30,
                      2.0 ** -120, -2.0 ** -126], np.float32),
            np.array([0x7f800001, 0xffc12345, 0x7fa00000], np.uint32).view(np.float32)])  # signalling / payload NaNs
        m = rng.random((h, w)) < rng.uniform(0.001, 0.05)
        d = d.copy()
        d[m] = rng.choice(specials, int(m.sum()))
    return d


@pytest.fixture(scope="module")
def env():
    import disparity_to_point_cloud_b200 as d2pc
    ctx = d2pc.Context(n_slots=3)
    cap = 600 * 420 * 16 + 256
    pin = d2pc.PinnedArray((cap,), np.uint8)
    reg = d2pc.RegisteredArray(np.empty(cap, np.uint8))
    pin_in = d2pc.PinnedArray((600 * 4 * 420 + 4096,), np.uint8)
    yield d2pc, ctx, pin, reg, pin_in
    pin.free()
    reg.free()
    pin_in.free()
    ctx.close()


KNOBS = {"fuse_median": [0, 0, -1], "direct_out": [0, 0, -1, 1], "zero_numer": [0, 0, 1, -1], "exact_variant": [0, 0, 1],
         "force_scalar": [0, 0, 0, 1], "median_strip": [0, 0, 2, 8, 64], "compact_variant": [0, 0, 0, 1],
         "force_generic": [0, 0, 0, 0, 1], "rows_per_unit": [0, 0, 2, 4, 16], "prefetch_dist": [0, 0, -1, 64]}


def _one_case(case, env):
    d2pc, ctx, pin, reg, pin_in = env
    rng = np.random.default_rng(SEED0 + case)
    w = int(rng.integers(81, 600)) if rng.random() < 0.93 else int(rng.integers(1, 90))
    h = int(rng.integers(81, 420)) if rng.random() < 0.93 else int(rng.integers(1, 90))
    mono = bool(rng.random() < 0.5)
    qkind, q = _draw_q(rng, d2pc)
    finite = bool(rng.random() < 0.4)
    frame = _draw_u8(rng, h, w) if mono else _draw_f32(rng, h, w)
    knobs = {k: int(rng.choice(v)) for k, v in KNOBS.items()}
    layout = rng.choice(["dense", "padded", "pinned"])
    dest = rng.choice(["library", "pinned", "registered", "pageable", "slot"])
    desc = dict(case=case, w=w, h=h, mono=mono, q=qkind, finite=finite, layout=str(layout), dest=str(dest), knobs=knobs)

    want = oracle.disparity_cb_mono8(frame, q) if mono else oracle.disparity_cb_f32(frame, q)
    if finite:
        want = oracle.filter_finite(want)

    src = frame
    if layout == "padded":
        pad = int(rng.integers(1, 40)) * (1 if mono else 1)
        buf = np.zeros((h, w + pad), frame.dtype)
        buf[:, :w] = frame
        src = buf[:, :w]
    elif layout == "pinned":
        view = pin_in.array[: frame.nbytes].view(frame.dtype).reshape(h, w)
        view[:] = frame
        src = view
    ctx.set_q(q)
    ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE if finite else d2pc.FILTER_CROP)
    for k, v in knobs.items():
        ctx.set_tuning(k, v)
    try:
        if dest == "library":
            got = ctx.process_mono8(src) if mono else ctx.process_f32(src)
        elif dest == "slot":
            slot = int(rng.integers(0, 3))
            ctx.submit(slot, src)
            got = ctx.wait(slot)
        else:
            dst = {"pinned": pin.array, "registered": reg.array, "pageable": np.empty(want.size + 64, np.uint8)}[str(dest)]
            off = int(rng.choice([0, 0, 16, 4]))
            dst = dst[off: off + want.size + 48]
            dst[:] = 0x5A
            got = ctx.process_into(src, dst)
            assert (dst[want.size:] == 0x5A).all(), ("wrote past the cloud", desc)
        assert got.size == want.size, (desc, got.size, want.size)
        if got.tobytes() != np.ascontiguousarray(want).tobytes():
            g = np.frombuffer(got.tobytes(), np.uint32)
            o = np.frombuffer(np.ascontiguousarray(want).tobytes(), np.uint32)
            bad = np.nonzero(g != o)[0]
            raise AssertionError(f"{desc}: {bad.size} words differ, first at word {bad[0]}: {g[bad[0]]:#x} vs {o[bad[0]]:#x}")
    finally:
        for k in knobs:
            ctx.set_tuning(k, 0)
        ctx.set_filter_mode(d2pc.FILTER_CROP)
    return desc


def test_fuzz_against_oracle(env):
    seen = {}
    for case in range(N_CASES):
        d = _one_case(case, env)
        key = (d["mono"], d["q"], d["finite"])
        seen[key] = seen.get(key, 0) + 1
    if N_CASES >= 80:
        assert len(seen) >= 12, seen  # the draw really spreads over entries x Q forms x filter modes
    print(f"fuzz: {N_CASES} cases from seed {SEED0}, {len(seen)} (entry, Q, filter) classes")


# ---------------------------------------------------------------------------------------------------------------
# depth_map_fusion: merge (rotate / crop / rule / median 3 / trim), geometry errors, score chain, fused -> cloud
# ---------------------------------------------------------------------------------------------------------------
N_FUSION = int(os.environ.get("D2PC_FUZZ_FUSION_CASES", "60"))


@pytest.fixture(scope="module")
def fctx():
    import disparity_to_point_cloud_b200 as d2pc
    with d2pc.Context(offset_x=-7, offset_y=15) as c:
        yield d2pc, c


def _fusion_case(case, fctx):
    d2pc, ctx = fctx
    rng = np.random.default_rng(SEED0 + 500000 + case)
    w, h = int(rng.integers(90, 520)), int(rng.integers(90, 420))
    ox, oy = int(rng.integers(-40, 41)), int(rng.integers(-40, 41))
    if rng.random() < 0.08:
        ox = int(rng.integers(-300, 300))  # often leaves the image: cv::Mat::operator() would throw in the reference
    rule = int(rng.choice([0, 0, 0, 1, 2, 3, 4, 5, 6, 7]))
    d1, d2, s1, s2 = (_draw_u8(rng, h, w) for _ in range(4))
    if rng.random() < 0.5:  # correlated maps: d2 is d1 seen by the rotated camera, plus noise
        d2 = np.ascontiguousarray(np.rot90(np.resize(d1, (w, h)), 1))
        d2 = np.clip(d2.astype(np.int32) + rng.integers(-6, 7, d2.shape), 0, 255).astype(np.uint8)
    desc = dict(case=case, w=w, h=h, ox=ox, oy=oy, rule=rule)
    ctx.set_tuning("offset_x", ox)
    ctx.set_tuning("offset_y", oy)
    ctx.set_tuning("fuse_rule", rule)
    try:
        try:
            of, oc = oracle.fuse(d1, d2, s1, s2, ox, oy, mode=rule)
        except ValueError:
            with pytest.raises(d2pc.D2pcError) as e:
                ctx.fuse(d1, d2, s1, s2)
            assert e.value.status == -7, desc
            return "geometry"
        fused, combined = ctx.fuse(d1, d2, s1, s2)
        assert fused.shape == of.shape and fused.tobytes() == of.tobytes(), desc
        assert combined.tobytes() == oc.tobytes(), desc
        # the score chain of both callbacks for this geometry
        rc1, r1 = oracle.crop_to_square(w, h, ox, oy, oy)
        rc2, r2 = oracle.crop_to_square(h, w, -ox, -oy, oy)
        if rc1 == 0 and rc2 == 0 and r1[2] > 0:
            p1, p2 = ctx.preprocess_score(s1, 1), ctx.preprocess_score(s2, 2)
            assert p1.tobytes() == oracle.score_preprocess(s1, r1, False).tobytes(), ("score 1", desc)
            assert p2.tobytes() == oracle.score_preprocess(oracle.rotate_cw(s2), r2, True).tobytes(), ("score 2", desc)
        # fused map -> DisparityCb without leaving the device (config 5), default rule only to bound the run time
        if rule == 0 and of.shape[0] > 0 and of.shape[1] > 0:
            q = ctx.get_q()
            got = ctx.fuse_then_process(d1, d2, s1, s2)
            assert got.tobytes() == oracle.disparity_cb_mono8(of, q).tobytes(), ("fuse_then_process", desc)
        return "ok"
    finally:
        ctx.set_tuning("offset_x", -7)
        ctx.set_tuning("offset_y", 15)
        ctx.set_tuning("fuse_rule", 0)


def test_fuzz_fusion_against_oracle(fctx):
    outcomes = {}
    for case in range(N_FUSION):
        r = _fusion_case(case, fctx)
        outcomes[r] = outcomes.get(r, 0) + 1
    assert outcomes.get("ok", 0) > 0
    print(f"fusion fuzz: {N_FUSION} cases from seed {SEED0}: {outcomes}")
