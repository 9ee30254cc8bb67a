"""GPU parity: reprojection + crop + pack (float entry) through the C ABI vs the CPU oracle.

Bar: bit-exact PointCloud2.data in EXACT mode (the default) -- count, order, the 1.0f pad word and every
XYZ bit, including +-inf and the x86 NaN pattern.  FAST mode: |err| <= 1e-5 * |z| (north_star's tolerance).
"""
import numpy as np
import pytest

import oracle
from conftest import assert_same_bits, golden
from disparity_to_point_cloud_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import disparity_to_point_cloud_b200 as d2pc
    with d2pc.Context() as c:
        yield c


def _default_q():
    return golden("q_golden.npz")["q"][0]


def test_default_q_equals_stereo_rectify(ctx):
    assert_same_bits(ctx.get_q(), _default_q(), "context Q")


@pytest.mark.parametrize("w,h", [(640, 480), (752, 480), (1280, 720), (81, 81), (665, 665), (100, 200), (333, 97),
                                 (208, 81), (209, 82), (3840, 88)])
def test_float_entry_bit_exact(ctx, w, h):
    ctx.set_q(_default_q())
    d = synth.s3_float(h, w, 1)
    got = ctx.process_f32(d)
    assert got.size == oracle.n_points(w, h) * 16
    assert_same_bits(got, oracle.disparity_cb_f32(d, _default_q()), f"{w}x{h}")
    cl = ctx.last_cloud
    assert (cl.height, cl.width, cl.point_step, cl.row_step) == (1, oracle.n_points(w, h), 16, 16 * oracle.n_points(w, h))
    assert cl.is_dense == 0 and cl.is_bigendian == 0 and cl.n_fields == 3
    assert [(f.name, f.offset, f.datatype, f.count) for f in cl.fields] == [(b"x", 0, 7, 1), (b"y", 4, 7, 1),
                                                                            (b"z", 8, 7, 1)]


def test_config1_640x480_known_answers(ctx):
    """BASELINE config 1: one 640x480 float frame; count, order, pad word, inf/NaN classes."""
    ctx.set_q(_default_q())
    d = synth.s3_float(480, 640, 0)
    d[240, :] = 0.0  # row v == 240 with d == 0 -> y is NaN (0/0)
    got = ctx.process_f32(d).view(np.uint32).reshape(-1, 4)
    assert got.shape[0] == 224000
    assert np.all(got[:, 3] == 0x3F800000)
    row240 = got[(240 - 40) * 560:(240 - 40 + 1) * 560]
    assert np.all(row240[:, 1] == 0xFFC00000)  # x86 default NaN, as the reference node publishes
    assert np.all(row240[:376 - 40, 0] == 0xFF800000) and np.all(row240[376 - 40:, 0] == 0x7F800000)
    assert np.all(row240[:, 2] == 0x7F800000)


def test_small_and_empty_frames(ctx):
    ctx.set_q(_default_q())
    for w, h in [(80, 80), (79, 300), (300, 40), (1, 1), (81, 80)]:
        got = ctx.process_f32(np.ones((h, w), dtype=np.float32))
        assert got.size == 0 and ctx.last_cloud.width == 0


def test_stress_floats_4k_width(ctx):
    """S4: arbitrary floats at 4K width -- the case where skipping the intermediate float32 cast breaks."""
    ctx.set_q(_default_q())
    d = synth.s4_stress(160, 3840, 4)
    assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, _default_q()), "S4 4K")


def test_special_values(ctx):
    ctx.set_q(_default_q())
    d = synth.s4_stress(120, 256, 9)
    flat = d.reshape(-1)
    flat[::11] = np.inf
    flat[1::13] = -np.inf
    flat[2::17] = np.nan
    flat[3::19] = -0.0
    flat[4::23] = np.float32(1e-42)   # denormal
    flat[5::29] = np.float32(-3.5)
    flat[6::31] = np.float32(3e38)
    flat[7::37] = np.uint32(0xFFC12345).view(np.float32)  # NaN with payload and sign
    flat[8::41] = np.uint32(0x7F812345).view(np.float32)  # signalling NaN
    assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, _default_q()), "special values")


@pytest.mark.parametrize("params", [(714.24, 713.5, 376.0, 240.0, 0.043), (500.5, 480.25, 300.0, 200.0, 0.2),
                                    (1400.0, 1390.0, 640.5, 360.25, 0.12)])
def test_other_intrinsics(ctx, params):
    import disparity_to_point_cloud_b200 as d2pc
    q = d2pc.q_from_intrinsics(*params)
    assert_same_bits(q, oracle.q_from_intrinsics(*params), "Q")
    ctx.set_q(q)
    d = synth.s3_float(300, 500, 2)
    assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, q), "intrinsics")
    ctx.set_q(_default_q())


def test_generic_q_matrix(ctx):
    g = golden("reproject_golden.npz")
    q = g["q_generic"]
    ctx.set_q(q)
    try:
        d = synth.s4_stress(200, 400, 3)
        assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, q), "generic Q")
        # rectified-form Q with non-zero q33 and negative q32
        q2 = _default_q().copy()
        q2[3, 3] = 0.37
        q2[3, 2] = -7.25
        q2[2, 3] = -0.0
        ctx.set_q(q2)
        assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, q2), "rectified, q33 != 0")
        # tiny negative q03 -> X rounds to -0.0f at u == 0 ... exercise the neg-zero numerator guard
        q3 = _default_q().copy()
        q3[0, 3] = -1e-300
        q3[1, 3] = -1e-300
        ctx.set_q(q3)
        ctx.set_tuning("border", 0)
        d3 = synth.s3_float(40, 64, 5)
        assert_same_bits(ctx.process_f32(d3), oracle.crop_pack(oracle.reproject_image_to_3d(d3, q3), 0), "neg zero")
    finally:
        ctx.set_tuning("border", 40)
        ctx.set_q(_default_q())


def test_golden_fixture_cv2(ctx):
    """Directly against cv2.reprojectImageTo3D output (committed fixture), not only via the oracle."""
    g = golden("reproject_golden.npz")
    for dn, qn in [("d_s3", "ref"), ("d_s3", "gen"), ("d_s4", "ref"), ("d_s4", "gen")]:
        q = g["q_ref"] if qn == "ref" else g["q_generic"]
        ctx.set_q(q)
        ctx.set_tuning("border", 0)
        try:
            got = ctx.process_f32(g[dn]).view(np.float32).reshape(-1, 4)
        finally:
            ctx.set_tuning("border", 40)
        want = g[f"xyz_{dn[2:]}_{qn}"].reshape(-1, 3)
        assert_same_bits(got[:, :3].copy(), want, f"{dn}/{qn}")
    ctx.set_q(_default_q())


def test_generic_q_zero_disparities_stay_straight_line(ctx):
    """The generic-Q arithmetic answers W == +-0 without the exact fall-back: +-inf by sign(numerator) ^ sign(W),
    NaN (x86 pattern) where the numerator is zero as well.  Frames that are half zeros, with -0.0f, inf, NaN and
    denormal disparities mixed in, under Q matrices that make W +0, -0 and non-zero for d == 0."""
    rng = np.random.default_rng(77)
    d = synth.s4_stress(240, 520, 9)
    d[rng.random(d.shape) < 0.5] = 0.0
    d[rng.random(d.shape) < 0.02] = -0.0
    for v in (np.inf, -np.inf, np.nan, 1e-42, -1e-42, 3.0e38):
        d[rng.integers(0, d.shape[0], 40), rng.integers(0, d.shape[1], 40)] = np.float32(v)
    qs = {"default (W = +0)": _default_q().copy()}
    qneg = _default_q().copy()       # every term of W is -0 for d == +0
    qneg[3, 0] = -0.0
    qneg[3, 1] = -0.0
    qneg[3, 2] = -abs(qneg[3, 2])
    qneg[3, 3] = -0.0
    qs["W = -0"] = qneg
    qgen = golden("reproject_golden.npz")["q_generic"].copy()
    qs["generic"] = qgen
    qz = qgen.copy()                 # W == 0 for d == 0 under a full matrix; X numerator 0 on the column u == 100
    qz[3, 0] = 0.0
    qz[3, 1] = 0.0
    qz[3, 3] = 0.0
    qz[0] = [1.0, 0.0, 0.0, -100.0]
    qs["generic, q33 = 0, zero numerators"] = qz
    ctx.set_tuning("force_generic", 1)
    try:
        for name, q in qs.items():
            ctx.set_q(q)
            assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, q), name)
    finally:
        ctx.set_tuning("force_generic", 0)
        ctx.set_q(_default_q())


@pytest.mark.parametrize("q32,q33", [(1.0 / 0.09, 0.0), (-1.0 / 0.09, -0.0), (-7.25, 0.37)])
def test_integral_principal_point_zero_numerators(ctx, q32, q33):
    """A calibration whose principal point falls on a pixel: X is exactly 0 on the column u == cx and Y on the row
    v == cy.  The guarded-multiply paths keep those pixels straight-line (0 * r is the IEEE quotient, sign included;
    0 / 0 where the disparity is zero as well is x86's NaN); the Markstein variant and q33 != 0 send them to the
    exact function.  Every code path, CROP and CROP_FINITE, against the oracle."""
    import disparity_to_point_cloud_b200 as d2pc
    q = np.array([[1, 0, 0, -376.0], [0, 1, 0, -240.0], [0, 0, 0, 713.5], [0, 0, q32, q33]], dtype=np.float64)
    rng = np.random.default_rng(5)
    d = synth.s4_stress(480, 752, 11)
    d[rng.random(d.shape) < 0.2] = 0.0
    d[rng.random(d.shape) < 0.02] = -0.0
    d[240, ::3] = 0.0                     # zeros on the Y == 0 row and the X == 0 column: 0 / 0
    d[::5, 376] = 0.0
    d[240, 376] = 2.5
    for v in (np.inf, -np.inf, np.nan, 1e-42, 3.0e38, -4.0):
        d[rng.integers(0, 480, 30), rng.integers(0, 752, 30)] = np.float32(v)
        d[240, rng.integers(40, 712, 3)] = np.float32(v)
        d[rng.integers(40, 440, 3), 376] = np.float32(v)
    want = oracle.disparity_cb_f32(d, q)
    ctx.set_q(q)
    try:
        # zero_numer: 0 picks the zero-numerator kernel variant for this Q, -1 forces the ordinary one (that column
        # then takes the exact function), 1 is what a Q without such a column never sees
        for knobs in ({}, {"zero_numer": -1}, {"zero_numer": 1, "force_scalar": 1}, {"force_scalar": 1},
                      {"exact_variant": 1}, {"force_generic": 1}, {"rows_per_unit": 3}):
            for k, v in knobs.items():
                ctx.set_tuning(k, v)
            try:
                assert_same_bits(ctx.process_f32(d), want, f"CROP {knobs}")
                ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE)
                for park in (0, 1):
                    ctx.set_tuning("force_park", park)
                    assert_same_bits(ctx.process_f32(d), oracle.filter_finite(want), f"CROP_FINITE park={park} {knobs}")
            finally:
                ctx.set_filter_mode(d2pc.FILTER_CROP)
                ctx.set_tuning("force_park", 0)
                for k in knobs:
                    ctx.set_tuning(k, 0)
        # the mono8 callback (bytes / 8) through the same Q
        img = synth.s2_scene(480, 752, 4)
        assert_same_bits(ctx.process_mono8(img), oracle.disparity_cb_mono8(img, q), "mono8")
        # and the zero-numerator variant under the default Q (zero row v == 240, no zero column)
        ctx.set_q(_default_q())
        ctx.set_tuning("zero_numer", 1)
        assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, _default_q()), "default Q, zero_numer=1")
        ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE)
        assert_same_bits(ctx.process_f32(d), oracle.filter_finite(oracle.disparity_cb_f32(d, _default_q())),
                         "default Q, zero_numer=1, CROP_FINITE")
    finally:
        ctx.set_filter_mode(d2pc.FILTER_CROP)
        ctx.set_tuning("zero_numer", 0)
        ctx.set_q(_default_q())


@pytest.mark.parametrize("knob", ["force_scalar", "force_generic", "exact_variant"])
def test_alternate_code_paths_agree(ctx, knob):
    ctx.set_q(_default_q())
    d = synth.s4_stress(300, 1000, 6)
    want = oracle.disparity_cb_f32(d, _default_q())
    ctx.set_tuning(knob, 1)
    try:
        assert_same_bits(ctx.process_f32(d), want, knob)
    finally:
        ctx.set_tuning(knob, 0)


def test_markstein_variant_special_values_and_4k(ctx):
    """exact_variant 1 (Markstein quotients) against the oracle on the inputs that stress the guards."""
    ctx.set_q(_default_q())
    ctx.set_tuning("exact_variant", 1)
    try:
        d = synth.s4_stress(120, 3840, 11)
        flat = d.reshape(-1)
        for i, v in enumerate([np.inf, -np.inf, np.nan, -0.0, 1e-42, -3.5, 3e38, 1e-38, 2e19, 1e25]):
            flat[i::31] = np.float32(v)
        assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, _default_q()), "markstein")
    finally:
        ctx.set_tuning("exact_variant", 0)


def test_guarded_multiply_large_and_tiny_disparities(ctx):
    """default variant: disparities beyond d_hi = 2^64/|q32| and near the float-denormal result range."""
    ctx.set_q(_default_q())
    rng = np.random.default_rng(12)
    d = np.exp(rng.uniform(np.log(1e-38), np.log(3e38), size=(200, 512))).astype(np.float32)
    d[::3, ::5] *= -1
    assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, _default_q()), "log-uniform disparities")


def test_unaligned_and_strided_input(ctx):
    ctx.set_q(_default_q())
    big = synth.s4_stress(200, 700, 7)
    view = big[3:190, 5:650]          # step != width*4, base not 16-byte aligned
    assert view.strides[0] == 700 * 4
    assert_same_bits(ctx.process_f32(view), oracle.disparity_cb_f32(view, _default_q()), "strided view")


@pytest.mark.parametrize("rows", [2, 4, 8, 32])
def test_rows_per_unit_knob(ctx, rows):
    ctx.set_q(_default_q())
    d = synth.s3_float(333, 777, 8)
    ctx.set_tuning("rows_per_unit", rows)
    try:
        assert_same_bits(ctx.process_f32(d), oracle.disparity_cb_f32(d, _default_q()), f"rows={rows}")
    finally:
        ctx.set_tuning("rows_per_unit", 0)


def test_fast_mode_within_tolerance(ctx):
    import disparity_to_point_cloud_b200 as d2pc
    ctx.set_q(_default_q())
    d = synth.s3_float(720, 1280, 3)
    want = oracle.disparity_cb_f32(d, _default_q()).view(np.float32).reshape(-1, 4)
    ctx.set_arith_mode(d2pc.ARITH_FAST)
    try:
        got = ctx.process_f32(d).view(np.float32).reshape(-1, 4)
    finally:
        ctx.set_arith_mode(d2pc.ARITH_EXACT)
    fin = np.isfinite(want[:, :3]).all(axis=1)
    assert np.array_equal(np.isfinite(got[:, :3]).all(axis=1), fin)          # same validity classes
    assert np.array_equal(got[~fin].view(np.uint32) & 0x7F800000, want[~fin].view(np.uint32) & 0x7F800000)
    err = np.abs(got[fin, :3] - want[fin, :3]).max(axis=1)
    tol = 1e-5 * np.abs(want[fin, 2])   # north_star: max abs error <= 1e-5 relative to depth
    assert np.all(err <= tol), float((err / np.abs(want[fin, 2])).max())
    assert np.all(got[:, 3] == 1.0)


# ---- CROP_FINITE (extension): order-preserving compaction ---------------------------------------
@pytest.mark.parametrize("park", [0, 1])
@pytest.mark.parametrize("w,h,kind", [(640, 480, "s2"), (1280, 720, "s2"), (333, 97, "s1"), (81, 81, "zeros"),
                                      (400, 300, "nozeros"), (2000, 90, "s2"), (700, 300, "special"),
                                      (3840, 200, "s2"), (208, 1200, "s2"), (85, 83, "s1"), (5000, 100, "s2")])
def test_crop_finite_compaction(ctx, w, h, kind, park):
    """compact_variant 0: band kernel (default for the rectified Q; TMA row loads when the rows are aligned floats);
    1: park-then-compact, the fallback every other Q takes.  Both single-pass with decoupled look-back."""
    import disparity_to_point_cloud_b200 as d2pc
    ctx.set_q(_default_q())
    ctx.set_tuning("force_park", park)
    if kind == "s2":
        d = synth.s2_scene(h, w, 2).astype(np.float32) * np.float32(0.125)
    elif kind == "s1":
        d = synth.s3_float(h, w, 2)
    elif kind == "zeros":
        d = np.zeros((h, w), dtype=np.float32)
    elif kind == "special":
        # the sliver the classifier can not decide from d alone: denormals and |d| < ~2^-117, plus inf/NaN/-0
        d = synth.s4_stress(h, w, 2)
        flat = d.reshape(-1)
        for i, v in enumerate([1e-45, 1e-40, 1.2e-38, 1e-37, 3e-36, 1e-35, 5e-34, -1e-37, -0.0, np.inf, -np.inf,
                               np.nan, 3e38, -2.5]):
            flat[i::29] = np.float32(v)
    else:
        d = np.full((h, w), 3.5, dtype=np.float32)
    want = oracle.filter_finite(oracle.disparity_cb_f32(d, _default_q()))
    ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE)
    try:
        got = ctx.process_f32(d)
        assert ctx.last_cloud.width == want.size // 16 and ctx.last_cloud.is_dense == 1
        assert_same_bits(got, want, f"compaction {kind}")
        # run it again: the epoch-tagged descriptors must not leak state between launches
        assert_same_bits(ctx.process_f32(d), want, f"compaction {kind} (2nd launch)")
    finally:
        ctx.set_filter_mode(d2pc.FILTER_CROP)
        ctx.set_tuning("force_park", 0)


def test_crop_finite_generic_q_and_fast(ctx):
    import disparity_to_point_cloud_b200 as d2pc
    q = golden("reproject_golden.npz")["q_generic"]
    d = synth.s4_stress(200, 640, 5)
    ctx.set_q(q)
    ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE)
    try:
        assert_same_bits(ctx.process_f32(d), oracle.filter_finite(oracle.disparity_cb_f32(d, q)), "generic Q")
        q2 = _default_q().copy()
        q2[3, 3] = 0.25  # rectified form with q33 != 0
        ctx.set_q(q2)
        assert_same_bits(ctx.process_f32(d), oracle.filter_finite(oracle.disparity_cb_f32(d, q2)), "q33 != 0")
    finally:
        ctx.set_filter_mode(d2pc.FILTER_CROP)
        ctx.set_q(_default_q())


# ---- device-resident batch entry ------------------------------------------------------------------
def test_device_batch_entry(ctx):
    import torch
    ctx.set_q(_default_q())
    f, h, w = 5, 200, 336
    frames = np.stack([synth.s4_stress(h, w, 20 + i) for i in range(f)])
    n = oracle.n_points(w, h)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.zeros((f, n * 16 + 64), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, h * w * 4, d_out.data_ptr(), n * 16 + 64)
    ctx.sync()
    got = d_out.cpu().numpy()
    for i in range(f):
        assert_same_bits(got[i, :n * 16], oracle.disparity_cb_f32(frames[i], _default_q()), f"frame {i}")
        assert not got[i, n * 16:].any()


def test_device_batch_compaction(ctx):
    import torch
    import disparity_to_point_cloud_b200 as d2pc
    ctx.set_q(_default_q())
    f, h, w = 4, 180, 400
    frames = np.stack([synth.s2_scene(h, w, 30 + i).astype(np.float32) * np.float32(0.125) for i in range(f)])
    frames[2] = 0.0
    n = oracle.n_points(w, h)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.zeros((f, n * 16), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE)
    try:
        for variant in (0, 1):  # band kernel, park kernel
            ctx.set_tuning("compact_variant", variant)
            d_out.zero_()
            d_cnt.zero_()
            for _ in range(2):
                ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, h * w * 4, d_out.data_ptr(), n * 16,
                                         d_cnt.data_ptr())
                ctx.sync()
            got, cnt = d_out.cpu().numpy(), d_cnt.cpu().numpy()
            for i in range(f):
                want = oracle.filter_finite(oracle.disparity_cb_f32(frames[i], _default_q()))
                assert cnt[i] == want.size // 16
                assert_same_bits(got[i, :want.size], want, f"frame {i} (variant {variant})")
                assert not got[i, want.size:].any(), "a compaction kernel wrote past the frame's last survivor"
    finally:
        ctx.set_filter_mode(d2pc.FILTER_CROP)
        ctx.set_tuning("compact_variant", 0)


def test_full_size_properties_4k(ctx):
    """BASELINE config 4 frame size: size-independent properties + oracle on one whole 4K frame."""
    ctx.set_q(_default_q())
    h, w = 2160, 3840
    d = synth.s3_float(h, w, 1024)
    got = ctx.process_f32(d)
    pts = got.view(np.float32).reshape(h - 80, w - 80, 4)
    assert got.size == 7820800 * 16
    assert np.all(pts[..., 3].view(np.uint32) == 0x3F800000)
    fin = np.isfinite(pts[..., 2])
    assert np.array_equal(fin, d[40:-40, 40:-40] != 0)
    # row-major order: x strictly increases along a row wherever depth is equal; here check monotone u via x/z
    ratio = np.where(fin, pts[..., 0] / pts[..., 2], np.nan)
    col = np.nanmedian(ratio, axis=0)
    assert np.all(np.diff(col) > 0)
    assert_same_bits(got, oracle.disparity_cb_f32(d, _default_q()), "4K frame")


@pytest.mark.parametrize("variant", [0, 1])
def test_full_size_compaction_4k(ctx, variant):
    """CROP_FINITE at BASELINE config 4's frame size (2080-band look-back chains): count, order and bits against
    the filtered oracle, plus the order-preservation property."""
    import disparity_to_point_cloud_b200 as d2pc
    ctx.set_q(_default_q())
    h, w = 2160, 3840
    d = synth.s3_float(h, w, 77)
    d[1000:1040, 500:2500] = 0.0  # a hole spanning many units
    d[240, :] = 0.0               # the row whose Y numerator is zero
    want = oracle.filter_finite(oracle.disparity_cb_f32(d, _default_q()))
    ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE)
    ctx.set_tuning("compact_variant", variant)
    try:
        got = ctx.process_f32(d)
    finally:
        ctx.set_filter_mode(d2pc.FILTER_CROP)
        ctx.set_tuning("compact_variant", 0)
    assert got.size == want.size == 16 * int(np.count_nonzero(d[40:-40, 40:-40]))
    assert_same_bits(got, want, f"4K compaction (variant {variant})")


# ---- BASELINE.json configs at their full sizes --------------------------------------------------------------
def test_config3_full_batch_1280x720x64(ctx):
    """configs[2]: a resident batch of 64 1280x720 float frames through one launch, every frame bit-exact against
    the oracle, in CROP and in CROP_FINITE mode (count, order and bits)."""
    import torch
    import disparity_to_point_cloud_b200 as d2pc
    ctx.set_q(_default_q())
    f, h, w = 64, 720, 1280
    base = synth.s3_float(h, w, 3)
    frames = np.stack([np.roll(base, 17 * i, axis=1) for i in range(f)])
    frames[5, 100:140, 300:900] = 0.0
    n = oracle.n_points(w, h)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.zeros((f, n * 16), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(f, dtype=torch.int32, device="cuda")
    ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, h * w * 4, d_out.data_ptr(), n * 16)
    ctx.sync()
    got = d_out.cpu().numpy()
    want = [oracle.disparity_cb_f32(frames[i], _default_q()) for i in range(f)]
    for i in range(f):
        assert_same_bits(got[i], want[i], f"frame {i}")
    ctx.set_filter_mode(d2pc.FILTER_CROP_FINITE)
    try:
        d_out.zero_()
        ctx.reproject_f32_device(d_in.data_ptr(), f, w, h, w * 4, h * w * 4, d_out.data_ptr(), n * 16, d_cnt.data_ptr())
        ctx.sync()
    finally:
        ctx.set_filter_mode(d2pc.FILTER_CROP)
    got, cnt = d_out.cpu().numpy(), d_cnt.cpu().numpy()
    for i in range(f):
        wf = oracle.filter_finite(want[i])
        assert cnt[i] == wf.size // 16
        assert_same_bits(got[i, :wf.size], wf, f"compacted frame {i}")


def test_config4_ring_of_4k_frames_1024(ctx):
    """configs[3] as bench.py runs it: 1024 3840x2160 frames streamed through a ring of 16 device slots (64
    launches).  The 16 distinct frames are checked bit for bit against the oracle after the last launch, and the
    ring is idempotent: every pass writes the same bytes (checksum of checksums over the passes)."""
    import torch
    ctx.set_q(_default_q())
    ring, h, w = 16, 2160, 3840
    base = synth.s3_float(h, w, 1000)
    frames = np.stack([np.roll(base, 131 * i + 7, axis=1) for i in range(ring)])
    n = oracle.n_points(w, h)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.zeros((ring, n * 16), dtype=torch.uint8, device="cuda")
    sums = []
    for launch in range(1024 // ring):
        ctx.reproject_f32_device(d_in.data_ptr(), ring, w, h, w * 4, h * w * 4, d_out.data_ptr(), n * 16)
        if launch in (0, 31, 63):
            ctx.sync()
            sums.append(int(d_out.view(torch.int32).sum(dtype=torch.int64).item()))
    ctx.sync()
    assert sums[0] == sums[1] == sums[2]
    got = d_out.cpu().numpy()
    for i in range(ring):
        assert_same_bits(got[i], oracle.disparity_cb_f32(frames[i], _default_q()), f"ring slot {i}")
