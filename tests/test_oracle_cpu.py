"""CPU: the C++ oracle against an independent numpy re-derivation (tests/ref_numpy.py), Pillow (sanity
bound for the Lanczos stage) and its own committed golden vectors. No GPU, no product code."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import ref_numpy as R
from tests.fixtures import CASES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", sorted(CASES))
def test_stats_match_numpy(case):
    dn = CASES[case](97, 131)
    db, mask = O.process_scalar_data_inplace(dn.astype(np.float32))
    st, hist = O.compute_histogram_stats(db, mask)
    dbn, maskn = R.db_and_mask(dn.astype(np.float32))
    assert np.array_equal(mask.astype(bool), maskn)
    assert np.allclose(db, dbn, rtol=1e-14, atol=0)
    ref = R.stats(dbn, maskn)
    if ref is None:
        assert st.valid_count == 0
        return
    assert st.valid_count == ref["n"]
    assert st.min_db == pytest.approx(ref["min"], rel=1e-14)
    assert st.max_db == pytest.approx(ref["max"], rel=1e-14)
    assert st.mean_db == pytest.approx(ref["mean"], rel=1e-11)
    assert st.std_db == pytest.approx(ref["std"], rel=1e-9, abs=1e-12)
    if "hist" in ref:
        assert int(np.abs(hist.astype(np.int64) - ref["hist"]).sum()) <= 2  # libm ulp at a bin edge at most
    for k in ("p01", "p02", "p05", "p10", "p25", "p75", "p90", "p95", "p98", "p99"):
        assert getattr(st, k) == pytest.approx(ref[k], rel=1e-9, abs=1e-9), k
    assert st.median_db == pytest.approx(ref["median"], rel=1e-9, abs=1e-9)


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("strategy", range(7))
@pytest.mark.parametrize("u8", [True, False])
def test_autoscale_matches_numpy(case, strategy, u8):
    dn = CASES[case](83, 117)
    v = dn.astype(np.float32)
    po = O.process_scalar_data_pipeline(v, O.U8 if u8 else O.U16, strategy, want_db=False)
    got = po.u8 if u8 else po.u16
    ref, _ = R.autoscale(v, u8, strategy)
    diff = got.astype(np.int64) - ref.astype(np.int64)
    # two independent evaluations may differ by 1 LSB where a libm ulp crosses a quantisation edge
    assert np.abs(diff).max() <= 1
    assert (diff != 0).mean() <= 1e-3


def test_branch_coverage_of_fixtures():
    """The fixtures really exercise every Standard / Adaptive branch (gamma identifies the branch)."""
    want_std = {"low_contrast": 1.1, "homogeneous": 1.0, "high_dynamic": 0.9, "speckle": 1.0}
    for case, gamma in want_std.items():
        dn = CASES[case](203, 317)
        po = O.process_scalar_data_pipeline(dn.astype(np.float32), O.U8, O.STANDARD, want_db=False)
        assert po.stats.gamma == gamma, case
    dn = CASES["homogeneous"](203, 317)
    st = O.process_scalar_data_pipeline(dn.astype(np.float32), O.U8, O.STANDARD, want_db=False).stats
    assert st.max_db - st.min_db >= 15 and st.p75 - st.p25 < 5
    want_adp = {"skew_pos": 0.9, "skew_neg": 1.1, "heavy_tail": 0.8, "speckle": 1.0}
    for case, gamma in want_adp.items():
        dn = CASES[case](203, 317)
        po = O.process_scalar_data_pipeline(dn.astype(np.float32), O.U8, O.ADAPTIVE, want_db=False)
        assert po.stats.gamma == gamma, case
    dn = CASES["narrow_output"](64, 64)
    db, mask = O.process_scalar_data_inplace(dn.astype(np.float32))
    q, _ = O.autoscale_db_image_advanced(db, mask, O.U8, O.DEFAULT)
    assert q.max() < 255 or q.min() > 0  # scale_u16_to_u8 really rescales


def test_zero_valid_and_all_equal():
    z = np.zeros((40, 50), np.float32)
    for s in range(7):
        po = O.process_scalar_data_pipeline(z, O.U8, s, want_db=False)
        assert po.stats.valid_count == 0 and not po.u8.any()
    e = np.full((40, 50), 77, np.float32)
    po = O.process_scalar_data_pipeline(e, O.U16, O.ROBUST, want_db=False)
    assert po.stats.p01 == po.stats.min_db == po.stats.p99
    neg = -np.ones((8, 8), np.float32)
    nan = np.full((8, 8), np.nan, np.float32)
    for a in (neg, nan):
        db, mask = O.process_scalar_data_inplace(a)
        assert not mask.any() and np.all(db == -100.0)


@pytest.mark.parametrize("shape", [(64, 64), (97, 131), (203, 317), (16, 300)])
def test_clahe_matches_numpy(shape):
    rng = np.random.default_rng(3)
    norm = rng.random(shape)
    norm[rng.random(shape) < 0.3] = 0.25  # spike -> clipping + remainder redistribution
    mask = rng.random(shape) > 0.1
    r0, r1 = shape[0] // 8, 2 * (shape[0] // 8)
    mask[r0:r1, : shape[1] // 8] = False  # one fully invalid tile
    out, cdfs = O.clahe_equalize_normalized(norm, mask.astype(np.uint8))
    ref, cdfn = R.clahe(norm, mask)
    assert np.allclose(cdfs, cdfn, rtol=0, atol=1e-15)
    assert np.allclose(out, ref, rtol=0, atol=1e-14)
    # top/left half tiles extrapolate (dy in [-0.5,0)): values may leave [0,1] before the final clamp
    assert out.min() >= -0.5 and out.max() <= 1.5


def test_scale_u16_to_u8_and_ops():
    rng = np.random.default_rng(4)
    d = rng.integers(3, 250, (50, 60)).astype(np.uint16)
    assert np.array_equal(O.scale_u16_to_u8(d), R.scale_u16_to_u8(d))
    flat = np.full((4, 4), 9, np.uint16)
    assert np.array_equal(O.scale_u16_to_u8(flat), np.zeros((4, 4), np.uint8))  # scale = 1.0, x - min = 0
    a = rng.gamma(2.0, 50.0, (30, 40)).astype(np.float32)
    b = rng.gamma(2.0, 20.0, (30, 40)).astype(np.float32)
    b[0, :5] = 0
    a[1, :5] = -b[1, :5]
    assert np.array_equal(O.pol_op(O.OP_SUM, a, b), a + b)
    assert np.array_equal(O.pol_op(O.OP_DIFF, a, b), a - b)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(np.abs(b) > np.float32(1e-10), a / b, np.float32(0))
        nd = np.where(np.abs(a + b) > np.float32(1e-10), (a - b) / (a + b), np.float32(0))
    assert np.array_equal(O.pol_op(O.OP_RATIO, a, b), ratio)
    assert np.array_equal(O.pol_op(O.OP_LOGRATIO, a, b), ratio)
    assert np.array_equal(O.pol_op(O.OP_NDIFF, a, b), nd)


@pytest.mark.parametrize("cols,rows,target", [(640, 480, 200), (480, 640, 200), (1000, 333, 333), (90, 100, 500),
                                              (25000, 16000, 2048), (1000, 1000, 128), (7, 5000, 100)])
def test_resize_dims(cols, rows, target):
    assert O.calculate_resize_dimensions(cols, rows, target) == R.resize_dims(cols, rows, target)
    if (cols, rows, target) == (25000, 16000, 2048):
        assert O.calculate_resize_dimensions(cols, rows, target) == (2048, 1311)
        assert O.resize_output_dims(cols, rows, target, True) == (2048, 2048)


@pytest.mark.parametrize("shape,tc,tr", [((120, 200), 50, 30), ((200, 120), 31, 52), ((64, 777), 64, 5), ((300, 301), 299, 298)])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_lanczos_matches_numpy(shape, tc, tr, dtype):
    rng = np.random.default_rng(6)
    img = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
    fn = O.resize_u8_image if dtype == np.uint8 else O.resize_u16_image
    assert np.array_equal(fn(img, tc, tr), R.resize_lanczos3(img, tc, tr))


def test_lanczos_vs_pillow_sanity():
    """Pillow's LANCZOS uses the same window formulation with a fixed 22-bit precision and rounding of the
    horizontal pass; expect +-1 LSB on a small fraction of samples. Sanity bound only (parity unpinned)."""
    from PIL import Image
    rng = np.random.default_rng(8)
    base = rng.integers(0, 256, (64, 96)).astype(np.uint8)
    img = np.kron(base, np.ones((8, 8), np.uint8))  # smooth-ish 512 x 768
    got = O.resize_u8_image(img, 200, 133)
    pil = np.asarray(Image.fromarray(img, "L").resize((200, 133), Image.LANCZOS))
    diff = np.abs(got.astype(int) - pil.astype(int))
    # measured here: 99.5 % identical, a handful of samples off by 2 (different fixed-point precision / bounds rounding)
    assert diff.max() <= 2
    assert (diff != 0).mean() < 0.02
    assert (diff > 1).mean() < 1e-3


def test_padding_and_meta():
    rng = np.random.default_rng(10)
    img = rng.integers(0, 256, (31, 50)).astype(np.uint8)
    assert np.array_equal(O.add_padding_to_square(img, O.U8), R.pad_square(img))
    img16 = rng.integers(0, 65536, (50, 31)).astype(np.uint16)
    assert np.array_equal(O.add_padding_to_square(img16, O.U16), R.pad_square(img16))
    out, meta = O.resize_image_data_with_meta(img, 25, O.U8, True)
    assert (meta.cols, meta.rows) == (25, 25) and out.shape == (25, 25)
    assert meta.scale_x == 25 / 50 and meta.scale_y == 16 / 31 and (meta.pad_left, meta.pad_top) == (0, 4)
    out, meta = O.resize_image_data_with_meta(img, 50, O.U8, False)  # long side == target: skipped (resize.rs:115-116)
    assert np.array_equal(out, img) and meta.scale_x == 1.0
    out, meta = O.resize_image_data_with_meta(img, 500, O.U8, False)  # no upscaling
    assert np.array_equal(out, img)


def test_synrgb_matches_numpy():
    rng = np.random.default_rng(12)
    b1 = rng.integers(0, 256, (120, 90)).astype(np.uint8)
    b2 = rng.integers(0, 256, (120, 90)).astype(np.uint8)
    b2[:10] = 0
    b1[:30] = 0
    d = O.create_synthetic_rgb(b1, b2).astype(int) - R.synrgb_default(b1, b2).astype(int)
    assert np.abs(d).max() <= 1 and (d != 0).mean() < 1e-3  # numpy powf vs libm powf: 1 ulp at a rounding edge
    d = O.create_synthetic_rgb_suppressed(b1, b2).astype(int) - R.synrgb_suppressed(b1, b2).astype(int)
    assert np.abs(d).max() <= 1 and (d != 0).mean() < 1e-3
    # quirks: b2 == 0 -> B = 0; both under the floor -> black; floor cushion capped at 40
    assert not O.create_synthetic_rgb(b1, b2)[:10, :, 2].any()
    hi = np.full((20, 20), 200, np.uint8)
    rgb = O.create_synthetic_rgb_suppressed(hi, hi)  # p05 = 200 -> floor capped at 40
    assert rgb[0, 0, 0] > 0
    assert np.array_equal(O.create_synthetic_rgb_by_mode_and_strategy(2, O.CLAHE, b1, b2), O.create_synthetic_rgb_suppressed(b1, b2))
    assert np.array_equal(O.create_synthetic_rgb_by_mode_and_strategy(3, O.ROBUST, b1, b2), O.create_synthetic_rgb(b1, b2))


def test_golden_vectors():
    """Committed outputs of the oracle (tests/golden/make_golden.py): pins the checker itself."""
    g = np.load(os.path.join(GOLDEN, "golden_small.npz"))
    vv, vh = g["vv"], g["vh"]
    for s in range(7):
        for bd, key in ((O.U8, "u8"), (O.U16, "u16")):
            po = O.process_scalar_data_pipeline(vv.astype(np.float32), bd, s, want_db=False)
            assert np.array_equal(po.u8 if bd == O.U8 else po.u16, g[f"autoscale_{s}_{key}"])
        rgb, _ = O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), s, 96, True)
        assert np.array_equal(rgb, g[f"synrgb_{s}"])
    assert np.array_equal(O.resize_u8_image(g["img8"], 57, 41), g["resize8"])
    assert np.array_equal(O.resize_u16_image(g["img16"], 41, 57), g["resize16"])


def _wide_scene():
    import hashlib
    from sarpro_b200.synth import synth_pair
    g = np.load(os.path.join(GOLDEN, "golden_wide.npz"))
    vv, vh = synth_pair(640, 2048, scene=5, point_targets=1e-4)
    assert np.array_equal(np.frombuffer(hashlib.sha256(vv.tobytes() + vh.tobytes()).digest(), np.uint8), g["input_sha256"])
    return g, vv, vh


def test_golden_wide_scene():
    """The wide committed scene (the shape the tensor-core pass B takes): the oracle still reproduces it."""
    g, vv, vh = _wide_scene()
    rgb, _ = O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), O.CLAHE, 256, True)
    assert np.array_equal(rgb, g["synrgb_clahe"])


# ---- downsample-on-read oracle (restated GDAL RasterIO resampling, parity unpinned) -------------------------------------
def test_read_dims_for_target_matches_reader_arithmetic():
    """sentinel1.rs:1083-1102: long side -> target, short side rounded, never enlarged; Average from a reduction of 4."""
    assert O.read_dims_for_target(25000, 16000, 2048) == (2048, 1311, O.RESAMPLE_AVERAGE)
    assert O.read_dims_for_target(16000, 25000, 2048) == (1311, 2048, O.RESAMPLE_AVERAGE)
    assert O.read_dims_for_target(4096, 4096, 2048) == (2048, 2048, O.RESAMPLE_LANCZOS)   # reduction 2
    assert O.read_dims_for_target(8192, 100, 2048) == (2048, 25, O.RESAMPLE_AVERAGE)     # reduction exactly 4
    assert O.read_dims_for_target(1000, 800, 4000) == (1000, 800, O.RESAMPLE_LANCZOS)    # target above the long side: no upscale
    assert O.read_dims_for_target(30000, 3, 100) == (100, 1, O.RESAMPLE_AVERAGE)         # short side at least 1


def test_read_average_known_answers():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 4000, (160, 240)).astype(np.uint16)
    # integer reduction: plain block means (every weight is 1)
    got = O.read_band_resampled(a, 24, 16, O.RESAMPLE_AVERAGE)
    ref = a.reshape(16, 10, 24, 10).astype(np.float64).mean(axis=(1, 3)).astype(np.float32)
    assert np.array_equal(got, ref)
    # fractional reduction 2.5 along x, one row: pixel d covers [2.5 d, 2.5 d + 2.5)
    row = np.arange(10, dtype=np.uint16)[None, :] * 100
    got = O.read_band_resampled(row, 4, 1, O.RESAMPLE_AVERAGE)[0]
    want = [(0 + 100 + 0.5 * 200) / 2.5, (0.5 * 200 + 300 + 400) / 2.5, (500 + 600 + 0.5 * 700) / 2.5, (0.5 * 700 + 800 + 900) / 2.5]
    assert np.allclose(got, np.float32(want), rtol=0, atol=1e-4)
    # a constant raster stays constant under both resamplers, u16 and f32 sources agree
    c = np.full((333, 517), 1234, np.uint16)
    for alg in (O.RESAMPLE_AVERAGE, O.RESAMPLE_LANCZOS):
        for oc, orr in ((100, 64), (200, 129), (517, 333)):
            r16 = O.read_band_resampled(c, oc, orr, alg)
            assert np.allclose(r16, 1234.0, rtol=1e-6), (alg, oc, orr)
            assert np.array_equal(r16, O.read_band_resampled(c.astype(np.float32), oc, orr, alg))


def test_read_lanczos_is_a_normalised_low_pass():
    rng = np.random.default_rng(4)
    a = rng.gamma(4.0, 50.0, (300, 420)).astype(np.float32)
    out = O.read_band_resampled(a, 210, 150, O.RESAMPLE_LANCZOS)       # reduction 2
    assert out.shape == (150, 210)
    assert abs(float(out.mean()) - float(a.mean())) < 0.01 * float(a.mean())
    assert float(out.std()) < float(a.std())
    ident = O.read_band_resampled(a, 420, 300, O.RESAMPLE_LANCZOS)     # same shape: the kernel collapses to the sample itself
    assert np.allclose(ident, a, rtol=1e-5, atol=1e-3)


# ---- adversarial inputs for the restated fast_image_resize Lanczos (parity unpinned: these are the properties any faithful
#      implementation of the crate's fixed-point scheme must have, plus two independent opinions: the numpy re-derivation and Pillow)
@pytest.mark.parametrize("in_w,out_w", [(16, 3), (97, 96), (1000, 7), (2048, 300), (25000, 2048), (4099, 512), (9, 8)])
def test_lanczos_adversarial_rows(in_w, out_w):
    from PIL import Image
    rows = 6

    def both(img, dtype):
        fn = O.resize_u8_image if dtype == np.uint8 else O.resize_u16_image
        got = fn(img.astype(dtype), out_w, rows)
        assert np.array_equal(got, R.resize_lanczos3(img.astype(dtype), out_w, rows)), "oracle != numpy re-derivation"
        return got

    for dtype, top in ((np.uint8, 255), (np.uint16, 65535)):
        # constants survive exactly (the quantised taps of every window sum close enough to 2^p), including the clamp ends
        for c in (0, 1, top // 2, top - 1, top):
            got = both(np.full((rows, in_w), c, np.int64), dtype)
            assert (got == c).all(), (dtype.__name__, c, np.unique(got))
        # impulses (negative lobes must clamp at 0, never wrap), at the edges and in the middle
        for pos in (0, 1, in_w // 2, in_w - 2, in_w - 1):
            img = np.zeros((rows, in_w), np.int64)
            img[:, pos] = top
            got = both(img, dtype)
            assert got.max() <= top and got.min() >= 0
            assert got[0].sum() > 0 or in_w > 40 * out_w   # the impulse shows up unless its tap weight rounds to nothing
            assert (got == got[0]).all()                   # rows are independent and identical
        # alternating columns / a step: the worst case for overshoot; the clamp must hold both ends
        alt = np.tile(np.array([0, top], np.int64), in_w // 2 + 1)[:in_w][None, :].repeat(rows, 0)
        step = np.where(np.arange(in_w) < in_w // 2, 0, top)[None, :].repeat(rows, 0)
        for img in (alt, step):
            got = both(img, dtype)
            assert got.min() >= 0 and got.max() <= top
        got = both(step, dtype)[0]
        assert got[0] == 0 and got[-1] == top                # far from the edge the step is flat
        if dtype == np.uint8:
            # third opinion on the same inputs: Pillow (same windows and weights, fixed 22-bit precision)
            for img in (alt, step):
                a = O.resize_u8_image(img.astype(np.uint8), out_w, rows)
                p = np.asarray(Image.fromarray(img.astype(np.uint8), "L").resize((out_w, rows), Image.LANCZOS))
                assert np.abs(a.astype(int) - p.astype(int)).max() <= 2
