"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Integer outputs must be bit-exact; percentiles / window f64 bit-exact; mean/std within 1e-9 (the
reference's serial Welford rounding is order-dependent and is not reproducible from histograms)."""
import os

import numpy as np
import pytest

import sarpro_b200 as S
from oracle import pyoracle as O
from tests.fixtures import CASES

pytestmark = pytest.mark.gpu

EXACT_STATS = ["valid_count", "min_db", "max_db", "median_db", "p01", "p02", "p05", "p10", "p25", "p75", "p90",
               "p95", "p98", "p99", "low_clip", "high_clip", "gamma"]


def check_stats(st, so):
    for k in EXACT_STATS:
        assert getattr(st, k) == getattr(so, k), k
    assert abs(st.mean_db - so.mean_db) <= 1e-9 * max(1.0, abs(so.mean_db))
    assert abs(st.std_db - so.std_db) <= 1e-9 * max(1.0, abs(so.std_db))


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("strategy", range(7))
@pytest.mark.parametrize("bit_depth", [S.U8, S.U16])
def test_process_scalar_data_pipeline(ctx, case, strategy, bit_depth):
    dn = CASES[case](203, 317)  # neither divisible by 8
    v = dn.astype(np.float32)
    po = O.process_scalar_data_pipeline(v, bit_depth, strategy, want_db=False)
    for arr in (v, dn):  # f32 API boundary and raw DN entry
        u8, u16, st = ctx.process_scalar_data_pipeline(arr, bit_depth, strategy)
        got, ref = (u8, po.u8) if bit_depth == S.U8 else (u16, po.u16)
        assert np.array_equal(got, ref)
        check_stats(st, po.stats)


@pytest.mark.parametrize("strategy", [S.STANDARD, S.ROBUST, S.CLAHE, S.TAMED])
@pytest.mark.parametrize("shape", [(1024, 1536), (777, 1201), (1500, 640)])
def test_pipeline_single_resized(ctx, strategy, shape):
    dn = CASES["speckle"](*shape)
    v = dn.astype(np.float32)
    for bd in (S.U8, S.U16):
        for target, pad in ((256, True), (300, False), (None, True)):
            ref, meta = O.pipeline_single(v, O.TIFF, bd, strategy, target, pad)
            img = ctx.process_single(dn, S.TIFF, bd, strategy, target, pad)
            got = img.gray if bd == S.U8 else img.gray16
            assert got.shape == ref.shape
            assert np.array_equal(got, ref), (strategy, bd, target, pad, int((got != ref).sum()))
            assert (img.scale_x, img.scale_y, img.pad_left, img.pad_top) == (meta.scale_x, meta.scale_y, meta.pad_left, meta.pad_top)


@pytest.mark.parametrize("strategy", range(7))
@pytest.mark.parametrize("tamed_step", [True, False])
def test_pipeline_synrgb(ctx, strategy, tamed_step):
    vv = CASES["speckle"](900, 1400)
    vh = CASES["speckle_vh"](900, 1400)
    ref, meta = O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), strategy, 512, True,
                                       tamed_band_step=tamed_step)
    img = ctx.process_synrgb_jpeg(vv, vh, strategy, 512, True, tamed_band_step=tamed_step)
    assert img.rgb.shape == ref.shape
    assert np.array_equal(img.rgb, ref), int((img.rgb != ref).sum())


@pytest.mark.parametrize("bit_depth", [S.U8, S.U16])
def test_pipeline_multiband_tiff(ctx, bit_depth):
    vv = CASES["speckle"](640, 1000)
    vh = CASES["speckle_vh"](640, 1000)
    r1, r2, meta = O.pipeline_multiband_tiff(vv.astype(np.float32), vh.astype(np.float32), bit_depth, S.CLAHE, 400, True)
    img = ctx.process_multiband_tiff(vv, vh, bit_depth, S.CLAHE, 400, True)
    g1, g2 = (img.gray, img.gray_band2) if bit_depth == S.U8 else (img.gray16, img.gray16_band2)
    assert np.array_equal(g1, r1) and np.array_equal(g2, r2)


@pytest.mark.parametrize("shape,target", [((480, 640), 200), ((640, 480), 200), ((333, 1000), 333), ((100, 90), 500),
                                          ((1000, 1000), 128), ((50, 2000), 64)])
@pytest.mark.parametrize("bit_depth", [S.U8, S.U16])
@pytest.mark.parametrize("pad", [False, True])
def test_resize_image_data_with_meta(ctx, shape, target, bit_depth, pad):
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256 if bit_depth == S.U8 else 65536, shape).astype(np.uint8 if bit_depth == S.U8 else np.uint16)
    ref, mo = O.resize_image_data_with_meta(data, target, bit_depth, pad)
    got, mg = ctx.resize_image_data_with_meta(data, target, bit_depth, pad)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref), int((got != ref).sum())
    for f in ("cols", "rows", "scale_x", "scale_y", "pad_left", "pad_top"):
        assert getattr(mg, f) == getattr(mo, f)


def test_u16_required_error(ctx):
    with pytest.raises(S.SarproError) as e:
        ctx.resize_image_data_with_meta(None, 100, S.U16, False)
    assert "U16 data required for U16 bit depth" in str(e.value)


@pytest.mark.parametrize("strategy", [S.STANDARD, S.CLAHE, S.TAMED])
def test_synthetic_rgb(ctx, strategy):
    rng = np.random.default_rng(7)
    b1 = rng.integers(0, 256, (300, 411)).astype(np.uint8)
    b2 = rng.integers(0, 256, (300, 411)).astype(np.uint8)
    b1[:40] = 0
    b2[:35] = 0
    b2[100:120] = 0
    ref = O.create_synthetic_rgb_by_mode_and_strategy(0, strategy, b1, b2)
    got = ctx.create_synthetic_rgb_by_mode_and_strategy(0, strategy, b1, b2)
    assert np.array_equal(got, ref)


def test_pol_ops_and_small_stages(ctx):
    rng = np.random.default_rng(9)
    a = rng.gamma(2.0, 100.0, (211, 307)).astype(np.float32)
    b = rng.gamma(2.0, 40.0, (211, 307)).astype(np.float32)
    b[0, :10] = 0
    a[1, :10] = -b[1, :10]
    for op, fn in ((O.OP_SUM, ctx.sum_arrays), (O.OP_DIFF, ctx.difference_arrays), (O.OP_RATIO, ctx.ratio_arrays),
                   (O.OP_NDIFF, ctx.normalized_diff_arrays), (O.OP_LOGRATIO, ctx.log_ratio_arrays)):
        assert np.array_equal(fn(a, b), O.pol_op(op, a, b))
    d16 = rng.integers(3, 200, (100, 33)).astype(np.uint16)
    assert np.array_equal(ctx.scale_u16_to_u8(d16), O.scale_u16_to_u8(d16))
    assert np.array_equal(ctx.add_padding_to_square(d16, S.U16), O.add_padding_to_square(d16, O.U16))
    db, mask = ctx.process_scalar_data_inplace(a)
    dbo, masko = O.process_scalar_data_inplace(a)
    assert np.array_equal(mask, masko)
    assert np.allclose(db, dbo, rtol=1e-12, atol=0)  # contract: 1e-5 relative


# ---- general f32 rasters: polarization ops and non-integer inputs (thresholds instead of DN tables) ----
def _check_stats_f32(st, so):
    for k in EXACT_STATS:
        assert getattr(st, k) == getattr(so, k), k
    # mean/std come from fp32 logs on this path (documented): log lines and the Adaptive test only
    assert abs(st.mean_db - so.mean_db) <= 1e-4
    assert abs(st.std_db - so.std_db) <= 1e-4


@pytest.mark.parametrize("op", [S.OP_SUM, S.OP_DIFF, S.OP_RATIO, S.OP_NDIFF, S.OP_LOGRATIO])
@pytest.mark.parametrize("strategy", range(7))
@pytest.mark.parametrize("bit_depth", [S.U8, S.U16])
def test_pipeline_single_pol_op(ctx, op, strategy, bit_depth):
    vv = CASES["speckle"](301, 423).astype(np.float32)
    vh = CASES["speckle_vh"](301, 423).astype(np.float32)
    comb = O.pol_op(op, vv, vh)
    po = O.process_scalar_data_pipeline(comb, bit_depth, strategy, want_db=False)
    img = ctx.process_single(vv, S.TIFF, bit_depth, strategy, None, False, op=op, band2=vh)
    got, ref = (img.gray, po.u8) if bit_depth == S.U8 else (img.gray16, po.u16)
    assert np.array_equal(got, ref), (int((got != ref).sum()), int(np.abs(got.astype(int) - ref.astype(int)).max()))
    if strategy != S.ADAPTIVE or True:
        _check_stats_f32(img.stats[0], po.stats)


@pytest.mark.parametrize("strategy", [S.STANDARD, S.ROBUST, S.CLAHE, S.EQUALIZED])
@pytest.mark.parametrize("bit_depth", [S.U8, S.U16])
def test_non_integer_band_resized(ctx, strategy, bit_depth):
    """A calibrated (non-integer) f32 band through autoscale + Lanczos + pad."""
    dn = CASES["speckle"](640, 900).astype(np.float32)
    v = (dn * dn * np.float32(3.7e-4)).astype(np.float32)  # sigma0-like
    v[:, :30] = 0
    ref, meta = O.pipeline_single(v, O.TIFF, bit_depth, strategy, 256, True)
    img = ctx.process_single(v, S.TIFF, bit_depth, strategy, 256, True)
    got = img.gray if bit_depth == S.U8 else img.gray16
    assert np.array_equal(got, ref), int((got != ref).sum())
    u8, u16, st = ctx.process_scalar_data_pipeline(v, bit_depth, strategy)
    po = O.process_scalar_data_pipeline(v, bit_depth, strategy, want_db=False)
    assert np.array_equal(u8 if bit_depth == S.U8 else u16, po.u8 if bit_depth == S.U8 else po.u16)
    _check_stats_f32(st, po.stats)


def test_f32_edge_cases(ctx):
    neg = -np.abs(np.random.default_rng(1).normal(size=(64, 80))).astype(np.float32)  # diff with a < b: all invalid
    u8, _, st = ctx.process_scalar_data_pipeline(neg, S.U8, S.ROBUST)
    assert st.valid_count == 0 and not u8.any()
    const = np.full((50, 70), 0.37, np.float32)  # degenerate: all equal
    po = O.process_scalar_data_pipeline(const, S.U16, S.DEFAULT, want_db=False)
    _, u16, st = ctx.process_scalar_data_pipeline(const, S.U16, S.DEFAULT)
    assert np.array_equal(u16, po.u16) and st.p99 == po.stats.p99
    nan = np.random.default_rng(2).gamma(2.0, 3.0, (40, 60)).astype(np.float32)
    nan[::7, ::5] = np.nan
    nan[1::9, ::4] = np.inf
    po = O.process_scalar_data_pipeline(nan, S.U8, S.EQUALIZED, want_db=False)
    u8, _, st = ctx.process_scalar_data_pipeline(nan, S.U8, S.EQUALIZED)
    assert np.array_equal(u8, po.u8)


@pytest.mark.parametrize("strategy", [S.STANDARD, S.ROBUST, S.ADAPTIVE, S.EQUALIZED, S.TAMED, S.DEFAULT])
@pytest.mark.parametrize("bit_depth", [S.U8, S.U16])
def test_pipeline_polops_two_operations(ctx, strategy, bit_depth):
    """BASELINE config 4's call (sarpro_pipeline_polops): log-ratio and normalised difference of the same pair, each autoscaled on
    its own into a full-resolution band, operands read once per pass for both. Each band must equal the oracle's
    pol_op -> process_scalar_data_pipeline of that operation alone; f32 operands and the raw u16 DN give the same bytes."""
    vv = CASES["speckle"](487, 1001)
    vh = CASES["speckle_vh"](487, 1001)
    vh[100:140, 200:260] = 0   # invalid in one operand only
    vv[300:310, :50] = 0
    ops = (S.OP_LOGRATIO, S.OP_NDIFF)
    refs = [O.process_scalar_data_pipeline(O.pol_op(op, vv.astype(np.float32), vh.astype(np.float32)), bit_depth, strategy, want_db=False)
            for op in ops]
    for a, b in ((vv.astype(np.float32), vh.astype(np.float32)), (vv, vh)):
        planes, stats = ctx.process_polops(a, b, ops, bit_depth, strategy)
        for k in range(2):
            ref = refs[k].u8 if bit_depth == S.U8 else refs[k].u16
            assert planes[k].shape == ref.shape and planes[k].dtype == ref.dtype
            assert np.array_equal(planes[k], ref), (k, a.dtype, int((planes[k] != ref).sum()))
            _check_stats_f32(stats[k], refs[k].stats)
    # the one-operation form and the other operations through the same entry
    for op in (S.OP_SUM, S.OP_DIFF, S.OP_RATIO):
        ref = O.process_scalar_data_pipeline(O.pol_op(op, vv.astype(np.float32), vh.astype(np.float32)), bit_depth, strategy, want_db=False)
        planes, stats = ctx.process_polops(vv, vh, (op,), bit_depth, strategy)
        assert np.array_equal(planes[0], ref.u8 if bit_depth == S.U8 else ref.u16), op


@pytest.mark.parametrize("strategy,bit_depth", [(S.EQUALIZED, S.U16), (S.ROBUST, S.U8), (S.TAMED, S.U16), (S.STANDARD, S.U16)])
def test_pipeline_polops_large_guarded_index(strategy, bit_depth, monkeypatch):
    """6 M pixels per band through the general f32 kernels with the guarded direct index (kernels_f32.cu: most samples get
    their stat bin / level from one fp32 evaluation, the rest compare thresholds) and with the shortcut switched off
    (SARPRO_F32_NO_GUARD=1): both must equal the oracle byte for byte, statistics included."""
    from sarpro_b200.synth import synth_pair
    vv, vh = synth_pair(2000, 3000, point_targets=1e-4)
    ops = (S.OP_LOGRATIO, S.OP_NDIFF)
    refs = [O.process_scalar_data_pipeline(O.pol_op(op, vv.astype(np.float32), vh.astype(np.float32)), bit_depth, strategy, want_db=False)
            for op in ops]
    for env in ("0", "1"):
        monkeypatch.setenv("SARPRO_F32_NO_GUARD", env)
        with S.Context(0) as c:
            planes, stats = c.process_polops(vv, vh, ops, bit_depth, strategy)
        for k in range(2):
            ref = refs[k].u8 if bit_depth == S.U8 else refs[k].u16
            assert np.array_equal(planes[k], ref), (env, k, int((planes[k] != ref).sum()))
            _check_stats_f32(stats[k], refs[k].stats)


def test_pipeline_polops_degenerate_and_errors(ctx):
    """One operation without a valid sample rides along with a regular one (diff with a < b everywhere: all zeros); CLAHE and a
    third operation are refused with INVALID_ARGUMENT."""
    vv = CASES["speckle"](120, 333)
    vh = (vv.astype(np.uint32) + 7).astype(np.uint16)
    ops = (S.OP_DIFF, S.OP_SUM)
    planes, stats = ctx.process_polops(vv, vh, ops, S.U16, S.EQUALIZED)
    assert stats[0].valid_count == 0 and not planes[0].any()
    ref = O.process_scalar_data_pipeline(O.pol_op(S.OP_SUM, vv.astype(np.float32), vh.astype(np.float32)), S.U16, S.EQUALIZED, want_db=False)
    assert np.array_equal(planes[1], ref.u16)
    with pytest.raises(S.SarproError):
        ctx.process_polops(vv, vh, ops, S.U8, S.CLAHE)
    with pytest.raises(S.SarproError):
        ctx.process_polops(vv, vh, (S.OP_SUM, S.OP_DIFF, S.OP_RATIO), S.U8, S.ROBUST)
    with pytest.raises(S.SarproError):
        ctx.process_polops(vv, vh, (7,), S.U8, S.ROBUST)


@pytest.mark.parametrize("strategy", [S.CLAHE, S.ROBUST, S.TAMED])
@pytest.mark.parametrize("shape,target", [((3001, 4999), 1024), ((2500, 9000), 700), ((5003, 2001), 512), ((3001, 5000), 1024),
                                          ((1031, 2048), 300), ((517, 25000), 2048), ((2100, 4096), 2048)])
def test_production_kernels_match_exact_kernels(strategy, shape, target, monkeypatch):
    """The production pass-B kernel against the generic exact kernels (SARPRO_FORCE_EXACT=1, the ones the other tests pin
    to the oracle) on rasters large enough for several strips, CLAHE cells and row groups: the tensor-core kernel
    (kernels_hmma.cu: IMMA taps, fp32 bilinear form with truncation-bit risk test, warp-cooperative exact fix-up, marker
    fix-up of saturated bins in the border cells; used when the column count is a multiple of 8) and the generic kernel with
    the fast per-pixel tables (SARPRO_HMMA=0). Bright point targets exercise the clamped table range, the zeroed block the
    invalid-pixel entry."""
    from sarpro_b200.synth import synth_pair
    vv, vh = synth_pair(*shape, point_targets=1e-4)
    vv[shape[0] // 2:, -300:] = 0  # invalid block on the right edge
    outs = []
    envs = ({"SARPRO_FORCE_EXACT": "1"}, {"SARPRO_HMMA": "0"}, {})
    for env in envs:
        for k in ("SARPRO_FORCE_EXACT", "SARPRO_HMMA"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with S.Context(0) as c:
            outs.append(c.process_synrgb_jpeg(vv, vh, strategy, target, True).rgb.copy())
    for i in range(1, len(outs)):
        assert np.array_equal(outs[0], outs[i]), (envs[i], int((outs[0] != outs[i]).sum()))


def test_tensor_core_pass_b_against_oracle(ctx):
    """kernels_hmma.cu straight against the CPU oracle (columns a multiple of 8, partial last row group, point targets,
    an invalid block), CLAHE and a LUT strategy, u8 single band and synRGB."""
    from sarpro_b200.synth import synth_pair
    vv, vh = synth_pair(1211, 2048, point_targets=1e-4)
    vv[300:500, -120:] = 0
    for strategy in (S.CLAHE, S.ROBUST):
        ref, _ = O.pipeline_single(vv.astype(np.float32), O.TIFF, S.U8, strategy, 640, True)
        img = ctx.process_single(vv, S.TIFF, S.U8, strategy, 640, True)
        assert np.array_equal(img.gray, ref), (strategy, int((img.gray != ref).sum()))
    ref, _ = O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), S.CLAHE, 1024, True)
    img = ctx.process_synrgb_jpeg(vv, vh, S.CLAHE, 1024, True)
    assert np.array_equal(img.rgb, ref), int((img.rgb != ref).sum())


@pytest.mark.parametrize("strategy", [S.CLAHE, S.ROBUST])
def test_full_size_scene_against_oracle(strategy):
    """BASELINE's full size (25,000 x 16,000 per band, the C3 / C2 scene of bench.py) through the production path (tensor-core
    pass B, second band on the side stream) against the CPU oracle on the very same rasters: the 2048 x 2048 synRGB must be
    byte-identical. Also: a different call in between and a repeat on the same context give the same bytes (no state leaks
    between calls). The oracle takes about a minute per strategy on the GPU box's host."""
    import torch
    from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
    dev = torch.device("cuda:0")
    vv = synth_band_torch(16000, 25000, SEED_VV, dev)
    vh = synth_band_torch(16000, 25000, SEED_VH, dev, cross_pol=True)
    torch.cuda.synchronize(dev)  # the library works on its own stream: the generators (torch's stream) must have finished
    with S.Context(0) as c:
        out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
        c.process_synrgb_jpeg(vv, vh, strategy, 2048, True, out=out)
        first = out.cpu().numpy().copy()
        c.process_synrgb_jpeg(vh, vv, strategy, 2048, True, out=out)  # different call in between
        c.process_synrgb_jpeg(vv, vh, strategy, 2048, True, out=out)
        again = out.cpu().numpy().copy()
    assert np.array_equal(first, again)
    assert first[(2048 - 1311) // 2 + 5:-(2048 - 1311) // 2 - 5].any()
    vv_h = vv.cpu().numpy().view(np.uint16)
    vh_h = vh.cpu().numpy().view(np.uint16)
    del vv, vh, out
    torch.cuda.empty_cache()
    O.set_resize_threads(os.cpu_count() or 1)  # (the oracle's Lanczos stage may use threads like the crate's rayon feature; same result)
    ref, _ = O.pipeline_synrgb_jpeg(vv_h.astype(np.float32), vh_h.astype(np.float32), strategy, 2048, True)
    d = first != ref
    assert not d.any(), (int(d.sum()), [int(d[..., ch].sum()) for ch in range(3)],
                         int(np.abs(first.astype(int) - ref.astype(int)).max()))


def _oracle_tamed(dn, is_copol):
    db, mask = O.process_scalar_data_inplace(dn.astype(np.float32))
    return O.autoscale_db_image_tamed_synrgb_u8(db, mask, is_copol)


def test_autoscale_tamed_synrgb_u8_stage(ctx):
    """autoscale.rs:710-742 through its own stage entry point (sarpro_autoscale_tamed_synrgb_u8), co- and cross-pol, on every
    fixture and on a raster of more than 64 K pixels; no scale_u16_to_u8 re-stretch on this path."""
    for case in sorted(CASES):
        dn = CASES[case](203, 317)
        for is_copol in (True, False):
            ref = _oracle_tamed(dn, is_copol)
            got = ctx.autoscale_db_image_tamed_synrgb_u8(dn.astype(np.float32), is_copol)
            assert np.array_equal(got, ref), (case, is_copol, int((got != ref).sum()))
    dn = CASES["speckle"](1100, 1900)
    for is_copol in (True, False):
        ref = _oracle_tamed(dn, is_copol)
        got = ctx.autoscale_db_image_tamed_synrgb_u8(dn.astype(np.float32), is_copol)
        assert np.array_equal(got, ref), (is_copol, int((got != ref).sum()))


def test_process_scalar_data_inplace_large(ctx):
    """pipeline.rs:8-40 on a raster well beyond 64 K pixels: mask exact, dB within 1e-12 relative (contract 1e-5)."""
    rng = np.random.default_rng(11)
    a = rng.gamma(2.0, 100.0, (1500, 2100)).astype(np.float32)
    a[::17, ::13] = 0
    a[5, :100] = -1.0
    a[6, :100] = np.float32(9.9e-6)   # just below the validity threshold
    a[7, :100] = np.float32(1.01e-5)  # just above it
    db, mask = ctx.process_scalar_data_inplace(a)
    dbo, masko = O.process_scalar_data_inplace(a)
    assert np.array_equal(mask, masko)
    assert np.allclose(db, dbo, rtol=1e-12, atol=0)


@pytest.mark.parametrize("strategy", [S.CLAHE, S.EQUALIZED])
def test_u16_resized_output_above_4mp(ctx, strategy):
    """U16 bit depth with a resize target that leaves more than 4 MP: the generic horizontal kernel with u16 samples (i32
    taps, i64 accumulate) and the u16 vertical pass, single band and multiband TIFF."""
    vv = CASES["speckle"](3000, 4400)
    vh = CASES["speckle_vh"](3000, 4400)
    ref, meta = O.pipeline_single(vv.astype(np.float32), O.TIFF, S.U16, strategy, 2600, True)
    img = ctx.process_single(vv, S.TIFF, S.U16, strategy, 2600, True)
    assert img.gray16.shape == ref.shape and ref.size > 4_000_000
    assert np.array_equal(img.gray16, ref), int((img.gray16 != ref).sum())
    r1, r2, _ = O.pipeline_multiband_tiff(vv.astype(np.float32), vh.astype(np.float32), S.U16, strategy, 2600, False)
    mb = ctx.process_multiband_tiff(vv, vh, S.U16, strategy, 2600, False)
    assert np.array_equal(mb.gray16, r1) and np.array_equal(mb.gray16_band2, r2)


def test_axis_cache_survives_many_shapes():
    """More than 64 distinct resize shapes on one context (the axis-plan cache is bounded and is only emptied between
    calls): every result still matches the oracle."""
    rng = np.random.default_rng(3)
    with S.Context(0) as c:
        for i in range(40):  # two axes per call
            rows, cols = 60 + 3 * i, 90 + 5 * i
            data = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
            ref, _ = O.resize_image_data_with_meta(data, 40 + i, S.U8, True)
            got, _ = c.resize_image_data_with_meta(data, 40 + i, S.U8, True)
            assert np.array_equal(got, ref), i


def test_enum_arguments_are_validated(ctx):
    dn = CASES["speckle"](64, 80)
    with pytest.raises(S.SarproError):
        ctx.process_single(dn.astype(np.float32), S.TIFF, S.U8, S.ROBUST, None, False, op=5, band2=dn.astype(np.float32))
    with pytest.raises(S.SarproError):
        ctx.process_scalar_data_pipeline(dn, S.U8, 7)
    with pytest.raises(S.SarproError):
        ctx.process_scalar_data_pipeline(dn, 2, S.ROBUST)


@pytest.mark.parametrize("strategy,name", [(S.CLAHE, "clahe"), (S.ROBUST, "robust"), (S.TAMED, "tamed")])
def test_golden_wide_scene_on_the_device(ctx, strategy, name):
    """Committed golden outputs (tests/golden/golden_wide.npz, written by make_golden.py from the oracle) for a scene wide
    enough for the tensor-core pass B: the CUDA path reproduces them byte for byte, from u16 DN and from f32 bands."""
    from tests.test_oracle_cpu import _wide_scene
    g, vv, vh = _wide_scene()
    for a, b in ((vv, vh), (vv.astype(np.float32), vh.astype(np.float32))):
        img = ctx.process_synrgb_jpeg(a, b, strategy, 256, True)
        assert np.array_equal(img.rgb, g[f"synrgb_{name}"]), int((img.rgb != g[f"synrgb_{name}"]).sum())


@pytest.mark.parametrize("strategy", [S.STANDARD, S.ADAPTIVE, S.CLAHE])
def test_present_list_planner_matches_dense_planner(strategy, monkeypatch):
    """The planner reads the device-compacted list of present DNs (k_hist_total); SARPRO_DENSE_PLAN=1 makes it scan the dense
    65,536-bin totals as before. Same statistics and samples, also for a raster with more distinct DNs than the list holds
    (every u16 value: the library falls back to the dense totals by itself), both against the oracle."""
    from sarpro_b200.synth import synth_band
    rng = np.random.default_rng(5)
    grd = synth_band(900, 1400, 3, block=16)
    grd[5:9, 100:300] = 60000
    wide = rng.integers(0, 65536, size=(700, 900), dtype=np.uint16)  # ~65k distinct DNs: list overflow
    for dn in (grd, wide):
        po = O.process_scalar_data_pipeline(dn.astype(np.float32), S.U8, strategy)
        got = {}
        for dense in ("", "1"):
            monkeypatch.delenv("SARPRO_DENSE_PLAN", raising=False)
            if dense:
                monkeypatch.setenv("SARPRO_DENSE_PLAN", "1")
            with S.Context(0) as c:
                u8, _, st = c.process_scalar_data_pipeline(dn, S.U8, strategy)
                got[dense] = (u8.copy(), st)
        assert np.array_equal(got[""][0], got["1"][0])
        assert np.array_equal(got[""][0], po.u8)
        for k in ("valid_count", "min_db", "max_db", "mean_db", "std_db", "median_db", "p01", "p99", "low_clip", "high_clip", "gamma"):
            assert getattr(got[""][1], k) == getattr(got["1"][1], k), k


# ---- the device planner (kernels_plan.cu) against the host planner (plan.cpp) -------------------------------------------
_DEV_PLAN_CASES = [(S.ROBUST, 0), (S.EQUALIZED, 0), (S.CLAHE, 0), (S.TAMED, 0), (S.DEFAULT, 0), (S.TAMED, 1), (S.TAMED, 2)]


def _fixture_histograms():
    from sarpro_b200.synth import synth_band
    rng = np.random.default_rng(17)
    hists = {}
    for case in sorted(CASES):
        hists[case] = np.bincount(CASES[case](203, 317).ravel(), minlength=65536)
    hists["grd"] = np.bincount(synth_band(900, 1400, 3, block=16, point_targets=1e-4).ravel(), minlength=65536)
    hists["every_dn"] = np.bincount(rng.integers(0, 65536, size=700 * 900, dtype=np.uint16), minlength=65536)
    hists["no_valid"] = np.bincount(np.zeros(1000, np.uint16), minlength=65536)
    one = np.zeros(65536, np.int64)
    one[777] = 12345
    hists["single_dn"] = one
    big = np.zeros(65536, np.int64)   # counts beyond 2^31 in total: 64-bit prefix sums
    big[100:4000] = 1_000_000
    big[0] = 5
    hists["four_billion"] = big
    return hists


@pytest.mark.parametrize("strategy,kind", _DEV_PLAN_CASES)
@pytest.mark.parametrize("bit_depth", [S.U8, S.U16])
def test_device_planner_matches_host_planner(ctx, strategy, kind, bit_depth):
    """Every percentile, the window, the whole DN -> sample / bin table and the table range of the tensor-core pass B are
    bit-identical between the single-CTA device planner and the host planner; mean / std agree to 1e-12 (the host sums in
    long double, the device in f64 trees; neither is the reference's serial Welford)."""
    for name, hist in _fixture_histograms().items():
        st_h, lut_h, hot_h = S.plan_kind_from_dn_histogram(hist, bit_depth, strategy, kind)
        st_d, lut_d, hot_d = ctx.plan_on_device(hist, bit_depth, strategy, kind)
        for k in EXACT_STATS:
            assert getattr(st_d, k) == getattr(st_h, k), (name, k, getattr(st_d, k), getattr(st_h, k))
        assert abs(st_d.mean_db - st_h.mean_db) <= 1e-12 * max(1.0, abs(st_h.mean_db)), name
        assert abs(st_d.std_db - st_h.std_db) <= 1e-12 * max(1.0, abs(st_h.std_db)), name
        assert np.array_equal(lut_d, lut_h), (name, int((lut_d != lut_h).sum()))
        if st_h.valid_count:
            assert hot_d == hot_h, (name, hot_d, hot_h)


@pytest.mark.parametrize("strategy", [S.CLAHE, S.ROBUST, S.TAMED])
def test_device_plan_and_host_plan_give_the_same_image(strategy, monkeypatch):
    """SARPRO_HOST_PLAN=1 (every band planned on the host, two round trips) against the default (planned on the device, none):
    same bytes, same statistics, and the default path reports zero host synchronisations."""
    from sarpro_b200.synth import synth_pair
    vv, vh = synth_pair(1211, 2048, point_targets=1e-4)
    got = {}
    for host_plan in ("", "1"):
        monkeypatch.delenv("SARPRO_HOST_PLAN", raising=False)
        if host_plan:
            monkeypatch.setenv("SARPRO_HOST_PLAN", "1")
        with S.Context(0) as c:
            img = c.process_synrgb_jpeg(vv, vh, strategy, 640, True)
            got[host_plan] = (img.rgb.copy(), img.stats, c.timing().host_syncs)
    assert np.array_equal(got[""][0], got["1"][0])
    assert got[""][2] == 0 and got["1"][2] >= 2
    for a, b in zip(got[""][1], got["1"][1]):
        for k in EXACT_STATS:
            assert getattr(a, k) == getattr(b, k), k


# ---- batch entry (config 5) -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [S.BATCH_MULTIBAND, S.BATCH_SYNRGB])
def test_pipeline_batch_matches_single_calls_and_oracle(ctx, kind):
    """sarpro_pipeline_batch (api/mod.rs:474-536's scene loop): scenes of differing shapes and dtypes, one skipped product,
    one failing scene (bands of different shape) with continue_on_error; every processed scene equals the oracle's pipeline
    on that scene, the report counts like BatchReport."""
    from sarpro_b200.synth import synth_pair
    shapes = [(900, 1400), (1211, 2048), (640, 777), (900, 1400)]
    scenes = []
    for i, sh in enumerate(shapes):
        vv, vh = synth_pair(*sh, scene=i)
        if i == 2:
            vv, vh = vv.astype(np.float32), vh.astype(np.float32)  # the f32 boundary in the middle of a batch
        scenes.append((vv, vh))
    scenes.insert(1, None)                                             # skipped product
    scenes.insert(3, (scenes[0][0], scenes[2][1]))                     # shapes differ -> error, loop continues
    results, statuses, rep, stats = ctx.process_batch(scenes, kind, S.U8, S.CLAHE, 512, True)
    assert (rep.processed, rep.skipped, rep.errors) == (4, 1, 1)
    assert statuses[3] != 0 and results[3] is None and results[1] is None
    for k, sc in enumerate(scenes):
        if sc is None or k == 3:
            continue
        a, b = (np.asarray(x).astype(np.float32) for x in sc)
        if kind == S.BATCH_SYNRGB:
            ref, _ = O.pipeline_synrgb_jpeg(a, b, S.CLAHE, 512, True)
            assert np.array_equal(results[k][0], ref), k
        else:
            r1, r2, _ = O.pipeline_multiband_tiff(a, b, S.U8, S.CLAHE, 512, True)
            assert np.array_equal(results[k][0], r1) and np.array_equal(results[k][1], r2), k
    # first error is returned when continue_on_error is off; scenes before it are processed
    with pytest.raises(S.SarproError):
        ctx.process_batch(scenes, kind, S.U8, S.CLAHE, 512, True, continue_on_error=False)
    # an empty batch is a no-op
    _, _, rep0, _ = ctx.process_batch([], kind, S.U8, S.ROBUST, 256, True)
    assert (rep0.processed, rep0.skipped, rep0.errors) == (0, 0, 0)


# ---- downsample-on-read (f2) ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,out", [((1600, 2500), (205, 131)), ((1311, 2048), (512, 328)), ((999, 1777), (100, 57)),
                                       ((64, 40000), (2048, 4)), ((700, 900), (900, 700)), ((3000, 180), (12, 200))])
@pytest.mark.parametrize("alg", [S.RESAMPLE_AVERAGE, S.RESAMPLE_LANCZOS])
def test_read_band_resampled_matches_oracle(ctx, shape, out, alg):
    """sarpro_read_band_resampled (gdal.rs:145-177) against the oracle's restatement of GDAL's RasterIO resampling: every f32
    sample bit-identical (f64 accumulation in the same order on both sides), u16 DN and f32 sources, host and device buffers,
    reductions from 1 (identity shape) to 20, spans wider than the staged segment (64 x 40000 -> 2048 columns is 19.5 per
    output pixel; 3000 x 180 -> 12 columns falls back to direct loads)."""
    import torch
    rows, cols = shape
    oc, orr = out
    if alg == S.RESAMPLE_LANCZOS and (cols / oc > 6 or rows / orr > 6):
        pytest.skip("the reader only picks Lanczos below a reduction of 4")
    dn = CASES["speckle"](rows, cols)
    ref = O.read_band_resampled(dn, oc, orr, alg)
    got = ctx.read_band_resampled(dn, oc, orr, alg)
    assert got.dtype == np.float32 and got.shape == (orr, oc)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), int((got != ref).sum())
    f = (dn.astype(np.float32) * np.float32(0.37)).astype(np.float32)
    ref_f = O.read_band_resampled(f, oc, orr, alg)
    dev_in = torch.from_numpy(f).cuda()
    dev_out = torch.empty((orr, oc), dtype=torch.float32, device="cuda")
    ctx.read_band_resampled(dev_in, oc, orr, alg, out=dev_out)
    assert np.array_equal(dev_out.cpu().numpy().view(np.uint32), ref_f.view(np.uint32))


def test_read_then_pipeline_is_the_cli_size_flow(ctx):
    """The flow the CLI takes with --size (sentinel1.rs:1074-1109 then save.rs:317-368): both bands averaged down on read, the
    f32 rasters (left on the device) through the synRGB pipeline at the same target. Equals the oracle's chain byte for byte."""
    import torch
    from sarpro_b200.synth import synth_pair
    vv, vh = synth_pair(3000, 4700)
    oc, orr, alg = S.Context.read_dims_for_target(4700, 3000, 512)
    assert (oc, orr, alg) == O.read_dims_for_target(4700, 3000, 512)
    small = []
    for band in (vv, vh):
        t = torch.empty((orr, oc), dtype=torch.float32, device="cuda")
        ctx.read_band_resampled(band, oc, orr, alg, out=t)
        small.append(t)
    img = ctx.process_synrgb_jpeg(small[0], small[1], S.CLAHE, 512, True)
    r1 = O.read_band_resampled(vv, oc, orr, alg)
    r2 = O.read_band_resampled(vh, oc, orr, alg)
    ref, _ = O.pipeline_synrgb_jpeg(r1, r2, S.CLAHE, 512, True)
    assert np.array_equal(img.rgb, ref), int((img.rgb != ref).sum())
    with pytest.raises(S.SarproError):
        ctx.read_band_resampled(vv, 5000, 3000, S.RESAMPLE_AVERAGE)  # never enlarges


# ---- encoder hand-off (f3) --------------------------------------------------------------------------------------------
def _decode_jpeg(stream):
    import io
    from PIL import Image
    im = Image.open(io.BytesIO(stream))
    im.load()
    return np.asarray(im), im


def test_encode_jpeg_decodes_to_the_same_pixels(ctx):
    """sarpro_encode_jpeg (io/writers/jpeg.rs:6-30 on the GPU): a baseline JPEG at quality 100 (all-ones quantisation tables,
    4:4:4) must decode to the pixels that went in, within what a q=100 encode / decode round trip costs on a noisy image: +-2
    for gray (DCT rounding on both sides), +-8 and a mean of 1.5 levels for RGB (the YCbCr round trip's two integer colour
    conversions on top; measured 6 / 1.0 with Pillow's libjpeg as the decoder). Host and device sources, odd sizes."""
    import torch
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:517, 0:771]
    gray = np.clip(128 + 90 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + rng.normal(0, 12, xx.shape), 0, 255).astype(np.uint8)
    dec, im = _decode_jpeg(ctx.encode_jpeg(gray, 100))
    assert im.mode == "L" and dec.shape == gray.shape
    assert int(np.abs(dec.astype(int) - gray.astype(int)).max()) <= 2
    rgb = np.stack([gray, np.roll(gray, 40, 1), 255 - gray], axis=-1).copy()
    for src in (rgb, torch.from_numpy(rgb).cuda()):
        dec, im = _decode_jpeg(ctx.encode_jpeg(src, 100))
        assert im.mode == "RGB" and dec.shape == rgb.shape
        d = np.abs(dec.astype(int) - rgb.astype(int))
        assert int(d.max()) <= 8 and float(d.mean()) < 1.5, (int(d.max()), float(d.mean()))
    with pytest.raises(S.SarproError):
        ctx.encode_jpeg(gray, 0)


def test_pipeline_result_encoded_where_it_lies(ctx):
    """The JPEG flow without the raw image crossing PCIe: sarpro_pipeline_synrgb with an output of location NONE, then
    sarpro_encode_last_jpeg. The decoded stream equals the oracle's RGB within the codec tolerance; the gray bands likewise."""
    vv = CASES["speckle"](900, 1400)
    vh = CASES["speckle_vh"](900, 1400)
    ref, _ = O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), S.CLAHE, 512, True)
    img = ctx.process_synrgb_jpeg(vv, vh, S.CLAHE, 512, True, out=S.Context.KEEP)
    assert img.rgb is None and (img.width, img.height) == (512, 512)
    t = ctx.timing()
    assert t.d2h_bytes == 0
    dec, _ = _decode_jpeg(ctx.encode_last_jpeg(0, 100))
    d = np.abs(dec.astype(int) - ref.astype(int))
    assert dec.shape == ref.shape and int(d.max()) <= 8 and float(d.mean()) < 1.5, (int(d.max()), float(d.mean()))
    g1, _ = _decode_jpeg(ctx.encode_last_jpeg(1, 100))
    one = ctx.process_single(vv, S.JPEG, S.U8, S.CLAHE, 512, True)
    assert int(np.abs(g1.astype(int) - one.gray.astype(int)).max()) <= 2
    # a call that does not leave a u8 image invalidates the hand-off
    ctx.process_polops(vv, vh, (S.OP_RATIO,), S.U16, S.ROBUST)
    with pytest.raises(S.SarproError):
        ctx.encode_last_jpeg(0, 100)


# ---- streamed upload (f4) ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("strategy", [S.CLAHE, S.ROBUST, S.STANDARD])
def test_streamed_upload_gives_the_same_bytes(strategy, monkeypatch):
    """Host u16 bands of 64 MB and more are uploaded in row chunks on a copy stream, pass A consumes the chunks as they land
    (work units ordered by their last row) and the first band's pass B runs beside the second band's upload. The result must
    equal the one-copy path (SARPRO_STREAM_UPLOAD=0) and the device-resident path byte for byte; stats included."""
    import torch
    from sarpro_b200.synth import synth_pair
    vv, vh = synth_pair(4200, 8192, point_targets=1e-4)   # 68.8 MB per band
    outs = []
    for env in ("1", "0"):
        monkeypatch.setenv("SARPRO_STREAM_UPLOAD", env)
        with S.Context(0) as c:
            img = c.process_synrgb_jpeg(vv, vh, strategy, 1024, True)
            outs.append(img.rgb.copy())
            mb = c.process_multiband_tiff(vv, vh, S.U8, strategy, 700, False)
            outs.append(np.stack([mb.gray, mb.gray_band2]))
            st = [s.as_dict() for s in mb.stats]
            again = c.process_synrgb_jpeg(vv, vh, strategy, 1024, True)  # a second call reuses the staging buffers
            assert np.array_equal(again.rgb, img.rgb)
        outs.append(st)
    assert np.array_equal(outs[0], outs[3]) and np.array_equal(outs[1], outs[4]) and outs[2] == outs[5]
    with S.Context(0) as c:
        dvv, dvh = torch.from_numpy(vv.view(np.int16)).cuda(), torch.from_numpy(vh.view(np.int16)).cuda()
        dev = c.process_synrgb_jpeg(dvv, dvh, strategy, 1024, True)
    assert np.array_equal(dev.rgb, outs[0])


def test_large_f32_host_rasters_are_narrowed_on_the_host(monkeypatch):
    """The reference's own boundary: host f32 rasters GDAL made of u16 bands (gdal.rs:123). Rasters of 64 MB of DNs and more
    are narrowed to u16 by the library's host threads, chunk by chunk into pinned staging while the previous chunk uploads, and
    then take the streamed path (half the PCIe bytes). The result must equal the oracle on the same f32 rasters, the u16-host
    path and the f32 upload with the device-side bridge (SARPRO_NARROW_UPLOAD=0) byte for byte, with invalid samples (NaN,
    negatives, zeros) mixed in; a raster that is NOT u16-valued must be noticed wherever the odd sample sits and give the
    bytes of the general path."""
    from sarpro_b200.synth import synth_pair
    rows, cols = 4200, 8192
    vv, vh = synth_pair(rows, cols, point_targets=1e-4)
    vv_f, vh_f = vv.astype(np.float32), vh.astype(np.float32)
    vv_f[100:140, 3000:3300] = np.nan          # no-data block
    vv_f[4100:, :50] = -7.0
    vh_f[::977, ::13] = 0.0
    vv_u = np.where(np.isfinite(vv_f) & (vv_f > 0), vv_f, 0).astype(np.uint16)
    vh_u = vh_f.astype(np.uint16)
    n = rows * cols
    monkeypatch.setenv("SARPRO_NARROW_UPLOAD", "2")   # narrow whatever the host's speed (the default gives up on a slow host)
    with S.Context(0) as c:
        img = c.process_synrgb_jpeg(vv_f, vh_f, S.CLAHE, 1024, True)
        assert c.timing().h2d_bytes == 2 * n * 2, c.timing().h2d_bytes     # both bands crossed PCIe as u16
        rgb = img.rgb.copy()
        again = c.process_synrgb_jpeg(vv_f, vh_f, S.CLAHE, 1024, True)     # the staging slots are reused
        assert np.array_equal(again.rgb, rgb)
        from_u16 = c.process_synrgb_jpeg(vv_u, vh_u, S.CLAHE, 1024, True).rgb.copy()
        mb = c.process_multiband_tiff(vv_f, vh_f, S.U16, S.ROBUST, 900, False)
        mb_n = np.stack([mb.gray, mb.gray_band2]).copy()
        # not u16-valued: one half-integer sample in the last chunk of the second band, then in the very first sample
        odd = []
        for pos in ((rows - 3, cols - 5), (0, 0)):
            w = vh_f.copy()
            w[pos] = 1234.5
            odd.append(c.process_synrgb_jpeg(vv_f, w, S.ROBUST, 1024, True).rgb.copy())
            assert c.timing().h2d_bytes == n * 2 + n * 4                   # VV as DNs, VH as f32 after the refused attempt
    monkeypatch.setenv("SARPRO_NARROW_UPLOAD", "0")
    with S.Context(0) as c:
        plain = c.process_synrgb_jpeg(vv_f, vh_f, S.CLAHE, 1024, True)
        assert c.timing().h2d_bytes == 2 * n * 4
        assert np.array_equal(plain.rgb, rgb)
        mb = c.process_multiband_tiff(vv_f, vh_f, S.U16, S.ROBUST, 900, False)
        assert np.array_equal(np.stack([mb.gray, mb.gray_band2]), mb_n)
        for k, pos in enumerate(((rows - 3, cols - 5), (0, 0))):
            w = vh_f.copy()
            w[pos] = 1234.5
            assert np.array_equal(c.process_synrgb_jpeg(vv_f, w, S.ROBUST, 1024, True).rgb, odd[k])
    assert np.array_equal(from_u16, rgb)
    O.set_resize_threads(os.cpu_count() or 1)
    ref, _ = O.pipeline_synrgb_jpeg(vv_f, vh_f, S.CLAHE, 1024, True)
    assert np.array_equal(rgb, ref), int((rgb != ref).sum())


# ---- rasters whose width is not a multiple of 8 (re-pitched for the aligned kernels) ----------------------------------
@pytest.mark.parametrize("cols", [4097, 4103, 5001])
def test_repitched_raster_matches_oracle(ctx, cols, monkeypatch):
    """A raster whose width is not a multiple of 8 is re-pitched (host rasters by their upload, device rasters by one kernel; the
    1..7 padding columns replicate the edge sample, carry no tap and stay out of every statistic) so that the tensor-core pass B
    and the aligned loads apply. Host and device inputs, CLAHE (position-dependent samples: the padding must not leak into the
    re-stretch extrema) and a LUT strategy, u8 synRGB and U16 bands: all equal to the oracle, and to SARPRO_REPITCH=0."""
    import torch
    from sarpro_b200.synth import synth_pair
    vv, vh = synth_pair(1300, cols, point_targets=1e-4)
    vv[vv == 0] = 1   # no invalid pixel in band 1: its minimum sample is not 0, the re-stretch decision depends on real pixels only
    dvv, dvh = torch.from_numpy(vv.view(np.int16)).cuda(), torch.from_numpy(vh.view(np.int16)).cuda()
    for strategy in (S.CLAHE, S.ROBUST):
        ref, _ = O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), strategy, 512, True)
        for a, b in ((vv, vh), (dvv, dvh)):
            img = ctx.process_synrgb_jpeg(a, b, strategy, 512, True)
            assert np.array_equal(img.rgb, ref), (strategy, type(a).__name__, int((img.rgb != ref).sum()))
        r1, r2, _ = O.pipeline_multiband_tiff(vv.astype(np.float32), vh.astype(np.float32), S.U16, strategy, 400, False)
        mb = ctx.process_multiband_tiff(dvv, dvh, S.U16, strategy, 400, False)
        assert np.array_equal(mb.gray16, r1) and np.array_equal(mb.gray16_band2, r2), strategy
    monkeypatch.setenv("SARPRO_REPITCH", "0")
    with S.Context(0) as c:
        plain = c.process_synrgb_jpeg(vv, vh, S.CLAHE, 512, True)
    assert np.array_equal(plain.rgb, O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), S.CLAHE, 512, True)[0])
