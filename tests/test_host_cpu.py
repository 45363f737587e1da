"""CPU: host logic of the product (planner, shape arithmetic, C-ABI surface). No compute call needs a GPU;
the library must load here and export every symbol the header declares, and must refuse to create a
context without a device (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

import sarpro_b200 as S
from oracle import pyoracle as O
from sarpro_b200 import _ffi
from tests.fixtures import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXACT = ["valid_count", "min_db", "max_db", "median_db", "p01", "p02", "p05", "p10", "p25", "p75", "p90", "p95", "p98",
         "p99", "low_clip", "high_clip", "gamma"]


def test_header_symbols_are_exported(lib_built):
    hdr = open(os.path.join(ROOT, "include", "sarpro_gpu.h")).read()
    declared = set(re.findall(r"\b(sarpro_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", _ffi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (sarpro_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    lib = _ffi.lib()
    assert lib.sarpro_abi_version() == 1


def test_no_cpu_fallback(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(S.SarproError) as e:
        S.Context(0)
    assert e.value.code == _ffi.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """The product tree never references oracle/ (the oracle is test infrastructure only)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sarpro_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("strategy", range(7))
def test_present_list_planner_equals_dense_planner(lib_built, case, strategy):
    """The production planner input is the device-compacted list of present DNs (k_hist_total: one {offset, count} entry per
    block of 256 DNs into a {dn, count} pair list, blocks allocated in arbitrary order). Emulated on the host, with the blocks
    shuffled, it must give the plan of the dense 65,536-bin histogram bit for bit (which the next test pins to the oracle);
    a list that overflowed its capacity must be refused (the library then reads the dense totals)."""
    dn = CASES[case](203, 317)
    rng = np.random.default_rng(strategy)
    dn.ravel()[rng.integers(0, dn.size, 40)] = rng.integers(1, 65536, 40)  # sparse bright DNs in the upper blocks
    hist = np.bincount(dn.ravel(), minlength=65536).astype(np.uint64)
    n_present = int((hist > 0).sum())
    for bit_depth in (S.U8, S.U16):
        st, lut = S.plan_from_dn_histogram(hist, bit_depth, strategy)
        blocks, pairs = S.present_list_from_histogram(hist, 8192, order=rng.permutation(256))
        got = S.plan_from_present_list(blocks, pairs, bit_depth, strategy)
        assert got is not None
        assert got[0].as_dict() == st.as_dict() or all(
            (a == b) or (a != a and b != b) for a, b in zip(got[0].as_dict().values(), st.as_dict().values()))
        assert np.array_equal(got[1], lut)
    if n_present > 8:
        blocks, pairs = S.present_list_from_histogram(hist, n_present - 1)
        assert S.plan_from_present_list(blocks, pairs, S.U8, strategy) is None


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("strategy", range(7))
@pytest.mark.parametrize("bit_depth", [S.U8, S.U16])
def test_planner_matches_oracle(lib_built, case, strategy, bit_depth):
    dn = CASES[case](203, 317)
    hist = np.bincount(dn.ravel(), minlength=65536).astype(np.uint64)
    st, lut = S.plan_from_dn_histogram(hist, bit_depth, strategy)
    v = dn.astype(np.float32)
    po = O.process_scalar_data_pipeline(v, bit_depth, strategy, want_db=True)
    for k in EXACT:
        assert getattr(st, k) == getattr(po.stats, k), k
    assert st.mean_db == pytest.approx(po.stats.mean_db, rel=1e-10, abs=1e-12)
    assert st.std_db == pytest.approx(po.stats.std_db, rel=1e-9, abs=1e-12)
    if strategy != S.CLAHE:
        ref = po.u8 if bit_depth == S.U8 else po.u16
        assert np.array_equal(lut[dn].astype(ref.dtype), ref)
    elif po.stats.valid_count:
        # the LUT holds the CLAHE bin of autoscale.rs:263: round(clamp(norm) * 255)
        lo, hi = po.stats.low_clip, po.stats.high_clip
        rng = max(hi - lo, 1.0)
        norm = (np.fmin(np.fmax(po.db, lo), hi) - lo) / rng
        bins = np.floor(np.clip(norm, 0, 1) * 255.0 + 0.5).astype(np.uint16)
        valid = po.mask.astype(bool)
        assert np.array_equal(lut[dn][valid], bins[valid])


def test_resize_output_dims_matches_oracle(lib_built):
    for cols, rows in ((25000, 16000), (16000, 25000), (640, 480), (333, 1000), (90, 100), (7, 5000), (1, 1)):
        for target in (None, 64, 333, 2048, 100000):
            for pad in (False, True):
                assert S.Context.resize_output_dims(cols, rows, target, pad) == O.resize_output_dims(cols, rows, target, pad)


def test_shard_rows(lib_built):
    for rows in (16000, 16001, 100, 7):
        for world in (1, 2, 4, 8):
            for clahe in (False, True):
                bands = [S.shard_rows(rows, world, r, clahe) for r in range(world)]
                assert bands[0][0] == 0 and bands[-1][1] == rows
                assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
                if clahe:
                    tile_h = -(-rows // 8)
                    assert all(b[0] % tile_h == 0 for b in bands)
    with pytest.raises(S.SarproError):
        S.shard_rows(100, 2, 2, False)


@pytest.mark.parametrize("in_size,out_size,max_span", [(25000, 2048, 3125), (25000, 2048, 0), (25000, 1024, 3125), (16000, 1311, 0),
                                                        (4096, 2048, 512), (9000, 700, 1125), (2048, 300, 256), (5000, 1024, 625),
                                                        (1400, 512, 175), (4999, 1024, 0), (640, 600, 0)])
@pytest.mark.parametrize("strip_nt", [0, 16, 8, 4, 1])
def test_tensor_core_tap_plan_replays_the_horizontal_pass(lib_built, in_size, out_size, max_span, strip_nt):
    """Host logic of kernels_hmma.cu: the n-tile / k-step plan and the permuted hi/lo tap bytes of the B fragments, replayed
    in the device's order on one row, must give the same bytes as the direct fixed-point Lanczos3 pass (and as the oracle's).
    strip_nt: strips of at most that many n-tiles (0 = 32), the finer work units of a sharded rank's short band."""
    rng = np.random.default_rng(in_size * 7 + out_size)
    row = rng.integers(0, 256, in_size).astype(np.uint8)
    row[: in_size // 9] = 255  # saturated run: the negative lobes must clamp identically
    direct = np.zeros(out_size, np.uint8)
    replay = np.full(out_size, 7, np.uint8)
    rc = _ffi.lib().sarpro_lanczos_row_plan_check(row.ctypes.data, in_size, out_size, max_span, strip_nt, direct.ctypes.data, replay.ctypes.data)
    assert rc in (0, 1)
    ref = O.resize_u8_image(np.tile(row, (1, 1)), out_size, 1) if hasattr(O, "resize_u8_image") else None
    if ref is not None:
        assert np.array_equal(direct, np.asarray(ref).reshape(-1))
    if rc == 0:
        # only axes the kernel does not take: widths that are not a multiple of 8, or scale factors below ~6.5, where more
        # than three n-tiles (8 output columns) meet one 64-column block; those run on the other pass-B kernels
        assert in_size % 8 != 0 or in_size < 7 * out_size
    else:
        assert np.array_equal(direct, replay)
    if in_size % 8 == 0 and in_size >= 8 * out_size:
        assert rc == 1


def test_rust_sys_crate_matches_the_header():
    """integration/sarpro-gpu-sys/src/lib.rs (source only, there is no Rust toolchain here) declares exactly the
    entry points of include/sarpro_gpu.h with the argument lists the generator derives from it."""
    r = subprocess.run([os.sys.executable, os.path.join(ROOT, "integration", "gen_sys.py"), "--check"])
    assert r.returncode == 0
    lib_rs = open(os.path.join(ROOT, "integration", "sarpro-gpu-sys", "src", "lib.rs")).read()
    assert set(re.findall(r"pub fn (sarpro_[a-z0-9_]+)\(", lib_rs)) == set(_ffi.SYMBOLS)


@pytest.mark.parametrize("low,rng,n,vmin,vmax", [(-12.3, 21.7, 65535, 2e-3, 80.0), (-48.0, 47.5, 65535, 1.6e-5, 0.999), (3.1, 9.0, 255, 0.5, 40.0),
                                                 (-30.0, 75.0, 4096, 1e-3, 3e4), (-0.7, 1.0, 65535, 0.7, 1.2), (10.0, 33.0, 4096, 10.0, 2e4)])
def test_f32_guard_bound_holds_under_worst_case_log2_error(lib_built, low, rng, n, vmin, vmax):
    """Host proof obligation of the general f32 kernels' shortcut (kernels_f32.cu f32_guarded_index): replaying the device's
    fp32 operation sequence with the mantissa logarithm pushed to BOTH ends of MUFU.LG2's documented error (+-2^-22), every
    sample the guard accepts gets the index the reference's f64 expression gives; and the guard accepts most samples."""
    import ctypes as C
    e0, f0, sc, gd = C.c_int(), C.c_float(), C.c_float(), C.c_float()
    assert _ffi.lib().sarpro_f32_guard_params(low, rng, n, vmin, vmax, C.byref(e0), C.byref(f0), C.byref(sc), C.byref(gd)) == 0
    g = np.float32(gd.value)
    assert 0 < g < 0.45
    r = np.random.default_rng(n + int(rng * 10))
    v = np.exp(r.uniform(np.log(vmin), np.log(vmax), 400_000)).astype(np.float32)
    # values sitting right at level boundaries as well
    kk = r.integers(0, n + 1, 100_000)
    vb = (10.0 ** ((low + rng * kk / n) / 10.0)).astype(np.float32)
    vb = np.concatenate([vb, np.nextafter(vb, np.float32(0)), np.nextafter(vb, np.float32(np.inf))])
    v = np.concatenate([v, vb[(vb >= vmin) & (vb <= vmax)]])
    db = 10.0 * np.log10(v.astype(np.float64))
    q = (np.clip(db, low, low + rng) - low) / rng * n
    exact = np.where(q >= n, n, np.floor(np.clip(q, 0, None))).astype(np.int64)
    bits = v.view(np.uint32)
    e = (bits >> 23).astype(np.int32) - 127
    m = ((bits & 0x7fffff) | 0x3f800000).view(np.float32)
    accepted = 0
    for delta in (-2.0 ** -22, 0.0, 2.0 ** -22):
        lg = (np.log2(m.astype(np.float64)) + delta).astype(np.float32)
        d = ((e - e0.value).astype(np.float32) + (lg - np.float32(f0.value)).astype(np.float32)).astype(np.float32)
        t = (d * np.float32(sc.value)).astype(np.float32)
        fl = np.floor(t)
        fr = (t - fl).astype(np.float32)
        dec = np.abs(fr - np.float32(0.5)) <= np.float32(0.5) - g   # the device's test: the fraction is clear of both integers
        idx = np.clip(fl.astype(np.int64), 0, n)                   # ... and then clamp(floor(t), 0, top) is the index
        assert np.array_equal(idx[dec], exact[dec]), int((idx[dec] != exact[dec]).sum())
        accepted = int(dec[:400_000].sum())
    assert accepted > 0.5 * 400_000


def test_rust_wrapper_covers_the_c_abi():
    """integration/gpu.rs (the safe wrapper the patches call) may only name entry points the header declares, and names every
    pipeline-level one; the patches only call wrapper methods that exist."""
    import re
    hdr = open(os.path.join(ROOT, "include", "sarpro_gpu.h")).read()
    declared = set(re.findall(r"\b(sarpro_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(ROOT, "integration", "gpu.rs")).read()
    used = set(re.findall(r"sys::(sarpro_[a-z0-9_]+)\s*\(", rs))
    assert used <= declared, used - declared
    for must in ("sarpro_pipeline_single", "sarpro_pipeline_multiband_tiff", "sarpro_pipeline_synrgb", "sarpro_pipeline_polops",
                 "sarpro_pipeline_batch", "sarpro_read_band_resampled", "sarpro_encode_last_jpeg", "sarpro_process_scalar_data_pipeline",
                 "sarpro_resize_image_data_with_meta", "sarpro_add_padding_to_square", "sarpro_create_synthetic_rgb_by_mode_and_strategy",
                 "sarpro_pol_op", "sarpro_autoscale_tamed_synrgb_u8", "sarpro_process_scalar_data_inplace"):
        assert must in used, must
    methods = set(re.findall(r"pub fn ([a-z0-9_]+)\s*\(", rs))
    pdir = os.path.join(ROOT, "integration", "patches")
    called = set()
    for name in sorted(os.listdir(pdir)):
        for line in open(os.path.join(pdir, name)):
            if line.startswith("+"):
                called |= set(re.findall(r"\|g\| g\.([a-z0-9_]+)\(", line))
                called |= set(re.findall(r"^\+\s+g\.([a-z0-9_]+)\(", line))
    assert called and called <= methods, called - methods


@pytest.mark.parametrize("in_size,out_size", [(25000, 2048), (16000, 1311), (1000, 1000), (1000, 999), (4097, 512), (77, 3), (8192, 2048)])
@pytest.mark.parametrize("alg", [0, 1])
def test_read_resample_tables_replay_the_oracle(lib_built, in_size, out_size, alg):
    """Host logic of downsample-on-read (plan_read.cpp: the spans and weights the kernels read): one row through the product's
    tables, accumulated like the kernels do, equals the oracle's restatement of GDAL's resampling bit for bit."""
    if alg == 1 and in_size > 6 * out_size:
        pytest.skip("the reader only picks Lanczos below a reduction of 4")
    rng = np.random.default_rng(in_size + out_size)
    row = rng.integers(0, 65536, in_size).astype(np.uint16)
    row[: in_size // 7] = 65535
    got = np.zeros(out_size, np.float32)
    assert _ffi.lib().sarpro_read_row_plan_check(row.ctypes.data, in_size, out_size, alg, got.ctypes.data) == 0
    ref = O.read_band_resampled(row[None, :], out_size, 1, alg)[0]
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), int((got != ref).sum())


_EDGE_CASES = [
    # kind, low, high, gamma, n_levels, vmin, vmax
    (0, -12.3, 14.7, 1.0, 65535, 1.0 / 1800.0, 1500.0),      # config 4: Equalized window of a VV/VH ratio, u16 levels
    (0, -21.4, 3.9, 1.0, 65535, 1.6e-5, 0.999),              # normalised difference, u16
    (0, -12.3, 14.7, 1.0, 255, 2e-3, 80.0),                  # u8 levels
    (0, -0.7, -0.2, 1.0, 65535, 0.7, 1.2),                   # window narrower than 1 dB: range raised to 1, top levels unreachable
    (0, -60.0, 90.0, 1.0, 65535, 1.1e-5, 3.0e8),             # window wider than the data
    (0, 3.0, 9.0, 1.0, 65535, 2.5, 2.6),                     # data inside a sliver of the window
    (0, -12.3, 14.7, 0.8, 65535, 2e-3, 80.0),                # gamma != 1: every entry by search
    (0, -12.3, 14.7, 1.6, 255, 2e-3, 80.0),
    (1, -18.0, 6.5, 1.0, 255, 1e-3, 40.0),                   # Tamed u8 levels
    (2, -18.0, 6.5, 1.0, 255, 1e-3, 40.0),                   # CLAHE bins (a round): every entry by search
    (-1, 0.0, 0.0, 1.0, 0, 1.0 / 1800.0, 1500.0),            # stat bins over the data range
    (-1, 0.0, 0.0, 1.0, 0, 1.6e-5, 3.0e8),
    (-1, 0.0, 0.0, 1.0, 0, 0.9991, 1.0007),                  # span of 7e-3 dB: bins 1.7e-6 dB wide
    (-1, 0.0, 0.0, 1.0, 0, 1.0, 1.0000002),                  # two adjacent f32 values
]


@pytest.mark.parametrize("kind,low,high,gamma,n_levels,vmin,vmax", _EDGE_CASES)
def test_f32_threshold_tables_analytic_equals_search(lib_built, kind, low, high, gamma, n_levels, vmin, vmax):
    """The general f32 path's threshold tables (plan_f32.cpp) place most boundaries analytically (one exp per boundary, with
    a margin argument) and search only the rest. The analytic table must equal, bit for bit, the table found by bracketing
    every boundary with the reference's own f64 expression; and that table must have the threshold property under an
    independent numpy evaluation of the expression (level(edge) >= k > level(previous f32))."""
    import ctypes as C
    na, bad = C.c_uint32(), C.c_uint32()
    size = 4096 if kind < 0 else n_levels + 1
    edges = np.zeros(size, np.float32)
    rc = _ffi.lib().sarpro_f32_edges_check(kind, low, high, gamma, n_levels, vmin, vmax, C.byref(na), C.byref(bad),
                                           edges.ctypes.data_as(C.c_void_p))
    assert rc == size
    assert bad.value == 0
    vmin32, vmax32 = np.float32(vmin), np.float32(vmax)

    def level(v):
        db = 10.0 * np.log10(np.maximum(v.astype(np.float64), 1e-10))
        if kind < 0:
            lo_db, hi_db = 10.0 * np.log10(np.float64(vmin32)), 10.0 * np.log10(np.float64(vmax32))
            t = np.clip((db - lo_db) * (1.0 / (hi_db - lo_db)), 0.0, 1.0)
            return np.minimum((t * 4096.0).astype(np.int64), 4095)
        rng = max(high - low, 1.0)
        n = (np.clip(db, low, high) - low) / rng
        if kind == 2:
            return np.clip(np.floor(np.clip(n, 0.0, 1.0) * 255.0 + 0.5), 0, 255).astype(np.int64)   # f64::round, positive arguments
        q = np.clip((n if gamma == 1.0 else np.power(n, gamma)) * float(n_levels), 0.0, float(n_levels))
        return np.where(q >= n_levels, n_levels, q.astype(np.int64))

    k = np.arange(1, size)
    e = edges[1:]
    finite = np.isfinite(e)
    top = int(level(np.array([vmax32]))[0])
    bottom = int(level(np.array([vmin32]))[0])
    # unreachable levels are +inf, reachable ones lie in [vmin, vmax]
    assert np.array_equal(finite, k <= top)
    ef, kf = e[finite], k[finite]
    assert np.all((ef >= vmin32) & (ef <= vmax32))
    assert np.all(level(ef) >= kf)
    inner = ef > vmin32                       # levels the smallest sample already reaches sit at vmin
    assert np.all(kf[~inner] <= bottom)
    assert np.all(level(np.nextafter(ef[inner], np.float32(0))) < kf[inner])
    analytic_expected = gamma == 1.0 and kind != 2 and size > 2 and top - bottom > 8 and not (kind < 0 and vmax / vmin < 1.0001)
    if analytic_expected:
        assert na.value > 0.9 * (top - bottom - 1), (na.value, top, bottom)
    if gamma != 1.0 or kind == 2:
        assert na.value == 0


def test_f32_threshold_tables_random_windows(lib_built):
    """Same comparison over random windows, ranges and level counts (analytic == search on every entry)."""
    import ctypes as C
    r = np.random.default_rng(20251018)
    na, bad = C.c_uint32(), C.c_uint32()
    placed = 0
    for i in range(40):
        vmin = float(np.exp(r.uniform(np.log(1.1e-5), np.log(50.0))))
        vmax = vmin * float(np.exp(r.uniform(0.0, 12.0)))
        low = 10 * np.log10(vmin) + r.uniform(-3, 8)
        high = low + r.uniform(0.2, 60.0)
        kind = int(r.integers(-1, 2))
        n_levels = 255 if kind == 1 else int(r.choice([255, 4095, 65535]))
        rc = _ffi.lib().sarpro_f32_edges_check(kind, low, high, 1.0, n_levels, vmin, vmax, C.byref(na), C.byref(bad), None)
        assert rc > 0 and bad.value == 0, (i, kind, low, high, n_levels, vmin, vmax, bad.value)
        placed += na.value
    assert placed > 100_000


def test_host_narrowing_of_f32_rasters(lib_built):
    """Large f32 host rasters are narrowed to their DN by the host threads before they cross PCIe (narrow.cpp). The rule is
    the one of the reference's validity test (pipeline.rs:19-22: 10 log10(max(v, 1e-10)) > -50) plus 'is a u16': checked here
    against numpy on u16-valued data with invalid samples mixed in, on every alignment of the vector loop, and on rasters that
    are not u16-valued (which must be refused, wherever the offending sample sits)."""
    import ctypes as C
    lib = _ffi.lib()
    r = np.random.default_rng(7)

    def run(a, dst_off=0):
        # dst_off shifts the destination: 32-byte aligned destinations take the non-temporal stores of the staging path
        buf = np.zeros(a.size + 64, np.uint16)
        base = (-buf.ctypes.data // 2) % 16          # elements up to the next 32-byte boundary
        out = buf[base + dst_off: base + dst_off + a.size]
        flag = C.c_int(-1)
        assert lib.sarpro_narrow_f32_check(a.ctypes.data_as(C.c_void_p), a.size, out.ctypes.data_as(C.c_void_p), C.byref(flag)) == 0
        return out.copy(), flag.value

    for n in (0, 1, 7, 15, 16, 17, 31, 33, 1000, 256 * 1024 + 5, 3 * 256 * 1024 + 123):
        dn = r.integers(0, 65536, n).astype(np.uint16)
        a = dn.astype(np.float32)
        special = r.random(n)
        a[special < 0.02] = np.nan
        a[(special >= 0.02) & (special < 0.04)] = -3.0
        a[(special >= 0.04) & (special < 0.06)] = 0.0
        a[(special >= 0.06) & (special < 0.08)] = 9.9e-6          # below -50 dB
        a[(special >= 0.08) & (special < 0.09)] = -np.inf
        a[(special >= 0.09) & (special < 0.10)] = -0.5             # not a whole number, but not valid either
        with np.errstate(invalid="ignore", divide="ignore"):
            valid = 10.0 * np.log10(np.maximum(a.astype(np.float64), 1e-10)) > -50.0
        want = np.where(valid, np.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0), 0).astype(np.uint16)
        for off in (0, 1, 3):                                     # unaligned starts of the source
            for dst_off in (0, 1, 8):
                got, ok = run(np.ascontiguousarray(a[off:]), dst_off)
                assert ok == 1
                assert np.array_equal(got, want[off:])
    # not u16-valued: a fraction, a value above 65535, +inf, 1e-5 (valid: -50 dB + a hair, and not a whole number)
    base = r.integers(1, 2000, 3 * 256 * 1024).astype(np.float32)
    for bad in (0.5, 1234.25, 65536.0, 1e9, np.inf, 1.0001e-5):
        for pos in (0, 5, 16, base.size // 2 + 3, base.size - 1):
            a = base.copy()
            a[pos] = bad
            assert run(a)[1] == 0, (bad, pos)
    assert run(base)[1] == 1
    assert run(np.full(40, 65535.0, np.float32)) [1] == 1


def test_planner_matches_oracle_on_random_distributions(lib_built):
    """Randomised version of test_planner_matches_oracle: uniform, speckle at five brightness levels, constant, half-invalid,
    three-valued, near-constant Gaussian, almost-empty and log-uniform rasters of random small shapes; every strategy and bit
    depth. Statistics bit-exact, LUT applied to the DNs == the oracle's samples."""
    rng = np.random.default_rng(2026)
    for it in range(24):
        rows, cols = int(rng.integers(3, 120)), int(rng.integers(3, 160))
        kind = it % 8
        if kind == 0:
            dn = rng.integers(0, 65536, (rows, cols))
        elif kind == 1:
            dn = np.rint(np.sqrt(rng.gamma(4.4, 1 / 4.4, (rows, cols))) * rng.choice([3, 30, 150, 900, 5000])).clip(0, 65535)
        elif kind == 2:
            dn = np.full((rows, cols), int(rng.integers(0, 65536)))
        elif kind == 3:
            dn = np.rint(np.sqrt(rng.gamma(1.0, 1.0, (rows, cols))) * 100).clip(0, 65535)
            dn[rng.random((rows, cols)) < 0.4] = 0
        elif kind == 4:
            dn = rng.choice([1, 2, 65535], (rows, cols), p=[.5, .49, .01])
        elif kind == 5:
            dn = np.rint(rng.normal(1000, rng.choice([0.5, 3, 30]), (rows, cols))).clip(0, 65535)
        elif kind == 6:
            dn = np.zeros((rows, cols))
            dn.ravel()[:int(rng.integers(0, 4))] = rng.integers(1, 65536)
        else:
            dn = np.rint(np.exp(rng.uniform(0, np.log(65535), (rows, cols))))
        dn = dn.astype(np.uint16)
        hist = np.bincount(dn.ravel(), minlength=65536).astype(np.uint64)
        v = dn.astype(np.float32)
        for strategy in range(7):
            for bit_depth in (S.U8, S.U16):
                st, lut = S.plan_from_dn_histogram(hist, bit_depth, strategy)
                po = O.process_scalar_data_pipeline(v, bit_depth, strategy)
                for k in EXACT:
                    a, b = getattr(st, k), getattr(po.stats, k)
                    assert a == b or (a != a and b != b), (it, strategy, bit_depth, k, a, b)
                if strategy != S.CLAHE:
                    ref = po.u8 if bit_depth == S.U8 else po.u16
                    assert np.array_equal(lut[dn].astype(ref.dtype), ref), (it, strategy, bit_depth)


def test_tensor_core_tap_plan_on_random_axes(lib_built):
    """Randomised version of test_tensor_core_tap_plan_replays_the_horizontal_pass: random widths (mostly multiples of 8),
    scale factors 1..40, strip lengths and CLAHE spans; wherever a plan exists its replay equals the direct pass."""
    rng = np.random.default_rng(12345)
    plans = 0
    for _ in range(150):
        in_size = int(rng.integers(64, 30000))
        if rng.random() < 0.8:
            in_size = (in_size + 7) // 8 * 8
        out_size = max(1, int(in_size / float(np.exp(rng.uniform(0.0, np.log(40.0))))))
        max_span = int(rng.choice([0, (in_size + 7) // 8, 512]))
        strip_nt = int(rng.choice([0, 32, 16, 8, 4, 2, 1]))
        row = rng.integers(0, 256, in_size).astype(np.uint8)
        if rng.random() < 0.3:
            row[: in_size // 5] = 255
        direct = np.zeros(out_size, np.uint8)
        replay = np.full(out_size, 7, np.uint8)
        rc = _ffi.lib().sarpro_lanczos_row_plan_check(row.ctypes.data, in_size, out_size, max_span, strip_nt, direct.ctypes.data, replay.ctypes.data)
        assert rc in (0, 1), (rc, in_size, out_size, max_span, strip_nt)
        if rc == 1:
            plans += 1
            assert np.array_equal(direct, replay), (in_size, out_size, max_span, strip_nt)
    assert plans > 40


def _valid_threshold_f32():
    """Smallest f32 whose dB (pipeline.rs:19-20) exceeds -50 (pipeline.rs:22)."""
    lo, hi = np.float32(0.0).view(np.uint32), np.float32(1.0).view(np.uint32)
    while hi - lo > 1:
        mid = np.uint32((int(lo) + int(hi)) // 2)
        if 10.0 * np.log10(max(float(mid.view(np.float32)), 1e-10)) > -50.0:
            hi = mid
        else:
            lo = mid
    return hi.view(np.float32)


def _f32_edges(kind, low, high, gamma, n_levels, vmin, vmax):
    import ctypes as C
    size = 4096 if kind < 0 else n_levels + 1
    edges = np.zeros(size, np.float32)
    na, bad = C.c_uint32(), C.c_uint32()
    assert _ffi.lib().sarpro_f32_edges_check(kind, low, high, gamma, n_levels, vmin, vmax, C.byref(na), C.byref(bad),
                                             edges.ctypes.data_as(C.c_void_p)) == size
    assert bad.value == 0
    return edges


@pytest.mark.parametrize("make", ["ratio", "ndiff", "scaled", "narrow"])
@pytest.mark.parametrize("strategy", range(7))
def test_general_f32_path_host_logic_replays_the_oracle(lib_built, make, strategy):
    """The general f32 path (rasters that are not u16-valued: polarization products, calibrated inputs) on the CPU: the device
    kernels only COMPARE samples with host-built threshold tables, so numpy's searchsorted stands in for them and the product's
    own host steps do the rest - stat-bin thresholds, percentiles and window from the 4096-bin histogram, level thresholds.
    The replay must give the oracle's statistics bit for bit and its u16 / u8 samples (CLAHE: its bins) for every strategy."""
    import ctypes as C
    rng = np.random.default_rng(strategy * 10 + len(make))
    a = np.rint(np.sqrt(rng.gamma(4.4, 1 / 4.4, (97, 131))) * 150).astype(np.float32)
    b = np.rint(np.sqrt(rng.gamma(4.4, 1 / 4.4, (97, 131))) * 50).astype(np.float32)
    a[:, :5] = 0
    b[40:44, :] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        if make == "ratio":
            v = np.where(np.abs(b) > 1e-10, a / b, np.float32(0)).astype(np.float32)            # ops.rs:9-19
        elif make == "ndiff":
            v = np.where(np.abs(a + b) > 1e-10, (a - b) / (a + b), np.float32(0)).astype(np.float32)  # ops.rs:22-33 (half of it negative: invalid)
        elif make == "scaled":
            v = (a * np.float32(0.0371)).astype(np.float32)                                       # a calibrated-looking raster
        else:
            v = (np.float32(2.5) + rng.random((97, 131)).astype(np.float32) * np.float32(0.3)).astype(np.float32)  # < 1 dB of range
    thr = _valid_threshold_f32()
    valid = v >= thr
    assert valid.any()
    vmin, vmax = float(v[valid].min()), float(v[valid].max())
    # pass 2 of the device: stat-bin index = number of thresholds <= sample
    se = _f32_edges(-1, 0.0, 0.0, 1.0, 0, vmin, vmax)
    idx = np.searchsorted(se[1:], v[valid], side="right")
    hist = np.bincount(idx, minlength=4096).astype(np.uint64)
    db = 10.0 * np.log10(v[valid].astype(np.float64))
    for bit_depth in (S.U8, S.U16):
        po = O.process_scalar_data_pipeline(v, bit_depth, strategy)
        st = _ffi.Stats()
        assert _ffi.lib().sarpro_plan_from_stat_histogram(hist.ctypes.data_as(C.c_void_p), int(valid.sum()), vmin, vmax, float(db.mean()),
                                                          float(db.std()), strategy, 0, C.byref(st)) == 0
        for k in EXACT:
            x, y = getattr(st, k), getattr(po.stats, k)
            assert x == y or (x != x and y != y), (k, x, y)
        if strategy == S.CLAHE:
            le = _f32_edges(2, st.low_clip, st.high_clip, 1.0, 255, vmin, vmax)
            bins = np.searchsorted(le[1:], v[valid], side="right")
            rng_db = max(st.high_clip - st.low_clip, 1.0)
            want = np.floor(np.clip((np.clip(db, st.low_clip, st.high_clip) - st.low_clip) / rng_db, 0, 1) * 255.0 + 0.5).astype(np.int64)
            assert np.array_equal(bins, want)   # autoscale.rs:585-587, 263 (the rest of CLAHE runs on the DN machinery, tested elsewhere)
            continue
        # pass 3: level = number of level thresholds <= sample; invalid samples are written as 0 (autoscale.rs:437-447)
        n_levels = 255 if bit_depth == S.U8 else 65535
        le = _f32_edges(0, st.low_clip, st.high_clip, st.gamma, n_levels, vmin, vmax)
        q = np.zeros(v.shape, np.uint16)
        q[valid] = np.searchsorted(le[1:], v[valid], side="right")
        if bit_depth == S.U16:
            assert np.array_equal(q, po.u16), int((q != po.u16).sum())
        else:
            assert np.array_equal(O.scale_u16_to_u8(q), po.u8)   # autoscale.rs:669-670


def _synrgb_through_product_luts(b1, b2, suppressed):
    """Composes RGB the way k_synrgb does, from the product's host-built LUT set (and, for the suppressed variant, the host
    mirror of the device's floor rule on the combined histogram)."""
    import ctypes as C
    r, g, b = np.zeros(256, np.uint8), np.zeros(256, np.uint8), np.zeros(65536, np.uint8)
    fwc = C.c_int(-7)
    hist = (np.bincount(b1.ravel(), minlength=256) + np.bincount(b2.ravel(), minlength=256)).astype(np.uint32)
    rc = _ffi.lib().sarpro_synrgb_lut_check(-1 if suppressed else 41, hist.ctypes.data_as(C.c_void_p), b1.size, C.byref(fwc),
                                            r.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    assert rc == 0
    v1, v2 = b1.astype(np.int64), b2.astype(np.int64)
    rgb = np.stack([r[v1], g[v2], b[(v1 << 8) | v2]], axis=-1)
    if suppressed:
        rgb[(v1 <= fwc.value) & (v2 <= fwc.value)] = 0   # water short-circuit, synthetic_rgb.rs:160-166
    return rgb, fwc.value


def test_synrgb_lut_sets_match_the_oracle_on_every_pair(lib_built):
    """Every (band1, band2) pair of u8 samples through the product's channel LUTs == the oracle's create_synthetic_rgb /
    _suppressed (synthetic_rgb.rs:10-67, 88-178). The suppressed variant depends on the data through floor_with_cushion (p05 of
    the combined histogram + 3, capped at 40): the 65,536-pair grid is extended with a run of one dark value so that the
    floor sweeps its whole range 3..40, and with bright pixels so that it stays at the bottom."""
    g1, g2 = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    g1, g2 = g1.ravel(), g2.ravel()
    got, _ = _synrgb_through_product_luts(g1[None, :], g2[None, :], False)
    assert np.array_equal(got, O.create_synthetic_rgb(g1[None, :], g2[None, :]))
    floors = set()
    for p in list(range(0, 45)) + [60, 200]:
        # p05 of the grid alone is 12; a run of value p long enough pulls it to p (below 12) or pushes it up to p (above)
        extra = int((6553.6 - 512 * (p + 1)) / 0.95) + 50 if p <= 12 else 10240 * p - 131072 + 3000
        for fill in (p, 255):
            run = np.full(max(extra, 0) // 2 + 1, fill, np.uint8)
            b1 = np.concatenate([g1, run])[None, :]
            b2 = np.concatenate([g2, run])[None, :]
            got, fwc = _synrgb_through_product_luts(b1, b2, True)
            floors.add(fwc)
            ref = O.create_synthetic_rgb_suppressed(b1, b2)
            assert np.array_equal(got, ref), (p, fill, fwc, int((got != ref).any(axis=-1).sum()))
    assert floors == set(range(3, 41)), sorted(floors)   # every suppressed set the device holds except 0..2 (unreachable: +3)


def test_u16_lanczos_table_replays_the_oracle(lib_built):
    """The u16 Lanczos3 table of the product (i32 taps, i64 accumulate; the vertical pass and the generic horizontal kernel read
    it) on single rows == the oracle's resize_u16_image: fixed shapes around the reference's sizes and random ones, with
    saturated runs at both ends of the range (the negative lobes must clamp identically)."""
    rng = np.random.default_rng(5)
    shapes = [(25000, 2048), (16000, 1311), (4096, 2048), (1000, 999), (1000, 1000), (77, 3), (8, 1), (5003, 512)]
    shapes += [(int(n), max(1, int(n / float(np.exp(rng.uniform(0, np.log(30))))))) for n in rng.integers(8, 20000, 40)]
    for in_size, out_size in shapes:
        row = rng.integers(0, 65536, in_size).astype(np.uint16)
        row[: in_size // 6] = 65535
        row[-(in_size // 9) - 1:] = 0
        got = np.zeros(out_size, np.uint16)
        assert _ffi.lib().sarpro_lanczos_row_check_u16(row.ctypes.data, in_size, out_size, got.ctypes.data) == 0
        ref = np.asarray(O.resize_u16_image(row[None, :], out_size, 1)).reshape(-1)
        assert np.array_equal(got, ref), (in_size, out_size, int((got != ref).sum()))
