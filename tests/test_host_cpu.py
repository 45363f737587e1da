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
def test_tensor_core_tap_plan_replays_the_horizontal_pass(lib_built, in_size, out_size, max_span):
    """Host logic of kernels_hmma.cu: the n-tile / k-step plan and the permuted hi/lo tap bytes of the B fragments, replayed
    in the device's order on one row, must give the same bytes as the direct fixed-point Lanczos3 pass (and as the oracle's)."""
    rng = np.random.default_rng(in_size * 7 + out_size)
    row = rng.integers(0, 256, in_size).astype(np.uint8)
    row[: in_size // 9] = 255  # saturated run: the negative lobes must clamp identically
    direct = np.zeros(out_size, np.uint8)
    replay = np.full(out_size, 7, np.uint8)
    rc = _ffi.lib().sarpro_lanczos_row_plan_check(row.ctypes.data, in_size, out_size, max_span, direct.ctypes.data, replay.ctypes.data)
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
