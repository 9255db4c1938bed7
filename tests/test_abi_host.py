"""CPU-only checks of the product library: the C ABI loads and exports every symbol include/radiofm_b200.h
declares, fails loudly without a device (no CPU fallback), and its HOST parts -- the planner that restates the
reference constructors and the integer RDS block-sync/FEC -- match the oracle and the golden fixtures.
No compute kernel is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, RATES, bits_equal

GOLDEN = os.path.join(ROOT, "tests", "golden")


def declared_symbols():
    out = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        txt = open(os.path.join(ROOT, "include", fn)).read()
        out += re.findall(r"RFM_API\s+[^;(]*?\b(rfm_\w+)\s*\(", txt)
    return out


def test_library_exports_every_declared_symbol(rfm):
    syms = declared_symbols()
    assert len(syms) >= 25
    lib = C.CDLL(rfm.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    assert set(syms) == set(rfm.EXPORTED_SYMBOLS)
    assert b"sm_100a" in rfm.lib().rfm_version()


def test_no_cpu_fallback(rfm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rfm.RadioFmError, match="no CUDA device"):
        rfm.FmDecoderBatch(1.0e6, -150e3, downsample=4)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "pvr.rtl.radiofm_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle|libradiofm_(oracle|ref)", txt, re.M), fn


@pytest.mark.parametrize("fs,ds", [(1.0e6, 4), (1.2e6, 5), (2.4e6, 11), (390625.0, 1), (2.048e6, 9), (960000.0, 4),
                                   (1.8e6, 8), (250000.0, 1)])
@pytest.mark.parametrize("usver", [False, True])
def test_planner_matches_oracle_constructors(rfm, port, fs, ds, usver):
    for off in (-0.15 * fs, 0.0, 37000.0):
        o = port.OracleFmDecoder(fs, off, downsample=ds, usver=usver)
        assert np.array_equal(o.constants()[:51], rfm.plan_constants(fs, off, downsample=ds, usver=usver)[:51])
        for w in range(6):
            assert bits_equal(o.table(w), rfm.plan_table(w, fs, off, downsample=ds, usver=usver)), f"table {w}"


def test_planner_matches_golden(rfm, port):
    for rate in ("1.0M", "1.2M", "2.4M", "390k"):
        g = np.load(os.path.join(GOLDEN, f"chain_{rate}.npz"))
        fs, ds = float(g["fs"]), int(g["ds"])
        assert np.array_equal(g["constants"][:51], rfm.plan_constants(fs, -0.15 * fs, downsample=ds)[:51])
        for w in range(6):
            assert bits_equal(g[f"table{w}"], rfm.plan_table(w, fs, -0.15 * fs, downsample=ds))
        # u8 conversion must be bit-exact (RTL_SDR_Source.cpp:207-211)
        assert bits_equal(g["cf32_first"], rfm.plan_table(6, fs, -0.15 * fs, downsample=ds))
    v = np.arange(256, dtype=np.float64) / (255.0 / 2.0) - 1.0
    assert bits_equal(v.astype(np.float32), rfm.plan_table(6, 1e6, 0.0))


def test_rds_check_block_kat(rfm):
    k = np.load(os.path.join(GOLDEN, "kat_rds_blocks.npz"))
    for w, osyn, fec, syn, fixed in zip(k["word"].tolist(), k["offset_syndrome"].tolist(), k["use_fec"].tolist(),
                                        k["syndrome"].tolist(), k["corrected"].tolist()):
        assert rfm.rds_check_block(w, osyn, bool(fec)) == (syn, fixed)


def test_rds_syndrome_tables_against_the_bit_loop(rfm):
    """the table-driven syndrome (four byte lookups) against the reference's 16-step loop (RDSProcess.cpp:386-396)
    restated here, on random 26-bit words and every single-bit word"""
    parity = [0x2DC, 0x16E, 0x0B7, 0x287, 0x39F, 0x313, 0x355, 0x376, 0x1BB, 0x201, 0x3DC, 0x1EE, 0x0F7, 0x2A7, 0x38F, 0x31B]

    def loop(word, osyn):
        block = word & 0x3FFFFFF
        syn = block >> 16
        for i in range(16):
            if block & 0x8000:
                syn ^= parity[i]
            block <<= 1
        return syn ^ osyn

    rng = np.random.default_rng(7)
    words = [1 << b for b in range(26)] + rng.integers(0, 1 << 26, 4000).tolist() + [0, (1 << 26) - 1, (1 << 32) - 1]
    for w in words:
        for osyn in (0x3D8, 0x258, 0x3CC):
            assert rfm.rds_check_block(int(w), osyn, False)[0] == loop(int(w), osyn)
    assert rfm.lib().rfm_source_block_length(70000) == 69632      # cRtlSdrSource's rule is host-only arithmetic


def test_rds_block_sync_matches_oracle(rfm, port, synth):
    rng = np.random.default_rng(3)
    groups = synth.rds_group_stream(0x1234, "ABCDEFGH", 40, radiotext="hello b200")
    bits = synth.rds_bits_from_groups(groups)
    noise = rng.integers(0, 2, 333).astype(np.uint8)
    stream = np.concatenate([noise, bits[:2000], noise[:57], bits[2000:]])
    # sprinkle burst errors (FEC) and a few uncorrectable ones (resync path)
    stream = stream.copy()
    for pos in rng.integers(400, stream.size - 10, 25):
        stream[pos:pos + int(rng.integers(1, 8))] ^= 1
    L = port.lib()
    h = L.rfo_rdssync_create()
    s = rfm.RdsBlockSync()
    got_o, got_p = [], []
    for chunk in np.array_split(stream, 17):  # ragged pushes incl. an empty one
        for part in (chunk, chunk[:0]):
            part = np.ascontiguousarray(part)
            L.rfo_rdssync_push_bits(h, part.ctypes.data_as(C.POINTER(C.c_uint8)), part.size)
            s.push_bits(part)
        out = np.zeros((256, 4), dtype=np.uint16)
        n = L.rfo_rdssync_take_groups(h, out.ctypes.data_as(C.POINTER(C.c_uint16)), 256)
        got_o.append(out[:n].copy())
        got_p.append(s.take_groups())
    L.rfo_rdssync_destroy(h)
    a, b = np.concatenate(got_o), np.concatenate(got_p)
    assert len(a) > 20 and np.array_equal(a, b)
    s.reset()
    assert len(s.take_groups()) == 0


def test_rds_block_sync_matches_golden_bits(rfm):
    g = np.load(os.path.join(GOLDEN, "rds_1.0M.npz"))
    s = rfm.RdsBlockSync()
    s.push_bits(g["bits"])
    assert np.array_equal(s.take_groups(), g["groups"])


def test_u8_conversion_formula(rfm):
    """The tiled front-end kernel converts u8 -> float with two FMAs instead of a table (rfm_u8_to_float in
    rfm_kernels.cu): t = fma(u, 0x1.01p-7, -1) [exact], v = fma(u, 0x1.010102p-23, t).  Emulated here with exact
    float64 products and a single rounding per FMA; must equal the reference expression for all 256 codes."""
    u = np.arange(256, dtype=np.float64)
    a_hi = float.fromhex("0x1.01p-7")
    a_lo = float.fromhex("0x1.010102p-23")
    t = u * a_hi - 1.0
    assert np.array_equal(t, t.astype(np.float32).astype(np.float64))      # first FMA is exact
    v = (u * a_lo + t).astype(np.float32)                                   # exact in float64, rounded once
    assert bits_equal(v, rfm.plan_table(6, 1e6, 0.0))
    assert bits_equal(v, (np.arange(256) / (255.0 / 2.0) - 1.0).astype(np.float32))


def test_source_block_length_rule(rfm):
    """cRtlSdrSource's block-length rule (RTL_SDR_Source.cpp:124-126): clamp to 4096 .. 2^20, then down to a multiple
    of 4096 -- host only; cRadioReceiver asks for 65536 (default_block_length)."""
    L = rfm.lib()
    L.rfm_source_block_length.restype = C.c_uint32
    L.rfm_source_block_length.argtypes = [C.c_uint32]
    for req, want in ((0, 4096), (1, 4096), (4095, 4096), (4096, 4096), (4097, 4096), (8191, 4096), (8192, 8192),
                      (65536, 65536), (65537, 65536), (100000, 98304), (1 << 20, 1 << 20), ((1 << 20) + 1, 1 << 20),
                      (0xFFFFFFFF, 1 << 20)):
        assert L.rfm_source_block_length(req) == want, (req, want)


def test_rtlsdr_adapter_binds_librtlsdr_at_run_time(rfm, tmp_path):
    """SURVEY.md 8f N4 (host logic, no GPU): librtlsdr is dlopen-ed -- absent, the calls answer RFM_ERR_UNSUPPORTED and
    nothing else breaks; present (here: the stand-in tests/cpp/fake_rtlsdr.c) every entry point the adapter needs resolves."""
    import ctypes as C
    import subprocess
    L = rfm.lib()
    L.rfm_rtlsdr_device_count.argtypes = [C.c_char_p]
    assert L.rfm_rtlsdr_device_count(str(tmp_path / "absent.so").encode()) == -3
    so = tmp_path / "libfake_rtlsdr.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests", "cpp", "fake_rtlsdr.c"), "-o", str(so)])
    assert L.rfm_rtlsdr_device_count(str(so).encode()) == 1
