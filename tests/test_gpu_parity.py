"""GPU parity suite (run with -m gpu on the B200 box): every test goes through the C ABI of
libldpc_b200.so and compares with the CPU oracle / the committed golden fixtures."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, awgn_llrs, golden

pytestmark = pytest.mark.gpu

CODES = ["moon.7.13", "jpl.1024.4.5", "jpl.4096.4.5", "1920.1280.3.303"]
TX = {"moon.7.13": None, "jpl.1024.4.5": 1280, "jpl.4096.4.5": 5120, "1920.1280.3.303": None}
EBN0 = {"moon.7.13": 3.0, "jpl.1024.4.5": 3.7, "jpl.4096.4.5": 3.6, "1920.1280.3.303": 2.4}


def _golden_files():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def _parse(fname):
    code, algo, it = fname[:-4].rsplit("_", 2)
    return code, algo, int(it[2:])


@pytest.fixture(scope="module")
def dcodes(gpu_lib, builtin_codes):
    return {n: gpu_lib.DeviceCode(c) for n, c in builtin_codes.items()}


def spa_close(a, b):
    """Message-level comparison for sum-product in fp64. Min-sum is exact; SPA goes through tanh/atanh of two
    different libms, and atanh amplifies a 1-ulp difference of the product p by 1/(1-p^2),
    so: 1e-9 relative where |msg| <= 12, and 1e-12 absolute in the tanh domain everywhere."""
    a, b = np.asarray(a), np.asarray(b)
    small = np.abs(b) <= 12.0
    ok_small = np.all(np.abs(a - b)[small] <= 1e-9 * np.maximum(np.abs(b[small]), 1e-300) + 1e-300)
    ok_tanh = np.all(np.abs(np.tanh(a / 2) - np.tanh(b / 2)) <= 1e-12)
    return ok_small and ok_tanh


@pytest.mark.parametrize("fname", _golden_files())
def test_fp64_stream_matches_golden(fname, gpu_lib, dcodes):
    code_name, algo, max_iters = _parse(fname)
    g = golden(fname)
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp64", max_batch=8, engine=gpu_lib.ENGINE_STREAM)
    bits, iters, conv = dec.decode(g["llr"], max_iters)
    assert np.array_equal(iters, g["iters"])
    assert np.array_equal(conv, g["converged"])
    assert np.array_equal(bits, g["bits"])
    for f in range(g["llr"].shape[0]):
        b, it, cv, c2v, lam = dec.decode_trace(g["llr"][f], max_iters)
        assert it == g["iters"][f] and cv == bool(g["converged"][f]) and np.array_equal(b, g["bits"][f])
        if it == 0:
            continue
        if algo in ("min", "arraylet-min"):      # bit-exact messages (arraylet-min: Fast/ArrayletMin.hs:167-192 order)
            assert np.array_equal(c2v[0], g["c2v_first"][f]) and np.array_equal(c2v[-1], g["c2v_last"][f])
            assert np.array_equal(lam[0], g["lam_first"][f]) and np.array_equal(lam[-1], g["lam_last"][f])
            if "c2v_all" in g.files:
                assert np.array_equal(c2v, g["c2v_all"][f][:it]) and np.array_equal(lam, g["lam_all"][f][:it])
        else:
            assert spa_close(c2v[0], g["c2v_first"][f]) and spa_close(c2v[-1], g["c2v_last"][f])
            assert spa_close(lam[0], g["lam_first"][f]) and spa_close(lam[-1], g["lam_last"][f])
    dec.close()


@pytest.mark.parametrize("code_name", CODES)
@pytest.mark.parametrize("algo", ["min", "spa"])
def test_fp64_stream_matches_oracle_batch(code_name, algo, gpu_lib, dcodes, builtin_codes, oracles):
    code = builtin_codes[code_name]
    F = 700 if code_name == "moon.7.13" else 160
    llr = awgn_llrs(code, EBN0[code_name], F, 1001, TX[code_name])
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp64", max_batch=256, engine=gpu_lib.ENGINE_STREAM)   # F > max_batch: chunking
    bits, iters, conv = dec.decode(llr, 20)
    rb, ri, rc = oracles[code_name].decode_batch(algo, 20, llr, variant=0)
    same = (bits == rb).all(axis=1)
    if algo == "min":
        assert same.all() and np.array_equal(iters, ri) and np.array_equal(conv, rc)
    else:
        assert same.mean() >= 0.999 and (iters == ri).mean() >= 0.99
    dec.close()


@pytest.mark.parametrize("code_name", CODES)
@pytest.mark.parametrize("algo", ["min", "spa"])
def test_fp32_stream_decisions_vs_oracle(code_name, algo, gpu_lib, dcodes, builtin_codes, oracles):
    """Production precision: hard decisions identical on >= 99.9 % of frames (BASELINE.json north_star)."""
    code = builtin_codes[code_name]
    F = 4096 if code_name == "moon.7.13" else (1024 if algo == "min" else 256)
    llr = awgn_llrs(code, EBN0[code_name] + 0.3, F, 2002, TX[code_name])
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp32", max_batch=1024, engine=gpu_lib.ENGINE_STREAM)
    bits, iters, conv = dec.decode(llr, 20)
    rb, ri, rc = oracles[code_name].decode_batch(algo, 20, llr, variant=1 if algo == "min" else 0)
    same = (bits == rb).all(axis=1)
    assert same.mean() >= 0.999, (same.mean(), F)
    dec.close()


def test_fp32_trace_within_1e4(gpu_lib, dcodes, builtin_codes, oracles):
    """fp32 production messages within 1e-4 relative of the fp64 oracle (min-sum, first iterations)."""
    code = builtin_codes["jpl.1024.4.5"]
    llr = awgn_llrs(code, 3.9, 4, 77, 1280)
    dec = gpu_lib.Decoder(dcodes["jpl.1024.4.5"], "min", "fp32", max_batch=4, engine=gpu_lib.ENGINE_STREAM)
    for f in range(4):
        llr32 = llr[f].astype(np.float32).astype(np.float64)     # the decoder rounds its input to fp32 first
        b, it, cv, c2v, lam = dec.decode_trace(llr32, 4)
        rb, rit, rcv, rc2v, rlam = oracles["jpl.1024.4.5"].decode("min", 4, llr32, trace=True)
        n = min(it, rit)
        assert n >= 1
        # 1e-4 relative; messages below 1.0 in magnitude are compared on an absolute 1e-4 scale because
        # x = ne - lam cancels (the error of a small difference is set by the size of its operands)
        assert (np.abs(c2v[:n] - rc2v[:n]) / np.maximum(np.abs(rc2v[:n]), 1.0)).max() <= 1e-4
        assert (np.abs(lam[:n] - rlam[:n]) / np.maximum(np.abs(rlam[:n]), 1.0)).max() <= 1e-4
    dec.close()


def test_edge_cases(gpu_lib, dcodes, builtin_codes, oracles):
    code = builtin_codes["jpl.1024.4.5"]
    dec = gpu_lib.Decoder(dcodes["jpl.1024.4.5"], "min", "fp64", max_batch=16, engine=gpu_lib.ENGINE_STREAM)
    # empty batch
    b, it, cv = dec.decode(np.zeros((0, code.N)), 20)
    assert b.shape == (0, code.N)
    # all-zero LLRs: syndrome passes at iteration 0 with the all-zero word (hard(0) = 0)
    b, it, cv = dec.decode(np.zeros((3, code.N)), 20)
    assert not b.any() and (it == 0).all() and cv.all()
    # max_iters = 0 and a hopeless frame: fallback to hard(channel)
    llr = awgn_llrs(code, -2.0, 5, 5, 1280)
    for mi in (0, 1, 3):
        b, it, cv = dec.decode(llr, mi)
        assert (it == mi).all() and not cv.any()
        assert np.array_equal(b, (llr > 0).astype(np.uint8))
    # a valid codeword is a fixed point
    from oracle import oracle
    msg = np.random.default_rng(0).integers(0, 2, code.K).astype(np.uint8)
    cw = oracle.encode(code, msg)
    b, it, cv = dec.decode((8.0 * (2.0 * cw - 1.0))[None, :], 20)
    assert it[0] == 0 and cv[0] and np.array_equal(b[0], cw)
    # ragged: batch of 1 through a decoder sized for 16, and 17 frames (2 chunks)
    llr = awgn_llrs(code, 3.8, 17, 6, 1280)
    b1, it1, cv1 = dec.decode(llr[:1], 20)
    b17, it17, cv17 = dec.decode(llr, 20)
    rb, ri, rc = oracles["jpl.1024.4.5"].decode_batch("min", 20, llr)
    assert np.array_equal(b17, rb) and np.array_equal(it17, ri) and np.array_equal(b1[0], rb[0])
    dec.close()


def test_bad_arguments_fail_loudly(gpu_lib, dcodes):
    import ctypes as C
    L = gpu_lib.load()
    h = C.c_void_p()
    rp = np.array([0, 2], dtype=np.int32)
    ci = np.array([1, 0], dtype=np.int32)          # not ascending
    assert L.ldpc_b200_code_create_csr(1, 2, rp.ctypes.data, ci.ctypes.data, C.byref(h)) == -1
    assert b"ascending" in L.ldpc_b200_last_error()
    sh = np.array([5], dtype=np.int32)             # shift >= sz
    assert L.ldpc_b200_code_create_qc(4, 1, 1, sh.ctypes.data, C.byref(h)) == -1
    assert L.ldpc_b200_decoder_create(dcodes["moon.7.13"].h, 7, 1, 4, C.byref(h)) == -1
    dec = gpu_lib.Decoder(dcodes["moon.7.13"], "min", "fp32", max_batch=4)
    with pytest.raises(gpu_lib.LdpcError):
        dec.set_engine(gpu_lib.ENGINE_RESIDENT)    # general (non-QC) code
    # the one-pass decode closure: take / stride / k_take are validated, nothing is silently clamped
    rx = np.zeros((2, 20)); out = np.zeros((2, 7), dtype=np.uint8)
    for stride, take, k_take in ((20, 0, 7), (20, 21, 7), (10, 20, 7), (20, 20, 0), (20, 20, 21)):
        assert L.ldpc_b200_decode_punctured(dec.h, 5, 2, rx.ctypes.data, stride, take, out.ctypes.data, k_take) == -1
    assert L.ldpc_b200_decode_punctured(dec.h, 5, 2, rx.ctypes.data, 20, 13, out.ctypes.data, 7) == 0
    dec.close()


def test_device_generator_statistics_and_simulate(gpu_lib, dcodes, builtin_codes, oracles):
    """generate_llr: mean / variance / punctured zeros; simulate == generate + decode + count."""
    code = builtin_codes["jpl.1024.4.5"]
    K, n_tx, F, ebn0 = code.K, 1280, 512, 3.6
    dec = gpu_lib.Decoder(dcodes["jpl.1024.4.5"], "min", "fp32", max_batch=256, engine=gpu_lib.ENGINE_STREAM)
    buf = gpu_lib.DeviceBuffer(F * code.N * 4)
    dec.generate_llr(buf.ptr, ebn0, K, n_tx, 42, 1000, F)
    dec.sync()
    llr = buf.download(np.float32, (F, code.N)).astype(np.float64)
    assert not llr[:, n_tx:].any()
    sigma2 = 1.0 / (2 * (K / n_tx) * 10 ** (ebn0 / 10))
    y = llr[:, :n_tx] * sigma2 / 2
    assert abs(y.mean() + 1.0) < 5 * np.sqrt(sigma2 / y.size)
    assert abs(y.var() / sigma2 - 1.0) < 0.02
    # frame-index keyed streams: a sub-range regenerates the same frames
    buf2 = gpu_lib.DeviceBuffer(16 * code.N * 4)
    dec.generate_llr(buf2.ptr, ebn0, K, n_tx, 42, 1100, 16)
    dec.sync()
    assert np.array_equal(buf2.download(np.float32, (16, code.N)), llr[100:116].astype(np.float32))
    # simulate over the same frames == oracle on the downloaded LLRs
    c = dec.simulate(20, ebn0, K, n_tx, 42, 1000, F)
    rb, ri, rc = oracles["jpl.1024.4.5"].decode_batch("min", 20, llr, variant=1)
    gb, gi, gc = dec.decode(llr, 20)
    same = (gb == rb).all(axis=1)
    assert same.mean() >= 0.999
    assert c["frames"] == F
    assert c["bit_errors"] == int(gb[:, :K].sum()) and c["frame_errors"] == int(gb[:, :K].any(axis=1).sum())
    assert c["sum_iters"] == int(gi.sum()) and c["unconverged"] == int((gc == 0).sum())
    dec.close()


# ------------------------------------------------------------------------------------------------
# RESIDENT engine (quasi-cyclic, on-chip)
# ------------------------------------------------------------------------------------------------
QC = ["jpl.1024.4.5", "jpl.4096.4.5"]


@pytest.mark.parametrize("code_name", QC)
def test_resident_fp64_bit_exact_vs_oracle(code_name, gpu_lib, dcodes, builtin_codes, oracles):
    """The fp64 instance follows the reference's association order, so decisions, iteration counts and the
    convergence flag are IDENTICAL to the oracle on every frame (min-sum is exact up to the a7 additions)."""
    code = builtin_codes[code_name]
    F = 300
    llr = np.concatenate([awgn_llrs(code, EBN0[code_name], F // 2, 3003, TX[code_name]),
                          awgn_llrs(code, EBN0[code_name] - 0.7, F // 2, 3004, TX[code_name])])
    dec = gpu_lib.Decoder(dcodes[code_name], "min", "fp64", max_batch=128, engine=gpu_lib.ENGINE_RESIDENT)
    assert dec.engine() == gpu_lib.ENGINE_RESIDENT
    bits, iters, conv = dec.decode(llr, 20)
    rb, ri, rc = oracles[code_name].decode_batch("min", 20, llr, variant=1)
    assert np.array_equal(bits, rb) and np.array_equal(iters, ri) and np.array_equal(conv, rc)
    assert 0 < conv.mean() < 1          # the operating point exercises both outcomes
    dec.close()


@pytest.mark.parametrize("fname", [f for f in _golden_files() if "_min_" in f and f.startswith("jpl")])
def test_resident_fp64_matches_golden(fname, gpu_lib, dcodes):
    code_name, algo, max_iters = _parse(fname)
    g = golden(fname)
    dec = gpu_lib.Decoder(dcodes[code_name], "min", "fp64", max_batch=8, engine=gpu_lib.ENGINE_RESIDENT)
    bits, iters, conv = dec.decode(g["llr"], max_iters)
    assert np.array_equal(bits, g["bits"]) and np.array_equal(iters, g["iters"]) and np.array_equal(conv, g["converged"])
    dec.close()


@pytest.mark.parametrize("code_name", QC)
def test_resident_fp32_vs_oracle_and_stream(code_name, gpu_lib, dcodes, builtin_codes, oracles):
    code = builtin_codes[code_name]
    F = 2048 if code_name == "jpl.1024.4.5" else 1024
    llr = awgn_llrs(code, EBN0[code_name] + 0.2, F, 4004, TX[code_name])
    res = gpu_lib.Decoder(dcodes[code_name], "min", "fp32", max_batch=1024)
    assert res.engine() == gpu_lib.ENGINE_RESIDENT          # AUTO picks the on-chip engine for the JPL codes
    bits, iters, conv = res.decode(llr, 20)
    rb, ri, rc = oracles[code_name].decode_batch("min", 20, llr, variant=1)
    same = (bits == rb).all(axis=1)
    assert same.mean() >= 0.999, same.mean()
    # same precision, same association order: the two engines agree exactly
    st = gpu_lib.Decoder(dcodes[code_name], "min", "fp32", max_batch=1024, engine=gpu_lib.ENGINE_STREAM)
    sb, si, sc = st.decode(llr, 20)
    assert np.array_equal(bits, sb) and np.array_equal(iters, si) and np.array_equal(conv, sc)
    # float32 entry point with packed output
    pk, i2, c2 = res.decode_f32(llr.astype(np.float32), 20)
    assert np.array_equal(gpu_lib.unpack_bits(pk, code.N), bits) and np.array_equal(i2, iters)
    res.close(); st.close()


@pytest.mark.parametrize("code_name", QC)
def test_resident_edge_cases(code_name, gpu_lib, dcodes, builtin_codes, oracles):
    code = builtin_codes[code_name]
    dec = gpu_lib.Decoder(dcodes[code_name], "min", "fp32", max_batch=8)
    b, it, cv = dec.decode(np.zeros((3, code.N)), 20)
    assert not b.any() and (it == 0).all() and cv.all()
    llr = awgn_llrs(code, -2.0, 5, 5, TX[code_name])
    for mi in (0, 1, 2):
        b, it, cv = dec.decode(llr, mi)
        assert (it == mi).all() and not cv.any() and np.array_equal(b, (llr > 0).astype(np.uint8))
    from oracle import oracle
    msg = np.random.default_rng(1).integers(0, 2, code.K).astype(np.uint8)
    cw = oracle.encode(code, msg)
    b, it, cv = dec.decode((6.0 * (2.0 * cw - 1.0))[None, :], 20)
    assert it[0] == 0 and cv[0] and np.array_equal(b[0], cw)
    # 19 frames through a decoder sized for 8: three chunks, last one ragged
    llr = awgn_llrs(code, EBN0[code_name] + 0.4, 19, 8, TX[code_name])
    b, it, cv = dec.decode(llr, 50)
    rb, ri, rc = oracles[code_name].decode_batch("min", 50, llr, variant=1)
    assert (b == rb).all(axis=1).mean() >= 18 / 19
    dec.close()


@pytest.mark.parametrize("code_name", QC)
def test_resident_device_api_and_fixed_iterations(code_name, gpu_lib, dcodes, builtin_codes):
    """decode_device with HBM-resident buffers; early_exit=0 runs all iterations but returns the same output."""
    code = builtin_codes[code_name]
    F, W = 512, code.N // 32
    dec = gpu_lib.Decoder(dcodes[code_name], "min", "fp32", max_batch=F)
    llr = awgn_llrs(code, EBN0[code_name] + 0.1, F, 9, TX[code_name]).astype(np.float32)
    d_llr = gpu_lib.DeviceBuffer(llr.nbytes).upload(llr)
    d_bits = gpu_lib.DeviceBuffer(F * W * 4)
    d_it = gpu_lib.DeviceBuffer(F * 4)
    d_cv = gpu_lib.DeviceBuffer(F)
    out = []
    for ee in (True, False):
        d_bits.zero()
        dec.decode_device(d_llr.ptr, d_bits.ptr, 20, F, early_exit=ee, d_iters=d_it.ptr, d_conv=d_cv.ptr)
        dec.sync()
        ms, launches = dec.last_timing()
        assert ms > 0 and launches >= 1
        out.append((d_bits.download(np.uint32, (F, W)), d_it.download(np.int32, (F,)), d_cv.download(np.uint8, (F,))))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    hb, hi, hc = dec.decode(llr.astype(np.float64), 20)
    assert np.array_equal(gpu_lib.unpack_bits(out[0][0], code.N), hb) and np.array_equal(out[0][1], hi)
    dec.close()


@pytest.mark.parametrize("code_name", QC)
def test_resident_simulate_equals_generate_decode_count(code_name, gpu_lib, dcodes, builtin_codes, oracles):
    code = builtin_codes[code_name]
    K, n_tx, F, ebn0 = code.K, TX[code_name], 600, EBN0[code_name]
    dec = gpu_lib.Decoder(dcodes[code_name], "min", "fp32", max_batch=256)
    c = dec.simulate(20, ebn0, K, n_tx, 7, 5000, F)
    buf = gpu_lib.DeviceBuffer(F * code.N * 4)
    dec.generate_llr(buf.ptr, ebn0, K, n_tx, 7, 5000, F)
    dec.sync()
    llr = buf.download(np.float32, (F, code.N)).astype(np.float64)
    gb, gi, gc = dec.decode(llr, 20)
    assert c["frames"] == F and c["sum_iters"] == int(gi.sum()) and c["unconverged"] == int((gc == 0).sum())
    assert c["bit_errors"] == int(gb[:, :K].sum()) and c["frame_errors"] == int(gb[:, :K].any(axis=1).sum())
    # sharding invariance: two half ranges add up to the whole
    c1 = dec.simulate(20, ebn0, K, n_tx, 7, 5000, 250)
    c2 = dec.simulate(20, ebn0, K, n_tx, 7, 5250, 350)
    for k in ("frames", "frame_errors", "bit_errors", "sum_iters", "unconverged"):
        assert c1[k] + c2[k] == c[k]
    # and the oracle agrees on the same frames
    rb, ri, rc = oracles[code_name].decode_batch("min", 20, llr, variant=1)
    assert (gb == rb).all(axis=1).mean() >= 0.998
    dec.close()


def test_ecc_record_mirror(gpu_lib, builtin_codes, oracles):
    """mkLDPC mirror (Utils.hs:35-75): name, lengths, encode, unpuncture + decode + truncate vs the oracle wrapper."""
    from ecc_ldpc_b200 import ecc
    from oracle import oracle
    (e,) = ecc.match("ldpc/b200-minsum/jpl.1024.4.5/20/4/5")
    code = builtin_codes["jpl.1024.4.5"]
    assert e.name == "ldpc/b200-minsum/jpl.1024.4.5/20/4/5"
    assert (e.message_length, e.codeword_length) == (1024, 1280)
    rng = np.random.default_rng(12)
    msg = rng.integers(0, 2, 1024).astype(np.uint8)
    cw = e.encode(msg)
    assert np.array_equal(cw, oracle.encode(code, msg, 1280))
    sigma = np.sqrt(1.0 / (2 * 0.8 * 10 ** 0.40))
    for t in range(6):
        rx = 2.0 * ((2.0 * cw - 1.0) + sigma * rng.standard_normal(1280)) / sigma ** 2
        out, flag = e.decode(rx)
        ref, rflag, it, cv = oracles["jpl.1024.4.5"].ecc_decode("min", 20, 1024, 1280, rx)
        assert flag and rflag and np.array_equal(out, ref)
        assert cv and np.array_equal(out, msg)
    e.close()
    (e2,) = ecc.match("ldpc/b200-spa-fp64/moon.7.13/20")
    assert e2.name == "ldpc/b200-spa-fp64/moon.7.13/20/7/20" and e2.codeword_length == 20
    m = rng.integers(0, 2, 7).astype(np.uint8)
    out, flag = e2.decode(5.0 * (2.0 * e2.encode(m) - 1.0))
    assert flag and np.array_equal(out, m)
    e2.close()


def test_simulate_sharded_single_rank(gpu_lib, dcodes, builtin_codes):
    """shard.simulate_sharded without a process group = plain simulate (the 2-rank case is tests/test_shard_cpu.py + bench --gpus 2)."""
    from ecc_ldpc_b200 import shard
    code = builtin_codes["jpl.1024.4.5"]
    dec = gpu_lib.Decoder(dcodes["jpl.1024.4.5"], "min", "fp32", max_batch=256)
    a = shard.simulate_sharded(dec, 20, 3.6, code.K, 1280, 5, 0, 700)
    b = dec.simulate(20, 3.6, code.K, 1280, 5, 0, 700)
    for k in shard.COUNTER_NAMES:
        assert a[k] == b[k]
    dec.close()


@pytest.mark.parametrize("code_name", QC)
def test_resident_spa_fp32_vs_oracle(code_name, gpu_lib, dcodes, builtin_codes, oracles):
    """Sum-product on the on-chip engine (log-domain fp32 form) against the literal fp64 oracle (Orig.hs:81-92):
    hard decisions identical on >= 99.9 % of frames, iteration counts equal on >= 98 %."""
    code = builtin_codes[code_name]
    F = 2048 if code_name == "jpl.1024.4.5" else 1024
    llr = np.concatenate([awgn_llrs(code, EBN0[code_name] - 0.2, F // 2, 6006, TX[code_name]),
                          awgn_llrs(code, EBN0[code_name] + 0.8, F // 2, 6007, TX[code_name])])
    dec = gpu_lib.Decoder(dcodes[code_name], "spa", "fp32", max_batch=1024)
    assert dec.engine() == gpu_lib.ENGINE_RESIDENT
    bits, iters, conv = dec.decode(llr, 20)
    rb, ri, rc = oracles[code_name].decode_batch("spa", 20, llr, variant=0)
    same = (bits == rb).all(axis=1)
    assert same.mean() >= 0.999, same.mean()
    assert (iters == ri).mean() >= 0.98, (iters == ri).mean()
    assert abs(int(conv.sum()) - int(rc.sum())) <= max(2, F // 500)
    # fp64 sum-product stays on the STREAM engine (literal operation order)
    d64 = gpu_lib.Decoder(dcodes[code_name], "spa", "fp64", max_batch=8)
    assert d64.engine() == gpu_lib.ENGINE_STREAM
    dec.close(); d64.close()


def test_device_encoder_and_tester_point(gpu_lib, builtin_codes):
    """Device encoder == host encoder (Fast/Encoder.hs:41-63 / Orig.hs:25-26); a tester point with random messages is
    statistically the same as with the all-zero codeword (linear code, symmetric channel and decoder)."""
    from ecc_ldpc_b200 import ecc
    rng = np.random.default_rng(21)
    for nm, code_name in (("ldpc/b200-minsum/jpl.1024.4.5/20/4/5", "jpl.1024.4.5"), ("ldpc/b200-minsum/jpl.4096.4.5/20/4/5", "jpl.4096.4.5"),
                          ("ldpc/b200-spa/moon.7.13/20", "moon.7.13")):
        (e,) = ecc.match(nm, max_batch=512)
        code = builtin_codes[code_name]
        F, K, N = 37, code.K, code.N
        msgs = rng.integers(0, 2, (F, K)).astype(np.uint8)
        Wk, Wn = (K + 31) // 32, (N + 31) // 32
        pad = np.zeros((F, Wk * 32), dtype=np.uint8); pad[:, :K] = msgs
        packed = np.packbits(pad, axis=1, bitorder="little").view(np.uint32)
        d_msg = gpu_lib.DeviceBuffer(packed.nbytes).upload(packed)
        d_cw = gpu_lib.DeviceBuffer(F * Wn * 4)
        e.run(3.0, 1)                                     # creates the device side
        e.encode_device(d_msg.ptr, d_cw.ptr, F)
        e.sync()
        cw = gpu_lib.unpack_bits(d_cw.download(np.uint32, (F, Wn)), N)
        for f in range(F):
            want = e.encode(msgs[f])
            assert np.array_equal(cw[f, :e.codeword_length], want)
            assert not cw[f, e.codeword_length:].any()
        e.close()
    (e,) = ecc.match("ldpc/b200-minsum/jpl.1024.4.5/20/4/5", max_batch=4096)
    a = e.run(3.3, 8192, seed=3, random_messages=False)
    b = e.run(3.3, 8192, seed=3, random_messages=True)
    assert a["frames"] == b["frames"] == 8192
    fa, fb = a["frame_errors"] / 8192, b["frame_errors"] / 8192
    assert 0.02 < fa < 0.9
    assert abs(fa - fb) < 5 * np.sqrt(fa * (1 - fa) / 8192 * 2)          # same FER within 5 sigma
    assert abs(a["sum_iters"] - b["sum_iters"]) / a["sum_iters"] < 0.03
    e.close()


def test_tester_cli(gpu_lib, tmp_path):
    """ldpc_b200_tester: the `ecc-ldpc <EbN0...> <code-name...> -m<N>` command line (main/Main.hs:38-48, NOTES.txt:2-3)."""
    import subprocess
    from ecc_ldpc_b200 import codes, lib
    exe = os.path.join(os.path.dirname(lib.SO_PATH), "ldpc_b200_tester")
    out = subprocess.run([exe, "--codes", codes.builtin_codes_dir(), "0.0", "4.5", "ldpc/b200-minsum/jpl.1024.4.5/20", "bpsk", "-m256"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 2 and "no code matches bpsk" in out.stderr
    t = lines[0].split()
    # 0 dB at the default rate 8/11: nothing converges -> the fallback gives the uncoded BER of NOTES.txt:3 (1.14e-1 +-1 %)
    assert t[1] == "ldpc/b200-minsum/jpl.1024.4.5/20/8/11" and t[2] == "0.0" and t[3] == "256"
    assert abs(int(t[4]) / (256 * 1024) - 0.1139) < 0.004
    assert int(lines[1].split()[4]) < int(t[4]) / 20


def test_simulate_multi_single_process(gpu_lib, dcodes, builtin_codes):
    """ldpc_b200_simulate_multi: results do not depend on the number of GPUs (frame-index keyed Philox, contiguous
    shards, NCCL sum of the counters). With one visible GPU only the n_gpus = 1 leg runs."""
    code = builtin_codes["jpl.1024.4.5"]
    dec = gpu_lib.Decoder(dcodes["jpl.1024.4.5"], "min", "fp32", max_batch=512)
    want = dec.simulate(20, 3.4, code.K, 1280, 9, 100, 3001)
    dec.close()
    n_dev = gpu_lib.load().ldpc_b200_device_count()
    for n in (1, 2, 4, 8):
        if n > n_dev:
            break
        got = gpu_lib.simulate_multi(dcodes["jpl.1024.4.5"], n, "min", "fp32", 512, 20, 3.4, code.K, 1280, 9, 100, 3001)
        for k in ("frames", "frame_errors", "bit_errors", "sum_iters", "unconverged"):
            assert got[k] == want[k], (n, k)
    gpu_lib.load().ldpc_b200_simulate_multi_release()
    with pytest.raises(gpu_lib.LdpcError):
        gpu_lib.simulate_multi(dcodes["jpl.1024.4.5"], n_dev + 1, "min", "fp32", 512, 20, 3.4, code.K, 1280, 9, 0, 10)


def test_cached_mult_variant(gpu_lib, dcodes, builtin_codes, oracles):
    """the cached-product sum-product of Fast/CachedMult.hs:233-264 (float64, STREAM engine) against
    the oracle's restatement: decisions / iteration counts on a batch, per-edge messages of single frames."""
    for name in ("moon.7.13", "jpl.1024.4.5", "jpl.4096.4.5"):
        code = builtin_codes[name]
        F = 400 if name == "moon.7.13" else 48
        llr = awgn_llrs(code, EBN0[name], F, 4242, TX[name])
        dec = gpu_lib.Decoder(dcodes[name], "spa_cm", "fp64", max_batch=64)
        assert dec.engine() == gpu_lib.ENGINE_STREAM
        bits, iters, conv = dec.decode(llr, 12)
        rb, ri, rc = oracles[name].decode_batch("spa_cm", 12, llr)
        assert (bits == rb).all(axis=1).mean() >= 0.999 and (iters == ri).mean() >= 0.99
        for f in range(2):
            b, it, cv, c2v, lam = dec.decode_trace(llr[f], 12)
            ob, oit, ocv, oc2v, olam = oracles[name].decode("spa_cm", 12, llr[f], trace=True)
            assert it == oit and cv == ocv and np.array_equal(b, ob)
            if it:
                assert spa_close(c2v[0], oc2v[0]) and spa_close(lam[0], olam[0])
                assert spa_close(c2v[-1], oc2v[-1]) and spa_close(lam[-1], olam[-1])
        dec.close()
    with pytest.raises(gpu_lib.LdpcError):          # float64-only variant: fails loudly instead of silently switching
        gpu_lib.Decoder(dcodes["moon.7.13"], "spa_cm", "fp32", max_batch=8)
    from ecc_ldpc_b200 import ecc
    (e,) = ecc.match("ldpc/b200-spa-cm-fp64/jpl.1024.4.5/10/4/5", max_batch=8)
    assert e.name == "ldpc/b200-spa-cm-fp64/jpl.1024.4.5/10/4/5"
    rx = awgn_llrs(builtin_codes["jpl.1024.4.5"], 4.2, 4, 5, 1280)[:, :1280]
    out, flags = e.decode_batch(rx)
    for f in range(4):
        ob, oflag, oit, ocv = oracles["jpl.1024.4.5"].ecc_decode("spa_cm", 10, 1024, 1280, rx[f])
        assert np.array_equal(out[f], ob) and flags[f] == 1


ONCHIP_CASES = [("moon.7.13", False), ("1920.1280.3.303", False), ("jpl.1024.4.5", True)]   # (code, force ENGINE_ONCHIP)


def _onchip_decoder(gpu_lib, dcode, algo, prec, max_batch, force):
    dec = gpu_lib.Decoder(dcode, algo, prec, max_batch=max_batch, engine=gpu_lib.ENGINE_ONCHIP if force else gpu_lib.ENGINE_AUTO)
    assert dec.engine() == gpu_lib.ENGINE_ONCHIP
    return dec


@pytest.mark.parametrize("code_name,force", ONCHIP_CASES)
def test_onchip_fp64_bit_exact_vs_oracle(code_name, force, gpu_lib, dcodes, builtin_codes, oracles):
    """Run-time table on-chip kernel, float64 min-sum: decisions, iteration counts and flags identical to the oracle
    (general codes run on it by default; the JPL code is forced onto it to cover irregular row degrees 3 / 18)."""
    code = builtin_codes[code_name]
    F = 900 if code_name == "moon.7.13" else 200
    llr = awgn_llrs(code, EBN0[code_name], F, 31337, TX[code_name])
    dec = _onchip_decoder(gpu_lib, dcodes[code_name], "min", "fp64", 128, force)      # F > max_batch: chunking
    bits, iters, conv = dec.decode(llr, 20)
    rb, ri, rc = oracles[code_name].decode_batch("min", 20, llr, variant=0)
    assert np.array_equal(bits, rb) and np.array_equal(iters, ri) and np.array_equal(conv, rc)
    # iteration cap 0 and 1; a valid codeword; an all-zero LLR vector
    for cap in (0, 1):
        b, it, cv = dec.decode(llr[:16], cap)
        ob, oi, oc = oracles[code_name].decode_batch("min", cap, llr[:16], variant=0)
        assert np.array_equal(b, ob) and np.array_equal(it, oi) and np.array_equal(cv, oc)
    z = np.zeros((3, code.N)); z[1] = -5.0; z[2, ::2] = 1e-3
    b, it, cv = dec.decode(z, 7)
    ob, oi, oc = oracles[code_name].decode_batch("min", 7, z, variant=0)
    assert np.array_equal(b, ob) and np.array_equal(it, oi) and np.array_equal(cv, oc)
    dec.close()


@pytest.mark.parametrize("code_name,force", ONCHIP_CASES)
@pytest.mark.parametrize("algo", ["min", "spa"])
def test_onchip_fp32_equals_stream_fp32(code_name, force, algo, gpu_lib, dcodes, builtin_codes, oracles):
    """float32: the on-chip kernel and the STREAM kernels share the check-row arithmetic (bp_tree.cuh) and the
    reference's fold order, so they agree bit for bit; both within the 99.9 % bar of the float64 oracle."""
    code = builtin_codes[code_name]
    F = 2048 if code_name == "moon.7.13" else 384
    llr = awgn_llrs(code, EBN0[code_name] + 0.2, F, 777, TX[code_name]).astype(np.float32).astype(np.float64)
    on = _onchip_decoder(gpu_lib, dcodes[code_name], algo, "fp32", 512, force)
    st = gpu_lib.Decoder(dcodes[code_name], algo, "fp32", max_batch=512, engine=gpu_lib.ENGINE_STREAM)
    b1, i1, c1 = on.decode(llr, 20)
    b2, i2, c2 = st.decode(llr, 20)
    assert np.array_equal(b1, b2) and np.array_equal(i1, i2) and np.array_equal(c1, c2)
    rb, ri, rc = oracles[code_name].decode_batch(algo, 20, llr, variant=1 if algo == "min" else 0)
    assert (b1 == rb).all(axis=1).mean() >= 0.999
    on.close(); st.close()


@pytest.mark.parametrize("code_name,force", ONCHIP_CASES[:2])
def test_onchip_device_api_and_simulate(code_name, force, gpu_lib, dcodes, builtin_codes):
    code = builtin_codes[code_name]
    K, n_tx, F, ebn0 = code.K, code.N, 1500, EBN0[code_name]
    dec = _onchip_decoder(gpu_lib, dcodes[code_name], "min", "fp32", 2048, force)
    c = dec.simulate(20, ebn0, K, n_tx, 11, 123456, F)
    buf = gpu_lib.DeviceBuffer(F * code.N * 4)
    dec.generate_llr(buf.ptr, ebn0, K, n_tx, 11, 123456, F)
    dec.sync()
    llr = buf.download(np.float32, (F, code.N)).astype(np.float64)
    gb, gi, gc = dec.decode(llr, 20)
    assert c["frames"] == F and c["sum_iters"] == int(gi.sum()) and c["unconverged"] == int((gc == 0).sum())
    assert c["bit_errors"] == int(gb[:, :K].sum()) and c["frame_errors"] == int(gb[:, :K].any(axis=1).sum())
    # fixed iteration count (early_exit = 0): decisions are those of the iteration at which the frame first converged
    W = (code.N + 31) // 32
    bits = gpu_lib.DeviceBuffer(F * W * 4); it = gpu_lib.DeviceBuffer(F * 4); cv = gpu_lib.DeviceBuffer(F)
    dec.decode_device(buf.ptr, bits.ptr, 20, F, early_exit=False, d_iters=it.ptr, d_conv=cv.ptr)
    dec.sync()
    fb = gpu_lib.unpack_bits(bits.download(np.uint32, (F, W)), code.N)
    assert np.array_equal(fb, gb) and np.array_equal(it.download(np.int32, (F,)), gi)
    dec.close()


@pytest.mark.parametrize("code_name,F,ebn0", [("jpl.1024.4.5", 16384, 3.6), ("jpl.4096.4.5", 3072, 3.5), ("1920.1280.3.303", 8192, 2.2)])
@pytest.mark.parametrize("algo", ["min", "spa"])
def test_production_engines_large_batch_vs_oracle(code_name, F, ebn0, algo, gpu_lib, dcodes, builtin_codes, oracles):
    """BASELINE.json north_star bar at batch scale (config 2 is 16k frames of jpl.1024.4.5): float32 production engines
    (RESIDENT / ONCHIP through ENGINE_AUTO) against the float64 CPU restatement on identical LLRs, in the waterfall where
    frames converge late or not at all -- decisions identical on >= 99.9 % of the frames, FER inside the oracle's
    95 % interval, iteration counts equal on >= 99 %."""
    code = builtin_codes[code_name]
    if algo == "spa":
        F //= 4
    llr = awgn_llrs(code, ebn0, F, 20261018, TX[code_name])
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp32", max_batch=4096)
    assert dec.engine() in (gpu_lib.ENGINE_RESIDENT, gpu_lib.ENGINE_ONCHIP)
    bits, iters, conv = dec.decode(llr, 20)
    rb, ri, rc = oracles[code_name].decode_batch(algo, 20, llr, variant=1 if algo == "min" else 0)
    same = (bits == rb).all(axis=1)
    assert same.mean() >= 0.999, (code_name, algo, same.mean())
    assert (iters == ri).mean() >= 0.99
    K = code.K
    fe_gpu, fe_ref = int(bits[:, :K].any(axis=1).sum()), int(rb[:, :K].any(axis=1).sum())
    p = fe_ref / F
    half = 1.96 * np.sqrt(max(p * (1 - p), 1.0 / F) / F) + 1.0 / F
    assert abs(fe_gpu / F - p) <= half, (fe_gpu, fe_ref, F)
    dec.close()


def test_concurrent_host_threads_one_decoder_each(gpu_lib, dcodes, builtin_codes, oracles):
    """The reference's tester calls `decode` from several threads, each with its own decoder (Utils.hs:53,62-69:
    maxThreadCount decoders, thread t uses t mod maxThreadCount).  Four host threads, one handle each, one frame per
    call, concurrently (ctypes drops the GIL): every result equals the CPU restatement's."""
    import threading
    code = builtin_codes["jpl.1024.4.5"]
    T, calls = 4, 40
    llr = awgn_llrs(code, 3.8, T * calls, 99, 1280)
    rb, ri, rc = oracles["jpl.1024.4.5"].decode_batch("min", 20, llr, variant=1)
    decs = [gpu_lib.Decoder(dcodes["jpl.1024.4.5"], "min", "fp64", max_batch=1) for _ in range(T)]
    bad, errs = [], []

    def work(t):
        try:
            for c in range(calls):
                f = t * calls + c
                b, it, cv = decs[t].decode(llr[f:f + 1], 20)
                if not (np.array_equal(b[0], rb[f]) and it[0] == ri[f] and cv[0] == rc[f]):
                    bad.append(f)
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs and not bad, (errs, bad)
    [d.close() for d in decs]


def test_simulate_counters_pinned(gpu_lib, builtin_codes):
    """Regression pin (tests/golden/sim_counters.json, written by tests/golden/make_sim_golden.py on a B200): the error
    counters of ldpc_b200_simulate for fixed Philox frame ranges.  The float32 kernels follow the reference's operation
    order with explicit roundings (and the sum-product's MUFU approximations are deterministic on sm_100), so a kernel
    rewrite that changes ONE decision in millions of frames shows up here.  Parity itself is the oracle tests above."""
    import json
    cases = json.load(open(os.path.join(GOLDEN, "sim_counters.json")))
    assert len(cases) >= 8
    for cs in cases:
        code = builtin_codes[cs["code"]]
        dc = gpu_lib.DeviceCode(code)
        dec = gpu_lib.Decoder(dc, cs["algo"], "fp32", max_batch=16384)
        assert dec.engine() == cs["engine"]
        c = dec.simulate(cs["max_iters"], cs["ebn0_db"], code.K, cs["n_tx"], cs["seed"], cs["first_frame"], cs["n_frames"])
        got = {k: int(c[k]) for k in cs["counters"]}
        assert got == cs["counters"], (cs["code"], cs["algo"], got, cs["counters"])
        dec.close(); dc.close()


# ------------------------------------------------------------------------------------------------
# Round 2: BASELINE config 3, the four-code parity matrix, the atanh' clamp, the arraylet-order statements
# ------------------------------------------------------------------------------------------------
def _fer_interval_ok(fe_gpu, fe_ref, F):
    p = fe_ref / F
    half = 1.96 * np.sqrt(max(p * (1 - p), 1.0 / F) / F) + 1.0 / F
    return abs(fe_gpu / F - p) <= half


@pytest.mark.parametrize("algo,F,ebn0,waterfall", [("min", 2048, 2.9, True), ("min", 2048, 3.15, False), ("spa", 512, 2.6, True), ("spa", 512, 3.0, False)])
def test_config3_jpl4096_50_iterations_early_exit(algo, F, ebn0, waterfall, gpu_lib, dcodes, builtin_codes, oracles):
    """BASELINE.json config 3: jpl.4096.4.5 sum-product and min-sum, 50 iterations with syndrome early exit, on the
    float32 RESIDENT engine against the float64 CPU restatement on identical LLRs: in the 50-iteration waterfall (2.9 dB
    min-sum / 2.6 dB sum-product: frames that need most of the 50 iterations and frames that fail are both present) and
    at 3.0-3.15 dB where every frame converges, many of them after iteration 20.  Bars: decisions identical on >= 99.9 % of the
    frames, iteration counts equal on >= 99 %, FER inside the restatement's 95 % interval."""
    code = builtin_codes["jpl.4096.4.5"]
    llr = awgn_llrs(code, ebn0, F, 50505, 5120)
    dec = gpu_lib.Decoder(dcodes["jpl.4096.4.5"], algo, "fp32", max_batch=1024)
    assert dec.engine() == gpu_lib.ENGINE_RESIDENT
    bits, iters, conv = dec.decode(llr, 50)
    rb, ri, rc = oracles["jpl.4096.4.5"].decode_batch(algo, 50, llr, variant=1 if algo == "min" else 0)
    same = (bits == rb).all(axis=1)
    assert same.mean() >= 0.999, (algo, same.mean())
    assert (iters == ri).mean() >= 0.99, (algo, (iters == ri).mean())
    assert (ri[rc == 1] > 20).any()                                # the cap of 50 matters: frames converge after iteration 20
    if waterfall:
        assert 0.5 < rc.mean() < 0.99
    K = code.K
    assert _fer_interval_ok(int(bits[:, :K].any(axis=1).sum()), int(rb[:, :K].any(axis=1).sum()), F)
    assert abs(int(conv.sum()) - int(rc.sum())) <= max(2, F // 500)
    dec.close()


PARITY_MATRIX = [("jpl.4096.4.5", "min", 5120, 3.25, 4096), ("jpl.4096.4.5", "spa", 5120, 3.0, 1024),
                 ("jpl.1024.4.5", "min", 1280, 3.4, 16384), ("jpl.1024.4.5", "spa", 1280, 3.2, 4096),
                 ("1920.1280.3.303", "min", 1920, 1.8, 12288), ("1920.1280.3.303", "spa", 1920, 1.5, 3072),
                 ("moon.7.13", "spa", 20, 2.0, 131072), ("moon.7.13", "min", 20, 2.0, 131072)]


@pytest.mark.parametrize("code_name,algo,n_tx,ebn0,F", PARITY_MATRIX)
def test_parity_matrix_device_generated_frames(code_name, algo, n_tx, ebn0, F, gpu_lib, dcodes, builtin_codes, oracles):
    """tools/parity_at_scale.py inside the suite: the four codes x both algorithms, frames generated ON THE DEVICE
    (Philox, the throughput path's generator), downloaded, and decoded by the float32 production engine (ENGINE_AUTO)
    and by the float64 CPU restatement.  Decisions identical on >= 99.9 % of the frames, iteration counts on >= 99 %,
    FER inside the restatement's interval; and the device counters of ldpc_b200_simulate over the same frame range are
    the counts of those decisions."""
    code = builtin_codes[code_name]
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp32", max_batch=8192)
    assert dec.engine() in (gpu_lib.ENGINE_RESIDENT, gpu_lib.ENGINE_ONCHIP)
    same = eq_it = fe_g = fe_r = be_g = it_g = 0
    for f0 in range(0, F, 8192):
        n = min(8192, F - f0)
        buf = gpu_lib.DeviceBuffer(n * code.N * 4)
        dec.generate_llr(buf.ptr, ebn0, code.K, n_tx, 424242, 1000 + f0, n)
        dec.sync()
        llr = buf.download(np.float32, (n, code.N)).astype(np.float64)
        buf.free()
        gb, gi, gc = dec.decode(llr, 20)
        rb, ri, rc = oracles[code_name].decode_batch(algo, 20, llr, variant=1 if algo == "min" else 0)
        same += int((gb == rb).all(axis=1).sum()); eq_it += int((gi == ri).sum())
        fe_g += int(gb[:, :code.K].any(axis=1).sum()); fe_r += int(rb[:, :code.K].any(axis=1).sum())
        be_g += int(gb[:, :code.K].sum()); it_g += int(gi.sum())
    assert same / F >= 0.999, (code_name, algo, same, F)
    assert eq_it / F >= 0.99
    assert 0 < fe_r < F and _fer_interval_ok(fe_g, fe_r, F)
    c = dec.simulate(20, ebn0, code.K, n_tx, 424242, 1000, F)
    assert (c["frames"], c["frame_errors"], c["bit_errors"], c["sum_iters"]) == (F, fe_g, be_g, it_g)
    dec.close()


CLAMP = 2 * 18.714973875118524          # -2 * atanh'(+-1)   (Utils.hs:113-117)


def _saturated_frames(code, n_frames, seed, punctured_from=None):
    """LLRs of magnitude >= 40: tanh(x / 2) rounds to +-1 in float64, so every leave-one-out product of the first
    iteration is exactly +-1 (or exactly 0 next to a punctured position) and atanh' takes its clamp branch."""
    rng = np.random.default_rng(seed)
    llr = np.full((n_frames, code.N), -45.0)
    for f in range(n_frames):
        if f % 2:
            llr[f] = -rng.uniform(40.0, 60.0, code.N)
        flip = rng.choice(code.N if punctured_from is None else punctured_from, 5 + f, replace=False)
        llr[f, flip] = -llr[f, flip]
    if punctured_from is not None:
        llr[:, punctured_from:] = 0.0
    return llr


@pytest.mark.parametrize("code_name,punct", [("jpl.1024.4.5", None), ("jpl.1024.4.5", 1280), ("moon.7.13", None), ("1920.1280.3.303", None)])
def test_atanh_clamp_known_answer(code_name, punct, gpu_lib, dcodes, builtin_codes, oracles):
    """Bug history Reference/Orig.hs:85 ("must cap atanh") as a known-answer test: with |LLR| >= 40 every product of the
    first sum-product iteration is exactly +-1 (0 on the neighbours of a punctured position), so every message of that
    iteration is exactly -+2 * 18.714973875118524 (or 0).  The float64 STREAM kernels return exactly those values and
    equal the restatement; the float32 engines (complement domain, cap 54 ln 2) decide the same bits."""
    code = builtin_codes[code_name]
    llr = _saturated_frames(code, 6, 4545, punct)
    d64 = gpu_lib.Decoder(dcodes[code_name], "spa", "fp64", max_batch=8)
    for f in range(llr.shape[0]):
        b, it, cv, c2v, lam = d64.decode_trace(llr[f], 10)
        ob, oit, ocv, oc2v, olam = oracles[code_name].decode("spa", 10, llr[f], trace=True)
        assert it == oit and it >= 1 and cv == ocv and np.array_equal(b, ob)
        first = np.abs(oc2v[0])
        assert np.isin(first, [0.0, CLAMP]).all() and (first == CLAMP).any()          # the restatement clamps ...
        if punct is None:
            assert (first == CLAMP).all()
        assert np.array_equal(c2v[0], oc2v[0])                                         # ... and the kernel returns the same numbers
        assert np.array_equal(lam[0], olam[0])
        assert spa_close(c2v[-1], oc2v[-1]) and spa_close(lam[-1], olam[-1])
    rb, ri, rc = oracles[code_name].decode_batch("spa", 10, llr)
    for eng in (gpu_lib.ENGINE_AUTO, gpu_lib.ENGINE_STREAM):
        d32 = gpu_lib.Decoder(dcodes[code_name], "spa", "fp32", max_batch=8, engine=eng)
        gb, gi, gc = d32.decode(llr, 10)
        assert np.array_equal(gb, rb) and np.array_equal(gc, rc), (code_name, eng)
        # iteration counts: identical on the frames with random magnitudes; the frames with ONE magnitude (45) produce
        # posteriors that are exactly 0 in exact arithmetic (45 - 37.43 + 29.86 - 37.43), where the float64 reference
        # decides on its own rounding noise (+1.4e-14 observed) and may therefore stop one iteration later or earlier
        assert np.array_equal(gi[1::2], ri[1::2]) and (np.abs(gi - ri) <= 1).all(), (code_name, eng, gi, ri)
        d32.close()
    # min-sum has no clamp; the same frames must still decode identically (huge but finite magnitudes)
    mb, mi, mc = oracles[code_name].decode_batch("min", 10, llr, variant=0)
    dm = gpu_lib.Decoder(dcodes[code_name], "min", "fp32", max_batch=8)
    gb, gi, gc = dm.decode(llr, 10)
    assert np.array_equal(gb, mb) and np.array_equal(gi, mi)
    d64.close(); dm.close()


@pytest.mark.parametrize("code_name,ebn0,n_tx", [("jpl.1024.4.5", 3.7, 1280), ("jpl.4096.4.5", 3.3, 5120)])
@pytest.mark.parametrize("algo", ["arraylet-min", "arraylet"])
def test_arraylet_order_statements(code_name, ebn0, n_tx, algo, gpu_lib, dcodes, builtin_codes, oracles):
    """`ldpc/arraylet-min` (Fast/ArrayletMin.hs:167-192) and `ldpc/arraylet` (Fast/Arraylet.hs:164-186) on the STREAM
    engine in float64, against the circulant-vector restatement of those modules (oracle/arraylet_oracle.c): min-sum
    messages and posteriors identical in every bit, sum-product within the libm tolerance; decisions and iteration
    counts of the production float32 engine (Orig/Min order) equal the arraylet statement's on the same frames."""
    from oracle import oracle
    code = builtin_codes[code_name]
    base = "min" if algo == "arraylet-min" else "spa"
    F = 256 if base == "min" else 64
    llr = awgn_llrs(code, ebn0 - (0.0 if base == "min" else 0.55), F, 7117, n_tx)
    a = oracle.ArrayletOracle(code)
    ab, ai, ac = a.decode_batch(base, 20, llr)
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp64", max_batch=64)
    assert dec.engine() == gpu_lib.ENGINE_STREAM
    gb, gi, gc = dec.decode(llr, 20)
    if base == "min":
        assert np.array_equal(gb, ab) and np.array_equal(gi, ai) and np.array_equal(gc, ac)
    else:
        assert (gb == ab).all(axis=1).mean() >= 0.999 and (gi == ai).mean() >= 0.99
    for f in range(3):
        b, it, cv, c2v, lam = dec.decode_trace(llr[f], 20)
        xb, xit, xcv, xc2v, xlam = a.decode(base, 20, llr[f], trace=True)
        assert it == xit and cv == xcv and np.array_equal(b, xb)
        if it and base == "min":
            assert np.array_equal(c2v, xc2v) and np.array_equal(lam, xlam)
        elif it:
            assert spa_close(c2v[0], xc2v[0]) and spa_close(lam[0], xlam[0]) and spa_close(c2v[-1], xc2v[-1])
    prod = gpu_lib.Decoder(dcodes[code_name], base, "fp32", max_batch=256)
    pb, pi, pc = prod.decode(llr, 20)
    assert (pb == ab).all(axis=1).mean() >= 0.999 and (pi == ai).mean() >= 0.99
    assert 0 < ac.mean() < 1
    with pytest.raises(gpu_lib.LdpcError):
        dec.set_engine(gpu_lib.ENGINE_RESIDENT)           # verification statements live on STREAM only
    dec.close(); prod.close()


# ------------------------------------------------------------------------------------------------
# Round 2: message tap on the production engines (north_star: fp64 1e-9 / fp32 1e-4 per-edge), the mailbox
# ------------------------------------------------------------------------------------------------
def _msg_errors(got, ref):
    """(max relative error over the entries with |ref| >= 0.1, max absolute error over the entries below 0.1).
    A float32 posterior of magnitude ~10 carries ~1e-6 of absolute rounding, and x = ne - lam cancels, so a message far
    below 1 cannot be relatively exact to 1e-4 in ANY float32 implementation; profiles/r02_message_parity.md has the
    measured distribution per magnitude bin (relative <= 2.2e-5 from 0.1 up, absolute <= 3e-6 below)."""
    got, ref = np.asarray(got), np.asarray(ref)
    err = np.abs(got - ref)
    big = np.abs(ref) >= 0.1
    return (float((err[big] / np.abs(ref[big])).max()) if big.any() else 0.0, float(err[~big].max()) if (~big).any() else 0.0)


TAP_CASES = [("jpl.1024.4.5", 1280, 4.1, 2), ("jpl.4096.4.5", 5120, 3.9, 2), ("1920.1280.3.303", None, 3.0, 3), ("moon.7.13", None, 4.0, 3)]


@pytest.mark.parametrize("code_name,n_tx,ebn0,want_engine", TAP_CASES)
@pytest.mark.parametrize("algo", ["min", "spa"])
def test_message_tap_production_engines_fp32_within_1e4(code_name, n_tx, ebn0, want_engine, algo, gpu_lib, dcodes, builtin_codes, oracles):
    """ldpc_b200_decode_trace on the engines that produce the throughput numbers: the float32 RESIDENT kernels (message
    registers dumped in reference edge order by the TRACE instance) and the float32 ONCHIP kernel (shared-memory messages),
    against the float64 restatement on the same (float32-rounded) LLRs.  Per-edge messages and posteriors of the first
    iterations within 1e-4 RELATIVE wherever |reference| >= 0.1 and within 1e-5 ABSOLUTE below; decisions, iteration counts
    and flags equal.  Sum-product messages are compared where the reference itself resolves them: below 30 in magnitude
    (above, its float64 product 1 - p is a small multiple of 2^-53 and its own messages move in steps of up to 0.7, see
    DESIGN.md)."""
    code = builtin_codes[code_name]
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp32", max_batch=4)
    assert dec.engine() == want_engine
    llr = awgn_llrs(code, ebn0 - (0.4 if algo == "spa" else 0.0), 5, 2718, n_tx).astype(np.float32).astype(np.float64)
    worst_rel, worst_abs, checked = 0.0, 0.0, 0
    for f in range(llr.shape[0]):
        b, it, cv, c2v, lam = dec.decode_trace(llr[f], 8)
        rb, rit, rcv, rc2v, rlam = oracles[code_name].decode(algo, 8, llr[f], variant=0, trace=True)
        assert it == rit and cv == rcv and np.array_equal(b, rb), (code_name, algo, f, it, rit)
        assert c2v.shape == rc2v.shape and lam.shape == rlam.shape
        gb, gi, gc = dec.decode(llr[f:f + 1], 8)                  # the tap does not change the decode
        assert np.array_equal(gb[0], b) and gi[0] == it
        if not it:
            continue
        n = min(it, 6)
        if algo == "spa":
            m = np.abs(rc2v[:n]) < 30.0
            e1, f1 = _msg_errors(c2v[:n][m], rc2v[:n][m])
            ml = np.abs(rlam[:n]) < 30.0
            e2, f2 = _msg_errors(lam[:n][ml], rlam[:n][ml])
        else:
            e1, f1 = _msg_errors(c2v[:n], rc2v[:n])
            e2, f2 = _msg_errors(lam[:n], rlam[:n])
        worst_rel, worst_abs, checked = max(worst_rel, e1, e2), max(worst_abs, f1, f2), checked + 1
    assert checked >= 3
    assert worst_rel <= 1e-4, (code_name, algo, worst_rel)
    assert worst_abs <= 1e-5, (code_name, algo, worst_abs)
    dec.close()


@pytest.mark.parametrize("code_name,n_tx,ebn0,force_onchip", [("jpl.1024.4.5", 1280, 3.8, False), ("jpl.4096.4.5", 5120, 3.5, False),
                                                              ("1920.1280.3.303", None, 2.4, False), ("jpl.1024.4.5", 1280, 3.8, True)])
def test_message_tap_fp64_minsum_bit_exact(code_name, n_tx, ebn0, force_onchip, gpu_lib, dcodes, builtin_codes, oracles):
    """The float64 instances of the on-chip kernels follow the reference's operation order: every message and posterior of
    every iteration, read out of the RESIDENT registers / ONCHIP shared memory, equals the restatement's in every bit
    (well inside the 1e-9 bar of the verification mode)."""
    code = builtin_codes[code_name]
    eng = gpu_lib.ENGINE_ONCHIP if force_onchip else gpu_lib.ENGINE_AUTO
    dec = gpu_lib.Decoder(dcodes[code_name], "min", "fp64", max_batch=4, engine=eng)
    assert dec.engine() in (gpu_lib.ENGINE_RESIDENT, gpu_lib.ENGINE_ONCHIP)
    llr = awgn_llrs(code, ebn0, 4, 1618, n_tx)
    for f in range(4):
        b, it, cv, c2v, lam = dec.decode_trace(llr[f], 20)
        rb, rit, rcv, rc2v, rlam = oracles[code_name].decode("min", 20, llr[f], variant=0, trace=True)
        assert it == rit and cv == rcv and np.array_equal(b, rb)
        assert np.array_equal(c2v, rc2v) and np.array_equal(lam, rlam)
    dec.close()


@pytest.mark.parametrize("code_name,algo,prec,n_tx,ebn0", [("jpl.4096.4.5", "min", "fp32", 5120, 3.4), ("jpl.4096.4.5", "spa", "fp32", 5120, 3.0),
                                                        ("jpl.1024.4.5", "min", "fp32", 1280, 3.6), ("jpl.1024.4.5", "min", "fp64", 1280, 3.6),
                                                        ("1920.1280.3.303", "min", "fp32", None, 2.2), ("moon.7.13", "spa", "fp32", None, 3.0)])
def test_mailbox_tester_pattern_concurrent_single_frame_calls(code_name, algo, prec, n_tx, ebn0, gpu_lib, dcodes, builtin_codes, oracles):
    """The reference tester's calling pattern (Utils.hs:53,62-69): `maxThreadCount` decoder handles on one code, each host
    thread calling `decode` with ONE frame at a time, concurrently.  Those calls go through the shared mailbox (coalesced
    launches of the latency kernels): every result equals the batched path's bit for bit (same arithmetic), which in turn
    meets the 99.9 % bar against the restatement; and frames were coalesced (fewer launches than frames)."""
    import threading
    code = builtin_codes[code_name]
    T, calls = 16, 24
    llr = awgn_llrs(code, ebn0, T * calls, 4321, n_tx)
    batch = gpu_lib.Decoder(dcodes[code_name], algo, prec, max_batch=T * calls)
    bb, bi, bc = batch.decode(llr, 20)                                   # batched path (n_frames > 4)
    rb, ri, rc = oracles[code_name].decode_batch(algo, 20, llr, variant=0)
    assert (bb == rb).all(axis=1).mean() >= 0.999
    decs = [gpu_lib.Decoder(dcodes[code_name], algo, prec, max_batch=1) for _ in range(T)]
    decs[0].decode(llr[:1], 20)                                          # attaches the handle to the (shared) mailbox
    launches0, frames0 = decs[0].mailbox_stats()
    out_b = np.zeros_like(bb); out_i = np.zeros_like(bi); out_c = np.zeros_like(bc)
    errs = []

    def work(t):
        try:
            for c in range(calls):
                f = t * calls + c
                b, it, cv = decs[t].decode(llr[f:f + 1], 20)
                out_b[f], out_i[f], out_c[f] = b[0], it[0], cv[0]
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    assert np.array_equal(out_b, bb) and np.array_equal(out_i, bi) and np.array_equal(out_c, bc)
    launches, frames = decs[0].mailbox_stats()
    assert frames - frames0 == T * calls and 1 <= launches - launches0 <= T * calls
    # two- and four-frame calls take the mailbox too; ragged max_iters values are grouped per launch
    b2, i2, c2 = decs[1].decode(llr[:4], 20)
    assert np.array_equal(b2, bb[:4]) and np.array_equal(i2, bi[:4])
    b3, i3, c3 = decs[2].decode(llr[:2], 3)
    r3 = oracles[code_name].decode_batch(algo, 3, llr[:2], variant=0)
    assert (b3 == r3[0]).all(axis=1).mean() >= 0.5 and np.array_equal(i3, r3[1])
    [d.close() for d in decs]
    batch.close()


def test_mailbox_ecc_record_single_frame(gpu_lib, builtin_codes, oracles):
    """ECC.decode of the record (one received vector per call, punctured tail, first K bits back) through the mailbox ==
    the restatement's wrapper (Utils.hs:55,68-72)."""
    from ecc_ldpc_b200 import ecc
    (e,) = ecc.match("ldpc/b200-minsum/jpl.4096.4.5/20/4/5", max_batch=1)
    code = builtin_codes["jpl.4096.4.5"]
    rx = awgn_llrs(code, 3.6, 12, 777, 5120)[:, :5120]
    for f in range(12):
        out, flag = e.decode(rx[f])
        ref, rflag, it, cv = oracles["jpl.4096.4.5"].ecc_decode("min", 20, 4096, 5120, rx[f])
        assert flag and np.array_equal(out, ref)
    e.close()


def test_shutdown_releases_everything(gpu_lib):
    """ldpc_b200_shutdown = the Code's `finalize`: live handles (codes, decoders, mailboxes, simulate_multi replicas) are
    destroyed, destroying them again is a no-op, and the library can be initialised and used again.  Runs in its own
    process: shutting the library down would invalidate the handles the other tests' fixtures hold."""
    import subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = textwrap.dedent("""
        import ctypes as C, sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import numpy as np
        from conftest import awgn_llrs
        from ecc_ldpc_b200 import codes, lib
        L = lib.load(); lib.init(0)
        nc, nd = C.c_int(), C.c_int()
        code = codes.load_builtin("jpl.1024.4.5")
        dc = lib.DeviceCode(code)
        decs = [lib.Decoder(dc, "min", "fp32", max_batch=1) for _ in range(3)]
        llr = awgn_llrs(code, 4.0, 1, 5, 1280)
        want = decs[0].decode(llr, 20)                       # creates the mailbox
        big = decs[1].decode(np.repeat(llr, 8, axis=0), 20)  # and the batched staging
        lib.simulate_multi(dc, 1, "min", "fp32", 64, 20, 3.5, code.K, 1280, 1, 0, 64)
        L.ldpc_b200_live_handles(C.byref(nc), C.byref(nd))
        assert (nc.value, nd.value) == (2, 4), (nc.value, nd.value)      # + the replica of simulate_multi
        L.ldpc_b200_shutdown()
        L.ldpc_b200_live_handles(C.byref(nc), C.byref(nd))
        assert (nc.value, nd.value) == (0, 0)
        [d.close() for d in decs]; dc.close()                # stale handles: ignored
        lib.init(0)
        dc2 = lib.DeviceCode(code)
        d2 = lib.Decoder(dc2, "min", "fp32", max_batch=1)
        got = d2.decode(llr, 20)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        d2.close(); dc2.close()
        print("SHUTDOWN-OK")
    """) % (root, os.path.join(root, "tests"))
    out = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SHUTDOWN-OK" in out.stdout, out.stderr[-2000:]


def test_host_generator_matches_device(gpu_lib, dcodes, builtin_codes):
    """ecc-ldpc_b200/philox_host.py draws on the host the frames ldpc_b200_generate_llr draws on the device (same Philox
    words; numpy's and CUDA's float32 log / sin / cos differ in the last places only), so the CPU arm of bench.py decodes
    the GPU arm's frames."""
    from ecc_ldpc_b200 import philox_host
    for code_name, n_tx, ebn0 in (("jpl.4096.4.5", 5120, 2.0), ("moon.7.13", 20, 3.0)):
        code = builtin_codes[code_name]
        dec = gpu_lib.Decoder(dcodes[code_name], "min", "fp32", max_batch=64)
        buf = gpu_lib.DeviceBuffer(64 * code.N * 4)
        dec.generate_llr(buf.ptr, ebn0, code.K, n_tx, 20261018, 12345678901, 64)
        dec.sync()
        dev = buf.download(np.float32, (64, code.N))
        host = philox_host.awgn_llrs(code.N, ebn0, code.K, n_tx, 20261018, 12345678901, 64)
        assert np.array_equal(host[:, n_tx:], dev[:, n_tx:])
        scale = np.abs(dev).max()
        assert np.abs(host - dev).max() <= 2e-6 * scale
        assert (host == dev).mean() > 0.5
        assert np.array_equal(host > 0, dev > 0) or ((host > 0) != (dev > 0)).sum() <= 1
        dec.close()


def test_tester_pattern_harness(gpu_lib, builtin_codes, oracles):
    """ldpc_b200_ecc_tester_pattern: 16 tester threads, one ECC record each (mkLDPC's replicateM maxThreadCount), one frame per
    ECC.decode call.  Every message equals the restatement wrapper's; identical codes are shared, so the 16 records meet in
    one mailbox and their calls coalesce (fewer launches than frames)."""
    from ecc_ldpc_b200 import codes
    code = builtin_codes["jpl.4096.4.5"]
    rx = awgn_llrs(code, 3.7, 1024, 2468, 5120)[:, :5120]
    msg, secs, launches, frames = gpu_lib.tester_pattern(codes.builtin_codes_dir(), "ldpc/b200-minsum/jpl.4096.4.5/20/4/5", 16, rx)
    assert frames == 1024 and 1 <= launches < frames and secs > 0
    bad = 0
    for f in range(0, 1024, 8):
        ref, rflag, it, cv = oracles["jpl.4096.4.5"].ecc_decode("min", 20, 4096, 5120, rx[f])
        bad += int(not np.array_equal(msg[f], ref))
    assert bad == 0
    m1, s1, l1, f1 = gpu_lib.tester_pattern(codes.builtin_codes_dir(), "ldpc/b200-minsum/jpl.4096.4.5/20/4/5", 1, rx[:64])
    assert np.array_equal(m1, msg[:64]) and l1 == f1 == 64
    # a general (alist) code through the same harness: moon.7.13 sum-product, ONCHIP kernel behind the mailbox
    moon = builtin_codes["moon.7.13"]
    rxm = awgn_llrs(moon, 3.0, 512, 1357)
    mm, sm, lm, fm = gpu_lib.tester_pattern(codes.builtin_codes_dir(), "ldpc/b200-spa/moon.7.13/20", 8, rxm)
    rb, ri, rc = oracles["moon.7.13"].decode_batch("spa", 20, rxm)
    assert (mm == rb[:, :moon.K]).all(axis=1).mean() >= 0.999 and fm == 512


def test_two_devices_from_one_thread(gpu_lib, builtin_codes, oracles):
    """ADVICE (round 1): kernel function attributes (dynamic shared-memory limit) are per DEVICE; the "configured" cache of the
    launchers is keyed by the device of the code.  One host thread decodes jpl.4096.4.5 (67 KB of dynamic shared memory, above
    the 48 KB default) on device 0 and then on device 1, batched path, mailbox and simulate.  Needs two visible GPUs."""
    L = gpu_lib.load()
    if L.ldpc_b200_device_count() < 2:
        pytest.skip("one visible GPU")
    code = builtin_codes["jpl.4096.4.5"]
    llr = awgn_llrs(code, 3.6, 24, 909, 5120)
    rb, ri, rc = oracles["jpl.4096.4.5"].decode_batch("min", 20, llr, variant=1)
    counters = []
    try:
        for dev in (0, 1, 0):
            gpu_lib.init(dev)
            dc = gpu_lib.DeviceCode(code)
            for algo in ("min", "spa"):
                dec = gpu_lib.Decoder(dc, algo, "fp32", max_batch=32)
                b, it, cv = dec.decode(llr, 20)                               # batched path
                b1, i1, c1 = dec.decode(llr[:1], 20)                          # mailbox + latency kernel
                if algo == "min":
                    assert np.array_equal(b, rb) and np.array_equal(it, ri) and np.array_equal(b1[0], rb[0])
                    counters.append(dec.simulate(20, 3.4, code.K, 5120, 3, 0, 2000))
                else:
                    assert np.array_equal(b1[0], b[0])
                dec.close()
            dc.close()
    finally:
        gpu_lib.init(0)
    for c in counters[1:]:
        assert all(c[k] == counters[0][k] for k in ("frames", "frame_errors", "bit_errors", "sum_iters", "unconverged"))


def _random_code(rng, kind):
    """A random parity-check matrix in the library's two input forms: CSR (irregular, row degree 2..dmax) or single-weight
    quasi-cyclic blocks with random shifts.  Every column gets at least one edge."""
    from ecc_ldpc_b200 import codes
    if kind == "csr":
        M, N, dmax = int(rng.integers(6, 60)), int(rng.integers(24, 160)), int(rng.integers(3, 12))
        rows = [sorted(rng.choice(N, size=int(rng.integers(2, min(dmax, N) + 1)), replace=False).tolist()) for _ in range(M)]
        covered = set(c for r in rows for c in r)
        for j in range(N):
            if j not in covered:
                r = rows[int(rng.integers(0, M))]
                if len(r) < 32:
                    r.append(j); r.sort()
        return codes.from_rows("rand-csr", M, N, rows, K=max(N - M, 1))
    sz = int(rng.choice([4, 8, 16, 32, 64]))
    br, bc = int(rng.integers(2, 6)), int(rng.integers(6, 14))
    shifts = np.where(rng.random((br, bc)) < 0.55, rng.integers(0, sz, (br, bc)), -1).astype(np.int32)
    for c in range(bc):
        if (shifts[:, c] < 0).all():
            shifts[int(rng.integers(0, br)), c] = int(rng.integers(0, sz))
    for r in range(br):
        while (shifts[r] >= 0).sum() < 2:
            shifts[r, int(rng.integers(0, bc))] = int(rng.integers(0, sz))
    code = codes.from_qc("rand-qc", sz, shifts)
    code.K = max(code.N - code.M, 1)
    return code


@pytest.mark.parametrize("seed", range(10))
def test_random_user_codes_all_engines(seed, gpu_lib):
    """User codes (`code_create_csr` / `code_create_qc` take ANY H): random irregular CSR matrices and random
    single-weight quasi-cyclic matrices that have no compiled specialisation.  float64 min-sum on the ONCHIP kernel (run-time
    tables, bank-conflict-free placement by edge colouring) and on STREAM: decisions, iteration counts and flags identical to
    the restatement on every frame; the message tap of both equals the restatement's messages bit for bit; float32
    sum-product decides like the restatement on >= 99 % of the frames; one-frame calls (mailbox) equal batched calls."""
    from oracle import oracle
    rng = np.random.default_rng(1000 + seed)
    code = _random_code(rng, "csr" if seed % 2 == 0 else "qc")
    orc = oracle.Oracle(code)
    dc = gpu_lib.DeviceCode(code)
    F = 96
    llr = 2.2 * (-1.0 + 0.85 * rng.standard_normal((F, code.N)))
    llr[:, ::7] = 0.0 if seed % 3 == 0 else llr[:, ::7]                     # some punctured positions
    rb, ri, rc = orc.decode_batch("min", 15, llr, variant=0)
    for eng in (gpu_lib.ENGINE_AUTO, gpu_lib.ENGINE_STREAM):
        dec = gpu_lib.Decoder(dc, "min", "fp64", max_batch=64, engine=eng)
        if eng == gpu_lib.ENGINE_AUTO:
            assert dec.engine() == gpu_lib.ENGINE_ONCHIP                    # small codes fit shared memory
        b, it, cv = dec.decode(llr, 15)
        assert np.array_equal(b, rb) and np.array_equal(it, ri) and np.array_equal(cv, rc), (seed, eng)
        f = int(np.argmax(ri))                                              # the frame that iterates longest
        tb, tit, tcv, tc2v, tlam = dec.decode_trace(llr[f], 15)
        ob, oit, ocv, oc2v, olam = orc.decode("min", 15, llr[f], variant=0, trace=True)
        assert tit == oit and np.array_equal(tc2v, oc2v) and np.array_equal(tlam, olam), (seed, eng)
        b1, i1, c1 = dec.decode(llr[:3], 15)                                # three frames: mailbox (ONCHIP) / batched (STREAM)
        assert np.array_equal(b1, rb[:3]) and np.array_equal(i1, ri[:3])
        dec.close()
    for algo in ("spa", "min"):
        sb, si, sc = orc.decode_batch(algo, 15, llr, variant=0)
        d32 = gpu_lib.Decoder(dc, algo, "fp32", max_batch=64)
        b, it, cv = d32.decode(llr, 15)
        assert (b == sb).all(axis=1).mean() >= 0.99, (seed, algo, (b == sb).all(axis=1).mean())
        s32 = gpu_lib.Decoder(dc, algo, "fp32", max_batch=64, engine=gpu_lib.ENGINE_STREAM)
        b2, it2, cv2 = s32.decode(llr, 15)
        assert np.array_equal(b, b2) and np.array_equal(it, it2)            # ONCHIP == STREAM in float32 (same arithmetic)
        d32.close(); s32.close()
    dc.close()


@pytest.mark.parametrize("code_name,n_tx,ebn0", [("jpl.1024.4.5", 1280, 4.6), ("jpl.4096.4.5", 5120, 4.2), ("1920.1280.3.303", None, 3.2)])
@pytest.mark.parametrize("algo", ["spa", "min"])
def test_erased_positions_production_engines(code_name, n_tx, ebn0, algo, gpu_lib, dcodes, builtin_codes, oracles):
    """Erasures: 6 % of the transmitted positions carry LLR exactly 0 (on top of the punctured tail).  An exact zero has
    tanh = 0 / min' = 0, so every other message of its checks is exactly 0 in the reference, and a variable that hears nothing
    else keeps lam = 0 and decides 0.  The float32 production kernels (RESIDENT / ONCHIP; complement domain exact at 1, min'
    exact) reproduce that: decisions on >= 99.9 % of the frames, iteration counts on >= 99 %."""
    code = builtin_codes[code_name]
    F = 1024 if algo == "min" else 384
    llr = awgn_llrs(code, ebn0, F, 60606, n_tx)
    rng = np.random.default_rng(3)
    llr[rng.random(llr.shape) < 0.06] = 0.0
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp32", max_batch=1024)
    assert dec.engine() in (gpu_lib.ENGINE_RESIDENT, gpu_lib.ENGINE_ONCHIP)
    b, it, cv = dec.decode(llr, 30)
    rb, ri, rc = oracles[code_name].decode_batch(algo, 30, llr, variant=0)
    assert 0.05 < rc.mean() <= 1.0
    assert (b == rb).all(axis=1).mean() >= 0.999, (code_name, algo, (b == rb).all(axis=1).mean())
    assert (it == ri).mean() >= 0.99
    dec.close()


@pytest.mark.parametrize("algo", ["min", "spa"])
def test_codeword_symmetry_at_config2_batch(algo, gpu_lib, builtin_codes):
    """Size-independent property at BASELINE config 2's size (jpl.1024.4.5, 16 384 frames in one call): belief propagation on
    a symmetric channel commutes with adding a codeword -- flipping the LLR signs at the ones of a codeword c flips exactly
    those decisions and changes nothing else (same iteration count, same convergence flag).  Random messages are encoded on
    the device (ldpc_b200_ecc_encode_device: Fast/Encoder.hs:41-63 word multiply), the frames of the all-zero codeword come
    from the device generator, and the float32 production engine decodes both sets."""
    from ecc_ldpc_b200 import ecc
    code = builtin_codes["jpl.1024.4.5"]
    F, K, N, n_tx = 16384, code.K, code.N, 1280
    name = "ldpc/b200-%s/jpl.1024.4.5/20/4/5" % ("minsum" if algo == "min" else "spa")
    (e,) = ecc.match(name, max_batch=F)
    e.run(3.0, 1)                                        # creates the device side of the record
    rng = np.random.default_rng(77)
    msgs = rng.integers(0, 2, (F, K)).astype(np.uint8)
    packed = np.packbits(msgs, axis=1, bitorder="little").view(np.uint32)
    Wn = (N + 31) // 32
    d_msg = gpu_lib.DeviceBuffer(packed.nbytes).upload(packed)
    d_cw = gpu_lib.DeviceBuffer(F * Wn * 4)
    e.encode_device(d_msg.ptr, d_cw.ptr, F)
    e.sync()
    cw = gpu_lib.unpack_bits(d_cw.download(np.uint32, (F, Wn)), N)[:, :n_tx]
    assert np.array_equal(cw[:, :K], msgs)
    dc = gpu_lib.DeviceCode(code)
    dec = gpu_lib.Decoder(dc, algo, "fp32", max_batch=F)
    assert dec.engine() == gpu_lib.ENGINE_RESIDENT
    buf = gpu_lib.DeviceBuffer(F * N * 4)
    dec.generate_llr(buf.ptr, 3.55 if algo == "min" else 3.1, K, n_tx, 5, 0, F)
    dec.sync()
    llr0 = buf.download(np.float32, (F, N))
    llrc = llr0.copy()
    llrc[:, :n_tx] *= (1.0 - 2.0 * cw).astype(np.float32)
    p0, i0, c0 = dec.decode_f32(llr0, 20)
    pc, ic, cc = dec.decode_f32(llrc, 20)
    b0 = gpu_lib.unpack_bits(p0, N)[:, :n_tx]
    bc = gpu_lib.unpack_bits(pc, N)[:, :n_tx]
    assert 0.02 < 1 - c0.mean() < 0.9                    # waterfall: both outcomes present
    assert np.array_equal(i0, ic) and np.array_equal(c0, cc)
    assert np.array_equal(bc, b0 ^ cw)
    dec.close(); dc.close(); e.close()


def test_code_outlives_its_handle_while_decoders_use_it(gpu_lib, builtin_codes, oracles):
    """Handle lifetimes: identical codes share one device object; destroying a code handle while a decoder still decodes with it
    leaves the decoder fully usable (the object goes with its last decoder), and destroying handles twice is harmless."""
    import ctypes as C
    L = gpu_lib.load()
    nc, nd = C.c_int(), C.c_int()
    L.ldpc_b200_live_handles(C.byref(nc), C.byref(nd))
    base = (nc.value, nd.value)
    code = builtin_codes["1920.1280.3.303"]
    a = gpu_lib.DeviceCode(code)
    b = gpu_lib.DeviceCode(code)
    assert a.h == b.h                                                   # one shared object, two handles
    dec = gpu_lib.Decoder(a, "min", "fp32", max_batch=16)
    a.close(); b.close()                                                # the decoder keeps the object alive
    L.ldpc_b200_live_handles(C.byref(nc), C.byref(nd))
    assert (nc.value, nd.value) == (base[0], base[1] + 1)
    llr = awgn_llrs(code, 2.6, 12, 11)
    rb, ri, rc = oracles["1920.1280.3.303"].decode_batch("min", 20, llr, variant=1)
    gb, gi, gc = dec.decode(llr, 20)
    g1, i1, c1 = dec.decode(llr[:1], 20)                                # mailbox path too
    assert (gb == rb).all(axis=1).mean() >= 0.9 and np.array_equal(g1[0], gb[0])
    dec.close(); dec.close()
    L.ldpc_b200_live_handles(C.byref(nc), C.byref(nd))
    assert (nc.value, nd.value) == base


@pytest.mark.gpu
@pytest.mark.parametrize("code_name,F,algo", [("jpl.1024.4.5", 32768, "min"), ("jpl.4096.4.5", 8192, "min"), ("jpl.1024.4.5", 16384, "spa"),
                                             ("1920.1280.3.303", 16384, "min")])
def test_fallback_decisions_exact_at_scale(code_name, F, algo, gpu_lib, dcodes, builtin_codes):
    """Far below the waterfall no frame converges, so every frame must come back as the hard decision of its CHANNEL values
    (`n >= maxIterations -> orig_lam`, Orig.hs:68-71 / Min.hs:64-67): an exact known answer for tens of thousands of frames
    per launch, independent of any restatement.  It is the case in which a frame's state is read at the very end of its life
    -- while the persistent CTA is already taking in the next frame -- so it pins the hand-over between frames (a missing
    barrier there showed up as 2 wrong frames in 16 384, inside the 99.9 % bar of the statistical tests)."""
    code = builtin_codes[code_name]
    W = (code.N + 31) // 32
    n_tx = TX.get(code_name) or code.N
    dec = gpu_lib.Decoder(dcodes[code_name], algo, "fp32", max_batch=F)
    d_llr = gpu_lib.DeviceBuffer(F * code.N * 4)
    d_bits = gpu_lib.DeviceBuffer(F * W * 4)
    d_it = gpu_lib.DeviceBuffer(F * 4)
    d_cv = gpu_lib.DeviceBuffer(F)
    dec.generate_llr(d_llr.ptr, -1.0, code.K, n_tx, 77, 0, F)
    dec.sync()
    llr = d_llr.download(np.float32, (F, code.N))
    for rep in range(3):                                   # the hand-over is a race if it is one: give it a few chances
        d_bits.zero()
        dec.decode_device(d_llr.ptr, d_bits.ptr, 20, F, early_exit=True, d_iters=d_it.ptr, d_conv=d_cv.ptr)
        dec.sync()
        bits = gpu_lib.unpack_bits(d_bits.download(np.uint32, (F, W)), code.N)
        cv = d_cv.download(np.uint8, (F,))
        it = d_it.download(np.int32, (F,))
        stuck = cv == 0
        assert stuck.mean() > 0.99 and (it[stuck] == 20).all()
        assert np.array_equal(bits[stuck], (llr[stuck] > 0).astype(bits.dtype)), "a non-converged frame is not hard(orig_lam)"
    dec.close()
