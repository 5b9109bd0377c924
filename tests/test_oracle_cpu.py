"""CPU suite: the C oracle against the dense second statement (golden fixtures), code invariants,
host logic.  Nothing here touches the GPU."""
import glob
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, awgn_llrs, golden

ALGOS = {"spa": "spa", "min": "min"}


def _golden_files():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def _parse(fname):
    base = fname[:-4]
    code, algo, it = base.rsplit("_", 2)
    return code, algo, int(it[2:])


def test_golden_present():
    names = _golden_files()
    assert len(names) >= 8, names


@pytest.mark.parametrize("fname", _golden_files())
def test_oracle_matches_dense_statement(fname, oracles):
    """Bit-exact: decisions, iteration counts and per-edge messages of the C oracle == dense_ref fixtures."""
    code_name, algo, max_iters = _parse(fname)
    g = golden(fname)
    o = oracles[code_name]
    for f in range(g["llr"].shape[0]):
        bits, it, cv, c2v, lam = o.decode(algo, max_iters, g["llr"][f], trace=True)
        assert it == g["iters"][f] and cv == bool(g["converged"][f])
        assert np.array_equal(bits, g["bits"][f])
        if it > 0:
            assert np.array_equal(c2v[0], g["c2v_first"][f])          # == treats -0.0 and 0.0 alike
            assert np.array_equal(c2v[-1], g["c2v_last"][f])
            assert np.array_equal(lam[0], g["lam_first"][f])
            assert np.array_equal(lam[-1], g["lam_last"][f])
        if "c2v_all" in g.files:
            assert np.array_equal(c2v, g["c2v_all"][f][:it])
            assert np.array_equal(lam, g["lam_all"][f][:it])


@pytest.mark.parametrize("code_name,ebn0,n_tx", [("moon.7.13", 3.0, None), ("jpl.1024.4.5", 3.8, 1280),
                                                 ("1920.1280.3.303", 2.5, None)])
def test_twomin_equals_literal_minsum(code_name, ebn0, n_tx, builtin_codes, oracles):
    """Model.hs:326-343 two-minimum form == Min.hs:79-87 literal foldr1 (messages identical)."""
    code, o = builtin_codes[code_name], oracles[code_name]
    llr = awgn_llrs(code, ebn0, 6, 5, n_tx)
    for f in range(llr.shape[0]):
        a = o.decode("min", 15, llr[f], variant=0, trace=True)
        b = o.decode("min", 15, llr[f], variant=1, trace=True)
        assert a[1] == b[1] and a[2] == b[2] and np.array_equal(a[0], b[0])
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])


def test_fixture_invariants(builtin_codes):
    """dimensions, edge counts, degree profiles, [I|G] H^T = 0."""
    c = builtin_codes
    assert (c["moon.7.13"].M, c["moon.7.13"].N, c["moon.7.13"].E, c["moon.7.13"].K) == (13, 20, 60, 7)
    assert (c["jpl.1024.4.5"].M, c["jpl.1024.4.5"].N, c["jpl.1024.4.5"].E, c["jpl.1024.4.5"].K) == (384, 1408, 4992, 1024)
    assert (c["jpl.4096.4.5"].M, c["jpl.4096.4.5"].N, c["jpl.4096.4.5"].E, c["jpl.4096.4.5"].K) == (1536, 5632, 19968, 4096)
    assert (c["1920.1280.3.303"].M, c["1920.1280.3.303"].N, c["1920.1280.3.303"].E, c["1920.1280.3.303"].K) == (1280, 1920, 5120, 640)
    rowdeg = np.diff(c["moon.7.13"].row_ptr).tolist()
    assert rowdeg == [4, 6, 4, 4, 6, 8, 4, 4, 4, 4, 2, 4, 6]
    for name, sz in (("jpl.1024.4.5", 32), ("jpl.4096.4.5", 128)):
        code = c[name]
        rd = np.diff(code.row_ptr)
        assert sorted(set(rd.tolist())) == [3, 18] and (rd == 3).sum() == 4 * sz and (rd == 18).sum() == 8 * sz
        cd = np.bincount(code.col_idx, minlength=code.N)
        assert {int(k): int((cd == k).sum()) for k in np.unique(cd)} == {1: 4 * sz, 2: 4 * sz, 3: 8 * sz, 4: 24 * sz, 6: 4 * sz}
        assert (code.shifts >= 0).sum() == 156
    rd = np.diff(c["1920.1280.3.303"].row_ptr)
    assert set(rd.tolist()) == {4}
    for name in ("moon.7.13", "jpl.1024.4.5", "jpl.4096.4.5"):
        code = c[name]
        H = code.dense().astype(np.int64)
        G = np.concatenate([np.eye(code.K, dtype=np.int64), code.g_dense().astype(np.int64)], axis=1)
        assert not ((G @ H.T) % 2).any()


def test_fixture_provenance(builtin_codes):
    """When the reference checkout is present, the JSON fixtures must hash back to its files."""
    ref = "/root/reference/codes"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present (GPU box)")
    for name, code in builtin_codes.items():
        for fn, digest in code.sha256.items():
            p = os.path.join(ref, fn) if name == "1920.1280.3.303" else os.path.join(ref, name, fn)
            assert hashlib.sha256(open(p, "rb").read()).hexdigest() == digest


def test_reference_format_readers(builtin_codes):
    """Loader.hs:58-81 search order and formats: the directory loader reproduces the JSON fixtures."""
    from ecc_ldpc_b200 import codes
    ref = "/root/reference/codes"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present (GPU box)")
    for name in ("moon.7.13", "jpl.1024.4.5", "jpl.4096.4.5"):
        a, b = codes.load_dir(name, ref), builtin_codes[name]
        assert (a.M, a.N, a.K) == (b.M, b.N, b.K)
        assert np.array_equal(a.row_ptr, b.row_ptr) and np.array_equal(a.col_idx, b.col_idx)
        assert np.array_equal(a.g_row_ptr, b.g_row_ptr) and np.array_equal(a.g_col_idx, b.g_col_idx)
    # H.alist of jpl.1024.4.5 is the expansion of H.q (Appendix B)
    n, m, rows = codes.parse_alist(open(os.path.join(ref, "jpl.1024.4.5", "H.alist")).read())
    b = builtin_codes["jpl.1024.4.5"]
    assert (n, m) == (b.M, b.N)
    assert [c for r in rows for c in r] == b.col_idx.tolist()


def test_valid_codeword_is_fixed_point(builtin_codes, oracles):
    """LLR = +-c on a codeword: syndrome passes at iteration 0, output = codeword (Orig.hs:68-69)."""
    from oracle import oracle
    rng = np.random.default_rng(3)
    for name in ("moon.7.13", "jpl.1024.4.5"):
        code = builtin_codes[name]
        msg = rng.integers(0, 2, code.K).astype(np.uint8)
        cw = oracle.encode(code, msg)
        assert not ((code.dense().astype(np.int64) @ cw) % 2).any()
        for algo in ("spa", "min"):
            bits, it, cv = oracles[name].decode(algo, 20, 4.0 * (2.0 * cw - 1.0))
            assert it == 0 and cv and np.array_equal(bits, cw)


def test_fallback_returns_channel_hard_decision(builtin_codes, oracles):
    """Orig.hs:70: when max iterations are exhausted the result is hard(orig_lam), not hard(lam)."""
    code, o = builtin_codes["jpl.1024.4.5"], oracles["jpl.1024.4.5"]
    llr = awgn_llrs(code, 0.0, 4, 9, 1280)               # hopeless SNR
    for f in range(4):
        bits, it, cv = o.decode("min", 5, llr[f])
        assert it == 5 and not cv
        assert np.array_equal(bits, (llr[f] > 0).astype(np.uint8))
    bits, it, cv = o.decode("min", 0, llr[0])            # max_iters = 0: only the iteration-0 syndrome test
    assert it == 0 and not cv and np.array_equal(bits, (llr[0] > 0).astype(np.uint8))


def test_hard_decision_ties_and_punctured_zeros(builtin_codes, oracles):
    """hard = (> 0): an LLR of exactly 0 (punctured position, Utils.hs:55) decides 0."""
    code, o = builtin_codes["moon.7.13"], oracles["moon.7.13"]
    bits, it, cv = o.decode("spa", 10, np.zeros(code.N))
    assert it == 0 and cv and not bits.any()


def test_ecc_wrapper_unpuncture_and_truncate(builtin_codes, oracles):
    """Utils.hs:55,68-72: pad the received vector with zeros to N, decode, keep the first K bits."""
    from oracle import oracle
    code, o = builtin_codes["jpl.1024.4.5"], oracles["jpl.1024.4.5"]
    c_len = code.c_length(4, 5)
    assert c_len == 1280 and code.c_length() == 1408
    rng = np.random.default_rng(4)
    msg = rng.integers(0, 2, code.K).astype(np.uint8)
    cw = oracle.encode(code, msg, c_len)
    assert len(cw) == c_len and np.array_equal(cw[:code.K], msg)
    llr = awgn_llrs(code, 5.0, 1, 8, c_len, codewords=np.concatenate([cw, np.zeros(code.N - c_len, np.uint8)])[None, :])[0]
    out, flag, it, cv = o.ecc_decode("min", 20, code.K, c_len, llr[:c_len])
    assert flag and cv and np.array_equal(out, msg)
    full = np.concatenate([llr[:c_len], np.zeros(code.N - c_len)])
    bits, it2, cv2 = o.decode("min", 20, full)
    assert it2 == it and np.array_equal(bits[:code.K], out)


def test_batch_equals_single(builtin_codes, oracles):
    code, o = builtin_codes["jpl.1024.4.5"], oracles["jpl.1024.4.5"]
    llr = awgn_llrs(code, 3.6, 12, 17, 1280)
    bits, iters, conv = o.decode_batch("min", 20, llr, variant=1, threads=4)
    for f in range(12):
        b, it, cv = o.decode("min", 20, llr[f])
        assert np.array_equal(b, bits[f]) and it == iters[f] and cv == bool(conv[f])


def test_uncoded_ber_identity(builtin_codes):
    """Zero.hs:25-26 / NOTES.txt:3: hard decision on the channel gives Q(sqrt(2 R Eb/N0))."""
    from math import erfc, sqrt
    code = builtin_codes["jpl.1024.4.5"]
    llr = awgn_llrs(code, 0.0, 256, 123)                 # default rate 1024/1408 (Utils.hs:46-48)
    ber = (llr[:, :code.K] > 0).mean()
    q = 0.5 * erfc(sqrt(2 * (code.K / code.N)) / sqrt(2))
    assert abs(ber - q) < 4 * sqrt(q * (1 - q) / (256 * code.K))
    assert abs(q - 0.11390) < 2e-4


def test_code_name_grammar():
    """Utils.hs:100-108: ldpc/<decoder>/<matrix>/<max-rounds>[/<num>/<den>]; anything else matches nothing."""
    from ecc_ldpc_b200 import ecc
    (e,) = ecc.match("ldpc/b200-minsum/jpl.4096.4.5/20/4/5")
    assert (e.name, e.message_length, e.codeword_length, e.unpunctured_length, e.max_iters) == \
        ("ldpc/b200-minsum/jpl.4096.4.5/20/4/5", 4096, 5120, 5632, 20)
    (m,) = ecc.match("ldpc/b200-spa/moon.7.13/200")
    assert (m.name, m.message_length, m.codeword_length) == ("ldpc/b200-spa/moon.7.13/200/7/20", 7, 20)   # default rate K/N
    (j,) = ecc.match("ldpc/b200-minsum/jpl.1024.4.5/20")
    assert j.name == "ldpc/b200-minsum/jpl.1024.4.5/20/8/11" and j.codeword_length == 1408          # 1024/1408 reduced
    for bad in ("ldpc/reference/moon.7.13/20", "ldpc/b200-minsum/moon.7.13/x", "ldpc/b200-minsum/moon.7.13",
                "bpsk", "ldpc/b200-minsum/moon.7.13/20/4", "ldpc/b200-minsum/moon.7.13/20/a/5", ""):
        assert ecc.match(bad) == []
    from ecc_ldpc_b200 import lib
    with pytest.raises(lib.LdpcError) as ei:
        ecc.match("ldpc/b200-minsum/no.such.code/20")
    assert ei.value.code == lib.ENOENT and "can not find any matrix files" in str(ei.value)


def test_cpp_loader_reads_reference_formats(builtin_codes, tmp_path):
    """The C++ readers (Loader.hs / QuasiCyclic.hs / Alist.hs / Matlab.hs semantics) against the JSON fixtures, through
    files written in the reference's formats; plus the reference's own files when the checkout is present."""
    from ecc_ldpc_b200 import codes, lib
    root = codes.write_reference_dir(str(tmp_path))
    roots = [root] + (["/root/reference/codes"] if os.path.isdir("/root/reference/codes") else [])
    for r in roots:
        for name in ("moon.7.13", "jpl.1024.4.5", "jpl.4096.4.5"):
            b = builtin_codes[name]
            H = lib.HostMatrix.load(r, name + "/H")
            G = lib.HostMatrix.load(r, name + "/G")
            assert (H.rows, H.cols, H.ones) == (b.M, b.N, b.E) and (G.rows, G.cols) == (b.K, b.N - b.K)
            rp, ci = H.csr()
            assert np.array_equal(rp, b.row_ptr) and np.array_equal(ci, b.col_idx)
            grp, gci = G.csr()
            assert np.array_equal(grp, b.g_row_ptr) and np.array_equal(gci, b.g_col_idx)
            if b.is_qc:
                assert H.sz == b.sz and np.array_equal(H.qc_shifts(), b.shifts)
                with pytest.raises(lib.LdpcError):
                    G.qc_shifts()                     # dense circulants: "non-powers of two" (Arraylet.hs:72-79)
    # search order q > alist > m (Loader.hs:58-62), and the 0/1 matrix format
    d = tmp_path / "tiny"
    d.mkdir()
    (d / "H.m").write_text("1 1 0\n0 1 1\n")
    m = lib.HostMatrix.load(str(tmp_path), "tiny/H")
    assert (m.rows, m.cols, m.ones) == (2, 3, 4) and m.csr()[1].tolist() == [0, 1, 1, 2]
    (d / "H.alist").write_text("2 3\n2 2\n2 1\n1 1 1\n1 3\n2\n1\n2\n1\n")
    a = lib.HostMatrix.load(str(tmp_path), "tiny/H")
    assert a.ones == 3 and a.csr()[1].tolist() == [0, 2, 1]
    (d / "H.q").write_text("2\n1 2 0\n")
    q = lib.HostMatrix.load(str(tmp_path), "tiny/H")
    assert (q.rows, q.cols, q.sz) == (2, 6, 2) and q.csr()[1].tolist() == [0, 3, 1, 2] and q.qc_shifts().tolist() == [[0, 1, -1]]
    with pytest.raises(lib.LdpcError):
        lib.HostMatrix.parse("m", "1 2\n")
    with pytest.raises(lib.LdpcError):
        lib.HostMatrix.parse("q", "4\n1 2\n1\n")


def test_cpp_encoder_matches_oracle(builtin_codes):
    """ldpc_b200_ecc_encode (Fast/Encoder.hs:41-63 word multiply; Orig.hs:25-26 for alist G) == the oracle's dense v*G,
    and every codeword satisfies H."""
    from ecc_ldpc_b200 import ecc
    from oracle import oracle
    rng = np.random.default_rng(2)
    for nm, code_name in (("ldpc/b200-minsum/moon.7.13/20", "moon.7.13"), ("ldpc/b200-minsum/jpl.1024.4.5/20/4/5", "jpl.1024.4.5"),
                          ("ldpc/b200-minsum/jpl.4096.4.5/20", "jpl.4096.4.5")):
        (e,) = ecc.match(nm)
        code = builtin_codes[code_name]
        for t in range(3):
            msg = rng.integers(0, 2, code.K).astype(np.uint8)
            cw = e.encode(msg)
            assert np.array_equal(cw, oracle.encode(code, msg, e.codeword_length))
            if e.codeword_length == code.N:
                assert not ((code.dense().astype(np.int64) @ cw) % 2).any()
        e.close()


def test_ber_line_format():
    """NOTES.txt:3 `39.70 ldpc/model/jpl.1K/200 0.0 256 29891 1.14e-1 +1% -1% 95%`."""
    from ecc_ldpc_b200 import lib
    line = lib.ber_line(39.70, "ldpc/model/jpl.1K/200", 0.0, 256, 29891, 1024)
    assert line == "39.70 ldpc/model/jpl.1K/200 0.0 256 29891 1.14e-1 +1% -1% 95%"


def test_cached_mult_statedness(builtin_codes, oracles):
    """Fast/CachedMult.hs (`ldpc/arraylet-cm`): the StableDiv row product divided by the own term is the leave-one-out
    product of Orig.hs up to rounding, so decisions and iteration counts agree with the plain sum-product and the
    messages differ only in the last digits; a valid codeword is a fixed point at iteration 0; the fallback rule holds."""
    for name, ebn0, n_tx, iters in (("moon.7.13", 3.0, None, 20), ("jpl.1024.4.5", 3.9, 1280, 10)):
        code, o = builtin_codes[name], oracles[name]
        llr = awgn_llrs(code, ebn0, 12, 99, n_tx)
        agree = 0
        for f in range(llr.shape[0]):
            a = o.decode("spa", iters, llr[f], trace=True)
            b = o.decode("spa_cm", iters, llr[f], trace=True)
            agree += int(np.array_equal(a[0], b[0]) and a[1] == b[1])
            n = min(a[1], b[1], 2)
            if n:
                # first iterations: same quantity computed two ways (|msg| <= 12: away from the atanh' clamp)
                m = np.abs(a[3][:n]) <= 12
                assert np.allclose(a[3][:n][m], b[3][:n][m], rtol=1e-9, atol=1e-12)
        assert agree >= llr.shape[0] - 1
        cw = np.zeros(code.N)
        bits, it, cv = o.decode("spa_cm", 5, 3.0 * (2.0 * cw - 1.0))
        assert it == 0 and cv and not bits.any()
    code, o = builtin_codes["jpl.1024.4.5"], oracles["jpl.1024.4.5"]
    llr = awgn_llrs(code, 1.0, 2, 3, 1280)                     # hopeless SNR: never converges in 3 iterations
    for f in range(2):
        bits, it, cv = o.decode("spa_cm", 3, llr[f])
        assert it == 3 and not cv and np.array_equal(bits, (llr[f] > 0).astype(np.uint8))     # hard(orig_lam)


# ------------------------------------------------------------------------------------------------
# Third statement: the arraylet modules in circulant-vector form (oracle/arraylet_oracle.c, oracle/arraylet_ref.py)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fname", [f for f in _golden_files() if "_arraylet" in f])
def test_arraylet_c_statement_matches_python_circulant_fixtures(fname, builtin_codes):
    """oracle/arraylet_oracle.c (Matrixlet of per-circulant vectors, Fast/ArrayletMin.hs:23-114,167-192 and
    Fast/Arraylet.hs:164-186) reproduces the fixtures written by the pure-Python transcription of the same modules bit
    for bit: decisions, iteration counts, every message and posterior."""
    from oracle import oracle
    code_name, algo, max_iters = _parse(fname)
    g = golden(fname)
    a = oracle.ArrayletOracle(builtin_codes[code_name])
    for f in range(g["llr"].shape[0]):
        bits, it, cv, c2v, lam = a.decode("spa" if algo == "arraylet" else "min", max_iters, g["llr"][f], trace=True)
        assert it == g["iters"][f] and cv == bool(g["converged"][f]) and np.array_equal(bits, g["bits"][f])
        if it:
            assert np.array_equal(c2v[0], g["c2v_first"][f]) and np.array_equal(c2v[-1], g["c2v_last"][f])
            assert np.array_equal(lam[0], g["lam_first"][f]) and np.array_equal(lam[-1], g["lam_last"][f])
        if "c2v_all" in g.files:
            assert np.array_equal(c2v, g["c2v_all"][f][:it]) and np.array_equal(lam, g["lam_all"][f][:it])


@pytest.mark.parametrize("code_name,ebn0,n_tx,frames", [("jpl.1024.4.5", 3.7, 1280, 600), ("jpl.4096.4.5", 3.3, 5120, 160)])
@pytest.mark.parametrize("algo", ["min", "spa"])
def test_arraylet_statement_agrees_with_orig_min_order(code_name, ebn0, n_tx, frames, algo, builtin_codes, oracles):
    """The arraylet modules differ from Reference/Orig.hs / Min.hs only in the association of the posterior sum
    (L + (a1 + (a2 + ...)) vs a1 + (a2 + (... + L))): on waterfall frames (late convergence and failures both present)
    the circulant-vector statement, the edge-list statement of the same association and the Orig/Min-order oracle give
    the same decisions, iteration counts and flags; the two arraylet statements agree in every message bit."""
    from oracle import oracle
    code = builtin_codes[code_name]
    if algo == "spa":          # sum-product converges ~0.5 dB earlier than the alpha = 3/4 min-sum
        frames //= 4
        ebn0 -= 0.55
    llr = awgn_llrs(code, ebn0, frames, 8080, n_tx)
    a = oracle.ArrayletOracle(code)
    ab, ai, ac = a.decode_batch(algo, 20, llr)
    eb, ei, ec = oracles[code_name].decode_batch("arraylet-min" if algo == "min" else "arraylet", 20, llr)
    ob, oi, oc = oracles[code_name].decode_batch(algo, 20, llr, variant=0)
    assert np.array_equal(ab, eb) and np.array_equal(ai, ei) and np.array_equal(ac, ec)
    assert 0 < oc.mean() < 1 and (oi[oc == 1] >= 8).any()
    assert np.array_equal(ab, ob) and np.array_equal(ai, oi) and np.array_equal(ac, oc)
    for f in range(2):
        x = a.decode(algo, 20, llr[f], trace=True)
        y = oracles[code_name].decode("arraylet-min" if algo == "min" else "arraylet", 20, llr[f], trace=True)
        z = oracles[code_name].decode(algo, 20, llr[f], trace=True)
        assert np.array_equal(x[3], y[3]) and np.array_equal(x[4], y[4])
        n = min(x[1], z[1])
        if n:     # same messages up to the few ulp the association moves
            assert np.allclose(x[4][:n], z[4][:n], rtol=1e-12, atol=1e-12) and np.allclose(x[3][:1], z[3][:1], rtol=0, atol=0)
