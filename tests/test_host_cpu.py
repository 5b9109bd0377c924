"""CPU suite: host-side pieces added in round 2 -- the host statement of the device frame generator and the CPU arm of
bench.py.  Nothing here touches the GPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_philox_host_matches_exported_block_function():
    """ecc-ldpc_b200/philox_host.py (numpy) == ldpc_b200_philox4x32 (the inline function the kernels use, host instance)
    on random counters and keys, and on the Random123 known answers."""
    from ecc_ldpc_b200 import lib, philox_host
    rng = np.random.default_rng(5)
    ctr = rng.integers(0, 2 ** 32, (64, 4), dtype=np.uint64)
    key = rng.integers(0, 2 ** 32, 2, dtype=np.uint64)
    got = np.stack(philox_host.philox4x32_10(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], int(key[0]), int(key[1])), axis=1)
    for i in range(64):
        assert got[i].tolist() == lib.philox4x32(ctr[i].astype(np.uint32), key.astype(np.uint32)).tolist()
    kat = philox_host.philox4x32_10(*[np.array([0xffffffff], dtype=np.uint64)] * 4, 0xffffffff, 0xffffffff)
    assert [int(v[0]) for v in kat] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_philox_host_frames_are_awgn_llrs():
    """Statistics of the host generator (the device one is checked against it on the GPU): y = -1 + sigma n, llr = 2y/sigma^2,
    punctured tail zero, streams keyed by the global frame index (a sub-range regenerates the same frames)."""
    from ecc_ldpc_b200 import philox_host
    N, K, n_tx, ebn0 = 5632, 4096, 5120, 3.0
    llr = philox_host.awgn_llrs(N, ebn0, K, n_tx, 99, 1000, 96).astype(np.float64)
    assert not llr[:, n_tx:].any()
    s2 = 1.0 / (2 * (K / n_tx) * 10 ** (ebn0 / 10))
    y = llr[:, :n_tx] * s2 / 2
    assert abs(y.mean() + 1.0) < 5 * np.sqrt(s2 / y.size) and abs(y.var() / s2 - 1.0) < 0.02
    again = philox_host.awgn_llrs(N, ebn0, K, n_tx, 99, 1040, 8)
    assert np.array_equal(again, llr[40:48].astype(np.float32))
    other = philox_host.awgn_llrs(N, ebn0, K, n_tx, 100, 1000, 8)
    assert not np.array_equal(other, llr[:8].astype(np.float32))


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (CPU arm): one JSON line with the contract's keys, on the GPU arm's config object and
    on Philox frames (no frame converges at 2.0 dB: all 20 iterations run)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-frames", "32"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoded_info_gbit_per_s" and d["unit"] == "Gbit/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["avg_iters"] == 20.0 and d["gpu_launches"] == 0
    assert set(d["config"]) == {"workload", "frames_per_step_per_gpu", "frames", "l2", "sharding"} and d["config"]["frames_per_step_per_gpu"] == 32768
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_cpp_matrix_readers_on_random_matrices():
    """The C++ readers of the reference's three matrix formats (ldpc_b200_matrix_parse: QuasiCyclic.hs:19-25,52-56, Alist.hs:30-46,
    BitMatrix/Matlab.hs:19-26) against the Python statements of the same formats on random matrices: dense 0/1 text, alist in the
    reference's orientation (zeros dropped, short rows padded with zeros), and .q with circulant sizes up to 200 (entries beyond
    64 and 128 bits, several ones per circulant)."""
    from ecc_ldpc_b200 import codes, lib
    rng = np.random.default_rng(2026)
    for trial in range(24):
        # ---- .m and .alist from the same random matrix ----
        M, N = int(rng.integers(1, 30)), int(rng.integers(2, 50))
        dense = (rng.random((M, N)) < 0.2).astype(np.uint8)
        if trial % 4:                      # most trials: no empty row or column, so that the alist dialect can express the matrix
            for m in np.nonzero(dense.sum(axis=1) == 0)[0]:
                dense[m, rng.integers(0, N)] = 1
            for j in np.nonzero(dense.sum(axis=0) == 0)[0]:
                dense[rng.integers(0, M), j] = 1
        rows = [np.nonzero(dense[m])[0].tolist() for m in range(M)]
        want_rp = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
        want_ci = np.array([c for r in rows for c in r], dtype=np.int32)
        m_text = "\n".join(" ".join(str(int(v)) for v in dense[m]) for m in range(M)) + "\n"
        a_text = codes._alist_text(M, N, rows)
        for kind, text in (("m", m_text), ("alist", a_text)):
            if kind == "alist" and (dense.sum(axis=1).min() == 0 or dense.sum(axis=0).min() == 0):
                continue            # the reference's reader drops EVERY zero token (Alist.hs:31), weights included: a zero-weight
                                    # row or column cannot be written in its alist dialect
            h = lib.HostMatrix.parse(kind, text)
            assert (h.rows, h.cols, h.ones) == (M, N, int(dense.sum())), (trial, kind)
            rp, ci = h.csr()
            assert np.array_equal(rp, want_rp) and np.array_equal(ci, want_ci), (trial, kind)
            assert codes.parse_m(text)[2] == rows if kind == "m" else codes.parse_alist(text)[2] == rows
            h.close()
        # ---- .q: random circulant words, possibly several bits per word ----
        sz = int(rng.choice([3, 8, 32, 64, 65, 128, 129, 200]))
        br, bc = int(rng.integers(1, 5)), int(rng.integers(1, 7))
        words = [[0 if rng.random() < 0.4 else sum(1 << int(k) for k in rng.choice(sz, size=int(rng.integers(1, 4)), replace=False)) for _ in range(bc)] for _ in range(br)]
        q = lib.HostMatrix.parse("q", codes._q_text(sz, words))
        qrows = codes.expand_qc_rows(sz, words)
        assert (q.rows, q.cols, q.sz) == (br * sz, bc * sz, sz) and q.ones == sum(len(r) for r in qrows)
        rp, ci = q.csr()
        assert ci.tolist() == [c for r in qrows for c in r] and rp[-1] == q.ones
        single = all(bin(w).count("1") <= 1 for row in words for w in row)
        if single:
            want = [[(w.bit_length() - 1) if w else -1 for w in row] for row in words]
            assert q.qc_shifts().tolist() == want
        else:
            with pytest.raises(lib.LdpcError):
                q.qc_shifts()
        q.close()
