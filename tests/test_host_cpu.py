"""CPU suite: host-side pieces added in round 2 -- the host statement of the device frame generator and the CPU arm of
bench.py.  Nothing here touches the GPU."""
import json
import os
import subprocess
import sys

import numpy as np

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
