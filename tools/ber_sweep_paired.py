#!/usr/bin/env python3
"""BER/FER sweeps of the BASELINE.json configs, round 2: PAIRED comparison with the CPU restatement.

    python tools/ber_sweep_paired.py [cpu seconds per point] > profiles/r02_ber_fer.md

Per Eb/N0 point: (1) the curve -- ldpc_b200_simulate over a large range of Philox frames (all-zero codeword, BPSK + AWGN,
errors on the first K bits), with 95 % Wilson intervals; (2) the pairing -- the FIRST C frames of that same range are
generated again on the device, downloaded, and decoded by the float32 production engine and by the float64 CPU
restatement (oracle/, all host cores): frame errors and bit errors of both on identical frames, the number of frames
whose decisions differ, and whether the restatement's interval on C frames contains the GPU's large-sample estimate
(north_star: "BER/FER curves within 95 % binomial confidence of the reference across the Eb/N0 sweep")."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from ecc_ldpc_b200 import codes, lib  # noqa: E402
from oracle import oracle  # noqa: E402


def wilson(k, n, z=1.959963984540054):
    if n == 0:
        return 0.0, 1.0
    p = k / n
    den = 1 + z * z / n
    c = (p + z * z / (2 * n)) / den
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / den
    return max(0.0, c - h), min(1.0, c + h)


budget = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
CONFIGS = [
    ("moon.7.13", "spa", 20, None, [0, 1, 2, 3, 4, 5, 6], 4000000),
    ("jpl.1024.4.5", "min", 20, 1280, [2.0, 2.5, 3.0, 3.5, 4.0, 4.5], 2000000),
    ("jpl.4096.4.5", "min", 50, 5120, [2.5, 2.75, 3.0, 3.25, 3.5, 4.0], 1000000),
    ("jpl.4096.4.5", "spa", 50, 5120, [2.25, 2.5, 2.75, 3.0, 3.5], 400000),
    ("1920.1280.3.303", "min", 50, None, [1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 5.0], 2000000),
    ("1920.1280.3.303", "spa", 50, None, [1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 5.0], 1000000),
]
lib.init(0)
cores = oracle.lib().oracle_max_threads()
print("# BER / FER sweeps, float32 production engines vs float64 CPU restatement on identical frames (round 2)\n")
print("All-zero codeword, BPSK + AWGN, sigma^2 = 1/(2 R Eb/N0), R = K/n_tx, errors on the first K bits; Philox frames [0, F) per point;")
print("CPU restatement on %d host threads, ~%.0f s per point: it decodes the first C frames of the GPU's range. `in CI` = the GPU's" % (cores, budget))
print("large-sample FER lies inside the restatement's 95 %% Wilson interval on its C frames.\n")
t_all = time.time()
for name, algo, iters, n_tx, grid, gframes in CONFIGS:
    code = codes.load_builtin(name)
    tx = n_tx or code.N
    dc = lib.DeviceCode(code)
    dec = lib.Decoder(dc, algo, "fp32", max_batch=16384)
    orc = oracle.Oracle(code)
    var = 1 if algo == "min" else 0
    buf = lib.DeviceBuffer(8192 * code.N * 4)
    dec.generate_llr(buf.ptr, grid[len(grid) // 2], code.K, tx, 7, 0, 128); dec.sync()
    probe = buf.download(np.float32, (128, code.N)).astype(np.float64)
    t0 = time.perf_counter(); orc.decode_batch(algo, iters, probe, variant=var); rate = 128 / (time.perf_counter() - t0)
    C = int(min(65536, max(512, rate * budget)) // 256 * 256)
    print("## %s, %s, max %d iterations, R = %d/%d, engine %s; paired sample C = %d frames per point\n" % (
        name, {"min": "min-sum (alpha=3/4)", "spa": "sum-product"}[algo], iters, code.K, tx, {1: "STREAM", 2: "RESIDENT", 3: "ONCHIP"}[dec.engine()], C))
    print("| Eb/N0 dB | GPU frames | GPU FER [95 %] | GPU BER | avg it | paired: GPU FE / CPU FE | GPU BE / CPU BE | frames differing | iteration counts equal | CPU FER [95 %] on C | in CI |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for eb in grid:
        c = dec.simulate(iters, eb, code.K, tx, 7, 0, gframes)
        glo, ghi = wilson(c["frame_errors"], c["frames"])
        gfe = gbe = cfe = cbe = diff = eqit = 0
        for f0 in range(0, C, 8192):
            n = min(8192, C - f0)
            dec.generate_llr(buf.ptr, eb, code.K, tx, 7, f0, n); dec.sync()
            llr = buf.download(np.float32, (n, code.N)).astype(np.float64)
            gb, gi, gc = dec.decode(llr, iters)
            rb, ri, rc = orc.decode_batch(algo, iters, llr, variant=var)
            gfe += int(gb[:, :code.K].any(axis=1).sum()); gbe += int(gb[:, :code.K].sum())
            cfe += int(rb[:, :code.K].any(axis=1).sum()); cbe += int(rb[:, :code.K].sum())
            diff += int((gb != rb).any(axis=1).sum()); eqit += int((gi == ri).sum())
        clo, chi = wilson(cfe, C)
        gf = c["frame_errors"] / c["frames"]
        print("| %.2f | %d | %.3e [%.2e, %.2e] | %.3e | %.2f | %d / %d | %d / %d | %d | %.5f | %.3e [%.2e, %.2e] | %s |" % (
            eb, c["frames"], gf, glo, ghi, c["bit_errors"] / (c["frames"] * code.K), c["sum_iters"] / c["frames"],
            gfe, cfe, gbe, cbe, diff, eqit / C, cfe / C, clo, chi, "yes" if (clo - 1e-12 <= gf <= chi + 1e-12) else "NO"))
        sys.stdout.flush()
    print()
    buf.free(); dec.close(); dc.close()
print("(total run time %.0f s)" % (time.time() - t_all))
