#!/usr/bin/env python3
"""BER/FER sweep of the BASELINE.json configs on the GPU, with the CPU oracle beside it on a bounded sample.

    python tools/ber_sweep.py [--quick] > profiles/r01_ber_fer.md

GPU: ldpc_b200_simulate (device-generated BPSK+AWGN frames, all-zero codeword).  CPU: oracle/ldpc_oracle.c on numpy
noise (independent samples), so the two columns agree within binomial confidence, not bit for bit; the bit-for-bit
comparison on identical LLRs is tests/test_gpu_parity.py.  95 % Wilson intervals.
"""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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


def cpu_point(code, orc, algo, iters, ebn0, n_tx, frames, seed):
    rng = np.random.default_rng(seed)
    R = code.K / n_tx
    sigma = math.sqrt(1.0 / (2 * R * 10 ** (ebn0 / 10)))
    llr = 2.0 * (-1.0 + sigma * rng.standard_normal((frames, code.N))) / sigma ** 2
    llr[:, n_tx:] = 0.0
    bits, its, conv = orc.decode_batch(algo, iters, llr, variant=1 if algo == "min" else 0)
    return int(bits[:, :code.K].any(axis=1).sum()), int(bits[:, :code.K].sum()), float(its.mean())


def main():
    quick = "--quick" in sys.argv
    lib.init(0)
    configs = [
        ("moon.7.13", "spa", 20, None, [0, 1, 2, 3, 4, 5, 6], 400000, 20000),
        ("jpl.1024.4.5", "min", 20, 1280, [2.0, 3.0, 3.5, 4.0, 4.5], 400000, 1500),
        ("jpl.4096.4.5", "min", 50, 5120, [2.5, 3.0, 3.25, 3.5, 4.0], 200000, 300),
        ("jpl.4096.4.5", "spa", 50, 5120, [2.5, 2.75, 3.0, 3.5], 100000, 120),
        ("1920.1280.3.303", "min", 50, None, [1.0, 2.0, 2.5, 3.0, 4.0, 5.0], 200000, 600),
        ("1920.1280.3.303", "spa", 50, None, [1.0, 1.5, 2.0, 2.5, 3.0], 100000, 300),
    ]
    print("# BER / FER sweeps, GPU (libldpc_b200, fp32) vs CPU oracle (fp64), round 1\n")
    print("All-zero codeword, BPSK + AWGN, sigma^2 = 1/(2 R Eb/N0) with R = K/n_tx, errors counted on the first K bits; "
          "95 % Wilson intervals. `agree` = the CPU estimate lies inside the GPU interval widened by the CPU interval.\n")
    for name, algo, iters, n_tx, grid, gframes, cframes in configs:
        code = codes.load_builtin(name)
        tx = n_tx or code.N
        if quick:
            gframes, cframes = gframes // 10, max(40, cframes // 4)
        dc = lib.DeviceCode(code)
        dec = lib.Decoder(dc, algo, "fp32", max_batch=16384)
        orc = oracle.Oracle(code)
        print("## %s, %s, max %d iterations, R = %d/%d, engine %s\n" % (name, {"min": "min-sum (alpha=3/4)", "spa": "sum-product"}[algo], iters, code.K, tx,
                                                                     {1: "STREAM", 2: "RESIDENT", 3: "ONCHIP"}[dec.engine()]))
        print("| Eb/N0 dB | GPU frames | GPU FER [95 %] | GPU BER | GPU avg it | CPU frames | CPU FER [95 %] | CPU BER | CPU avg it | agree |")
        print("|---|---|---|---|---|---|---|---|---|---|")
        for eb in grid:
            c = dec.simulate(iters, eb, code.K, tx, 7, 0, gframes)
            glo, ghi = wilson(c["frame_errors"], c["frames"])
            fe, be, ait = cpu_point(code, orc, algo, iters, eb, tx, cframes, 1000 + int(eb * 100))
            clo, chi = wilson(fe, cframes)
            agree = not (chi < glo or clo > ghi)
            print("| %.2f | %d | %.3e [%.2e, %.2e] | %.3e | %.2f | %d | %.3e [%.2e, %.2e] | %.3e | %.2f | %s |" % (
                eb, c["frames"], c["frame_errors"] / c["frames"], glo, ghi, c["bit_errors"] / (c["frames"] * code.K),
                c["sum_iters"] / c["frames"], cframes, fe / cframes, clo, chi, be / (cframes * code.K), ait, "yes" if agree else "NO"))
            sys.stdout.flush()
        print()
        dec.close(); dc.close()


if __name__ == "__main__":
    main()
