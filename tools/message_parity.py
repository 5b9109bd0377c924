#!/usr/bin/env python3
"""Per-edge message parity of the production engines (message tap of ldpc_b200_decode_trace) against the float64 CPU
restatement, with the numbers behind the tolerances of tests/test_gpu_parity.py::test_message_tap_*.

    python tools/message_parity.py > profiles/r02_message_parity.md      (on a GPU box)

For every case: frames x first iterations; the error of every check->variable message and posterior, binned by the
magnitude of the reference value, as the maximum absolute and relative error per bin."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from conftest import awgn_llrs  # noqa: E402
from ecc_ldpc_b200 import codes, lib  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = [("jpl.4096.4.5", 5120, 3.5), ("jpl.1024.4.5", 1280, 3.7), ("1920.1280.3.303", None, 2.6), ("moon.7.13", None, 4.0)]
BINS = [(0, 1e-3), (1e-3, 1e-1), (1e-1, 1), (1, 10), (10, 30), (30, 1e9)]
lib.init(0)
print("# Per-edge message parity: float32 production engines vs float64 restatement (round 2)\n")
print("`ldpc_b200_decode_trace` on the decoder's own kernel (RESIDENT: TRACE instance; ONCHIP: run-time tap). 6 frames, first 6")
print("iterations, float32-rounded LLRs on both sides. Per magnitude bin of the REFERENCE value: entries, max absolute error, max")
print("relative error, share of entries within 1e-4 relative.\n")
for name, n_tx, ebn0 in CASES:
    code = codes.load_builtin(name)
    orc = oracle.Oracle(code)
    dc = lib.DeviceCode(code)
    for algo in ("min", "spa"):
        dec = lib.Decoder(dc, algo, "fp32", max_batch=4)
        eng = {1: "STREAM", 2: "RESIDENT", 3: "ONCHIP"}[dec.engine()]
        llr = awgn_llrs(code, ebn0 - (0.4 if algo == "spa" else 0.0), 6, 2718, n_tx).astype(np.float32).astype(np.float64)
        G, R = [], []
        for f in range(6):
            b, it, cv, c2v, lam = dec.decode_trace(llr[f], 8)
            rb, rit, rcv, rc2v, rlam = orc.decode(algo, 8, llr[f], variant=0, trace=True)
            n = min(it, rit, 6)
            if n:
                G += [c2v[:n].ravel(), lam[:n].ravel()]; R += [rc2v[:n].ravel(), rlam[:n].ravel()]
        g, r = np.concatenate(G), np.concatenate(R)
        err = np.abs(g - r)
        print("## %s, %s, %s, Eb/N0 %.1f dB: %d values; max |err| / max(|ref|, 1) = %.2e; within 1e-4 relative: %.4f\n" % (
            name, algo, eng, ebn0 - (0.4 if algo == "spa" else 0.0), r.size, (err / np.maximum(np.abs(r), 1)).max(), (err <= 1e-4 * np.abs(r)).mean()))
        print("| |ref| in | entries | max abs err | max rel err | within 1e-4 rel |")
        print("|---|---|---|---|---|")
        for lo, hi in BINS:
            m = (np.abs(r) >= lo) & (np.abs(r) < hi) & (r != 0)
            if m.any():
                print("| [%g, %g) | %d | %.2e | %.2e | %.4f |" % (lo, hi, m.sum(), err[m].max(), (err[m] / np.abs(r[m])).max(), (err[m] <= 1e-4 * np.abs(r[m])).mean()))
        z = r == 0
        if z.any():
            print("| exactly 0 | %d | %.2e | - | %.4f |" % (z.sum(), err[z].max(), (err[z] == 0).mean()))
        print()
        dec.close()
    dc.close()
