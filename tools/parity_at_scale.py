#!/usr/bin/env python3
"""Frame-by-frame parity of the float32 production engines with the float64 CPU restatement at scale (the north-star bar:
hard decisions identical on >= 99.9 % of the frames), in the waterfall where frames converge late or fail.

    python tools/parity_at_scale.py [seconds_of_cpu_per_case] > profiles/r01_parity_at_scale.md

LLRs are generated on the GPU (Philox), downloaded, and decoded by both sides; the CPU side (oracle/, test infrastructure)
uses all host cores, so the sample size is what the CPU finishes in the given time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from ecc_ldpc_b200 import codes, lib  # noqa: E402
from oracle import oracle  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 15.0
CASES = [("jpl.4096.4.5", "min", 20, 5120, 3.25), ("jpl.4096.4.5", "spa", 20, 5120, 3.0), ("jpl.1024.4.5", "min", 20, 1280, 3.4),
         ("jpl.1024.4.5", "spa", 20, 1280, 3.2), ("1920.1280.3.303", "min", 20, 1920, 1.8), ("1920.1280.3.303", "spa", 20, 1920, 1.5),
         ("moon.7.13", "spa", 20, 20, 2.0)]
lib.init(0)
print("# Parity at scale: float32 production engines vs float64 CPU restatement on identical LLRs (%s)\n" % (sys.argv[2] if len(sys.argv) > 2 else "round 1"))
print("| code | algorithm | Eb/N0 dB | engine | frames | identical frames | fraction | differing bits in differing frames (mean) | iteration counts equal | GPU FER | CPU FER |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, algo, iters, n_tx, ebn0 in CASES:
    code = codes.load_builtin(name)
    orc = oracle.Oracle(code)
    dc = lib.DeviceCode(code)
    dec = lib.Decoder(dc, algo, "fp32", max_batch=8192)
    # calibrate the CPU rate on a small sample, then size the batch for the time budget
    probe = 256
    buf = lib.DeviceBuffer(probe * code.N * 4)
    dec.generate_llr(buf.ptr, ebn0, code.K, n_tx, 424242, 0, probe); dec.sync()
    llr = buf.download(np.float32, (probe, code.N)).astype(np.float64)
    t0 = time.perf_counter(); orc.decode_batch(algo, iters, llr, variant=1 if algo == "min" else 0); rate = probe / (time.perf_counter() - t0)
    F = int(min(2000000, max(2048, rate * budget)) // 1024 * 1024)
    same = eq_it = fe_g = fe_c = 0; diffbits = []
    for f0 in range(0, F, 8192):
        n = min(8192, F - f0)
        buf = lib.DeviceBuffer(n * code.N * 4)
        dec.generate_llr(buf.ptr, ebn0, code.K, n_tx, 424242, 1000 + f0, n); dec.sync()
        llr = buf.download(np.float32, (n, code.N)).astype(np.float64)
        gb, gi, gc = dec.decode(llr, iters)
        rb, ri, rc = orc.decode_batch(algo, iters, llr, variant=1 if algo == "min" else 0)
        s = (gb == rb).all(axis=1)
        same += int(s.sum()); eq_it += int((gi == ri).sum())
        fe_g += int(gb[:, :code.K].any(axis=1).sum()); fe_c += int(rb[:, :code.K].any(axis=1).sum())
        diffbits += list((gb[~s] != rb[~s]).sum(axis=1))
    print("| %s | %s | %.2f | %s | %d | %d | %.6f | %s | %.6f | %.3e | %.3e |" % (
        name, algo, ebn0, {1: "STREAM", 2: "RESIDENT", 3: "ONCHIP"}[dec.engine()], F, same, same / F,
        ("%.1f" % np.mean(diffbits)) if diffbits else "-", eq_it / F, fe_g / F, fe_c / F))
    sys.stdout.flush()
    dec.close(); dc.close()
