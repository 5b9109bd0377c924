#!/usr/bin/env python3
"""Deep FER / BER points on the GPU alone (ldpc_b200_simulate; the CPU restatement cannot reach these frame counts):
error-floor check of the float32 production engines far down the waterfall.

    python tools/fer_deep.py [scale] > profiles/r01_fer_deep.md      (scale multiplies the frame counts, default 1)

All-zero codeword, BPSK + AWGN, sigma^2 = 1/(2 R Eb/N0), errors over the first K bits, 95 % Wilson intervals.
Frames come from Philox streams keyed by (seed, global frame index), so any point can be re-run or extended.
"""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ecc_ldpc_b200 import codes, lib  # noqa: E402
from ber_sweep import wilson  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    lib.init(0)
    configs = [
        ("jpl.4096.4.5", "min", 50, 5120, [(3.5, 2e7), (3.75, 1e8), (4.0, 2e8), (4.25, 2e8)]),
        ("jpl.4096.4.5", "spa", 50, 5120, [(3.25, 1e7), (3.5, 5e7), (3.75, 5e7)]),
        ("jpl.1024.4.5", "min", 50, 1280, [(4.0, 1e8), (4.5, 5e8), (5.0, 5e8)]),
        ("1920.1280.3.303", "min", 50, None, [(2.0, 2e7), (2.25, 1e8), (2.5, 2e8)]),
        ("1920.1280.3.303", "spa", 50, None, [(1.75, 2e7), (2.0, 5e7)]),
    ]
    print("# Deep FER / BER points, GPU only (libldpc_b200 float32 production engines), round %s\n" % os.environ.get("LDPC_ROUND", "2"))
    print(__doc__.split("\n\n")[2].strip() + "\n")
    for name, algo, iters, n_tx, pts in configs:
        code = codes.load_builtin(name)
        tx = n_tx or code.N
        dc = lib.DeviceCode(code)
        dec = lib.Decoder(dc, algo, "fp32", max_batch=16384)
        print("## %s, %s, max %d iterations, R = %d/%d, engine %s\n" % (name, {"min": "min-sum (alpha=3/4)", "spa": "sum-product"}[algo], iters, code.K, tx,
                                                                     {1: "STREAM", 2: "RESIDENT", 3: "ONCHIP"}[dec.engine()]))
        print("| Eb/N0 dB | frames | frame errors | FER [95 %] | bit errors | BER | avg it | unconverged | M frames/s | info Gbit/s |")
        print("|---|---|---|---|---|---|---|---|---|---|")
        for eb, n in pts:
            n = int(n * scale)
            t0 = time.perf_counter()
            c = dec.simulate(iters, eb, code.K, tx, 20261018, 0, n)
            dt = time.perf_counter() - t0
            lo, hi = wilson(c["frame_errors"], c["frames"])
            print("| %.2f | %.2e | %d | %.2e [%.1e, %.1e] | %d | %.2e | %.2f | %d | %.2f | %.1f |" % (
                eb, c["frames"], c["frame_errors"], c["frame_errors"] / c["frames"], lo, hi, c["bit_errors"],
                c["bit_errors"] / (c["frames"] * code.K), c["sum_iters"] / c["frames"], c["unconverged"], n / dt / 1e6, n / dt * code.K / 1e9))
            sys.stdout.flush()
        print()
        dec.close(); dc.close()


if __name__ == "__main__":
    main()
