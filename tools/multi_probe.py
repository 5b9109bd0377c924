#!/usr/bin/env python3
"""ldpc_b200_simulate_multi on all visible GPUs from ONE process (the Haskell tester is one process): counters must equal
the single-GPU run of the same global frame range; throughput printed.   python tools/multi_probe.py [frames] [ebn0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecc_ldpc_b200 import codes, lib
F = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000000
ebn0 = float(sys.argv[2]) if len(sys.argv) > 2 else 3.5
code = codes.load_builtin("jpl.4096.4.5")
lib.init(0)
n_dev = lib.load().ldpc_b200_device_count()
dc = lib.DeviceCode(code)
keys = ("frames", "frame_errors", "bit_errors", "sum_iters", "unconverged")
ref = None
for n in sorted({1, 2, n_dev} & set(range(1, n_dev + 1))):
    lib.simulate_multi(dc, n, "min", "fp32", 32768, 50, ebn0, code.K, 5120, 7, 0, 100000)      # warm-up (NCCL init, kernels)
    t0 = time.perf_counter()
    c = lib.simulate_multi(dc, n, "min", "fp32", 32768, 50, ebn0, code.K, 5120, 7, 1000, F)
    dt = time.perf_counter() - t0
    got = {k: int(c[k]) for k in keys}
    if ref is None:
        ref = got
    print("%d GPU(s): %.2f s  %.1f M frames/s  %.1f Gbit/s  %s  %s" % (n, dt, F / dt / 1e6, F / dt * code.K / 1e9, got,
                                                                      "== 1-GPU counters" if got == ref else "DIFFERS from 1 GPU"))
    assert got == ref
lib.load().ldpc_b200_simulate_multi_release()
