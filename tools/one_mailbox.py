#!/usr/bin/env python3
"""A few one-frame host-buffer decodes through the mailbox (latency kernel), for ncu and for timing:
python tools/one_mailbox.py [algo] [ebn0] [calls]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecc_ldpc_b200 import codes, lib, philox_host
algo = sys.argv[1] if len(sys.argv) > 1 else "min"
ebn0 = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 200
code = codes.load_builtin("jpl.4096.4.5")
lib.init(0)
dc = lib.DeviceCode(code)
dec = lib.Decoder(dc, algo, "fp32", max_batch=1)
llr = philox_host.awgn_llrs(code.N, ebn0, code.K, 5120, 3, 0, 16).astype(np.float64)
for r in range(8):
    dec.decode(llr[r % 16:r % 16 + 1], 20)
t0 = time.perf_counter()
its = 0
for r in range(calls):
    b, it, cv = dec.decode(llr[r % 16:r % 16 + 1], 20)
    its += int(it[0])
us = (time.perf_counter() - t0) / calls * 1e6
print("mailbox one-frame decode (%s, %.1f dB): %.1f us per call incl. ctypes, avg %.1f iterations, stats %s" % (algo, ebn0, us, its / calls, dec.mailbox_stats()))
