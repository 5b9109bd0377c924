#!/usr/bin/env python3
"""Per-call latency of the one-frame-per-call pattern of the reference tester (Utils.hs:68-69): python tools/latency_probe.py [ebn0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecc_ldpc_b200 import codes, lib, ecc
ebn0 = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
code = codes.load_builtin("jpl.4096.4.5")
lib.init(0)
dc = lib.DeviceCode(code)
dec = lib.Decoder(dc, "min", "fp32", max_batch=64)
llr = lib.DeviceBuffer(64 * code.N * 4); bits = lib.DeviceBuffer(64 * (code.N // 32) * 4); it = lib.DeviceBuffer(64 * 4); cv = lib.DeviceBuffer(64)
dec.generate_llr(llr.ptr, ebn0, code.K, 5120, 1, 0, 64); dec.sync()
for nf in (1, 8, 64):
    for r in range(20):
        dec.decode_device(llr.ptr, bits.ptr, 20, nf, early_exit=True, d_iters=it.ptr, d_conv=cv.ptr); dec.sync()
    t0 = time.perf_counter()
    for r in range(200):
        dec.decode_device(llr.ptr, bits.ptr, 20, nf, early_exit=True, d_iters=it.ptr, d_conv=cv.ptr); dec.sync()
    us = (time.perf_counter() - t0) / 200 * 1e6
    print("decode_device n_frames=%d: %.1f us per call (kernel %.1f us), avg_it %.1f" % (nf, us, dec.last_timing()[0] * 1e3, it.download(np.int32, (64,))[:nf].mean()))
(rec,) = ecc.match("ldpc/b200-minsum/jpl.4096.4.5/20/4/5", max_batch=64)
rx = llr.download(np.float32, (64, code.N))[:, :5120].astype(np.float64)
for nf in (1, 8, 64):
    for r in range(20):
        rec.decode_batch(rx[:nf])
    t0 = time.perf_counter()
    for r in range(200):
        rec.decode_batch(rx[:nf])
    us = (time.perf_counter() - t0) / 200 * 1e6
    print("ECC.decode n_frames=%d: %.1f us per call" % (nf, us))
