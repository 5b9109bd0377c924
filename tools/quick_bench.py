#!/usr/bin/env python3
"""Kernel-only timing of the RESIDENT decode for tuning: python tools/quick_bench.py [frames] [code]
(LDPC_B200_SO selects a library variant)."""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecc_ldpc_b200 import codes, lib

F = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
name = sys.argv[2] if len(sys.argv) > 2 else "jpl.4096.4.5"
algo = sys.argv[3] if len(sys.argv) > 3 else "min"
code = codes.load_builtin(name)
n_tx = {"jpl.4096.4.5": 5120, "jpl.1024.4.5": 1280}[name]
lib.init(0)
dc = lib.DeviceCode(code)
dec = lib.Decoder(dc, algo, "fp32", max_batch=F)
llr = lib.DeviceBuffer(F * code.N * 4)
bits = lib.DeviceBuffer(F * (code.N // 32) * 4)
it = lib.DeviceBuffer(F * 4)
cv = lib.DeviceBuffer(F)
out = []
for ebn0, ee in ((2.0, True), (4.0, True)):
    dec.generate_llr(llr.ptr, ebn0, code.K, n_tx, 1, 0, F)
    dec.sync()
    ms = []
    for r in range(8):
        dec.decode_device(llr.ptr, bits.ptr, 20, F, early_exit=ee, d_iters=it.ptr, d_conv=cv.ptr)
        dec.sync()
        ms.append(dec.last_timing()[0])
    m = float(np.median(ms[3:]))
    its = it.download(np.int32, (F,))
    avg_it = its.mean()
    crc = zlib.crc32(its.tobytes(), zlib.crc32(bits.download(np.uint32, (F, code.N // 32)).tobytes()))   # same decisions <=> same crc (variants)
    out.append("ebn0 %.1f: %.3f ms  %.3f Mframes/s  %.2f Gbit/s  avg_it %.2f crc %08x" % (ebn0, m, F / m / 1e3, F / m / 1e3 * code.K / 1e3, avg_it, crc))
print(os.environ.get("LDPC_B200_SO", "default"), name, algo, "engine", dec.engine(), "|", " | ".join(out))
