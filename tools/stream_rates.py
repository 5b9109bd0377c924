#!/usr/bin/env python3
"""STREAM engine rates against the HBM roof (algorithmic bytes 3E+3N words per frame and iteration): python tools/stream_rates.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecc_ldpc_b200 import codes, lib
lib.init(0)
for name, ntx, F in (("jpl.4096.4.5", 5120, 8192), ("jpl.4096.4.5", 5120, 32768), ("1920.1280.3.303", 1920, 32768)):
    code = codes.load_builtin(name); dc = lib.DeviceCode(code)
    for algo in ("min", "spa"):
        dec = lib.Decoder(dc, algo, "fp32", max_batch=F, engine=lib.ENGINE_STREAM)
        llr = lib.DeviceBuffer(F * code.N * 4); bits = lib.DeviceBuffer(F * ((code.N + 31) // 32) * 4); it = lib.DeviceBuffer(F * 4); cv = lib.DeviceBuffer(F)
        for eb in (0.0 if "1920" in name else 2.0, 3.0 if "1920" in name else 4.0):
            dec.generate_llr(llr.ptr, eb, code.K, ntx, 1, 0, F); dec.sync()
            ms = []
            for r in range(6):
                dec.decode_device(llr.ptr, bits.ptr, 20, F, early_exit=True, d_iters=it.ptr, d_conv=cv.ptr); dec.sync(); ms.append(dec.last_timing()[0])
            m = float(np.median(ms[2:])); ai = it.download(np.int32, (F,)).mean()
            ba = ai * 4 * (3 * code.E + 3 * code.N) + 4 * ntx + code.K // 8
            print(name, algo, "F", F, "ebn0", eb, "%.3f ms" % m, "%.3f Mframes/s" % (F / m / 1e3), "avg_it %.2f" % ai, "alg GB/s %.0f" % (F / m * 1e3 * ba / 1e9),
                  "frac %.3f" % (F / m * 1e3 * ba / 1e9 / 6552.3), "launches", dec.last_timing()[1], flush=True)
        dec.close()
        for b in (llr, bits, it, cv): b.free()
    dc.close()
