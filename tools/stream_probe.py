#!/usr/bin/env python3
"""One decode_device call on a chosen engine (for ncu launch lists and quick timing):
python tools/stream_probe.py [frames] [code] [algo] [stream|auto|onchip|resident] [ebn0]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecc_ldpc_b200 import codes, lib
F = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
name = sys.argv[2] if len(sys.argv) > 2 else "jpl.4096.4.5"
algo = sys.argv[3] if len(sys.argv) > 3 else "min"
eng = {"stream": lib.ENGINE_STREAM, "auto": lib.ENGINE_AUTO, "onchip": lib.ENGINE_ONCHIP, "resident": lib.ENGINE_RESIDENT}[sys.argv[4] if len(sys.argv) > 4 else "stream"]
code = codes.load_builtin(name)
n_tx = {"jpl.4096.4.5": 5120, "jpl.1024.4.5": 1280}.get(name, code.N)
lib.init(0)
dc = lib.DeviceCode(code)
dec = lib.Decoder(dc, algo, "fp32", max_batch=F, engine=eng)
llr = lib.DeviceBuffer(F * code.N * 4); bits = lib.DeviceBuffer(F * ((code.N + 31) // 32) * 4); it = lib.DeviceBuffer(F * 4); cv = lib.DeviceBuffer(F)
ebn0 = float(sys.argv[5]) if len(sys.argv) > 5 else (2.0 if name.startswith("jpl") else 0.0)
dec.generate_llr(llr.ptr, ebn0, code.K, n_tx, 1, 0, F); dec.sync()
for r in range(3):
    dec.decode_device(llr.ptr, bits.ptr, 20, F, early_exit=True, d_iters=it.ptr, d_conv=cv.ptr); dec.sync()
    print(name, algo, "engine", dec.engine(), "ms %.3f" % dec.last_timing()[0], "launches", dec.last_timing()[1], "avg_it %.2f" % it.download(np.int32, (F,)).mean())
