#!/usr/bin/env python3
"""One configuration, a few decode_device launches (for ncu): python tools/one_decode.py <algo> <ebn0> [frames] [early_exit]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecc_ldpc_b200 import codes, lib
algo, ebn0 = sys.argv[1], float(sys.argv[2])
F = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
ee = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
code = codes.load_builtin("jpl.4096.4.5")
lib.init(0)
dc = lib.DeviceCode(code)
dec = lib.Decoder(dc, algo, "fp32", max_batch=F)
llr = lib.DeviceBuffer(F * code.N * 4); bits = lib.DeviceBuffer(F * (code.N // 32) * 4); it = lib.DeviceBuffer(F * 4); cv = lib.DeviceBuffer(F)
dec.generate_llr(llr.ptr, ebn0, code.K, 5120, 1, 0, F); dec.sync()
for r in range(3):
    dec.decode_device(llr.ptr, bits.ptr, 20, F, early_exit=ee, d_iters=it.ptr, d_conv=cv.ptr); dec.sync()
    print(algo, ebn0, "ms %.3f" % dec.last_timing()[0], "avg_it %.2f" % it.download(np.int32, (F,)).mean())
