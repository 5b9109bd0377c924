import os, sys, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecc_ldpc_b200 import codes, lib
F = 16384
code = codes.load_builtin("jpl.4096.4.5")
lib.init(0)
dc = lib.DeviceCode(code)
llr = lib.DeviceBuffer(F * code.N * 4); bits = lib.DeviceBuffer(F * (code.N // 32) * 4); it = lib.DeviceBuffer(F * 4); cv = lib.DeviceBuffer(F)
def smi():
    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    return o
for algo in sys.argv[1:] or ("spa", "min"):
    dec = lib.Decoder(dc, algo, "fp32", max_batch=F)
    for ebn0 in (2.0, 3.0, 6.0):
        dec.generate_llr(llr.ptr, ebn0, code.K, 5120, 1, 0, F); dec.sync()
        ms = []
        t0 = time.time(); clk = None
        while time.time() - t0 < 1.5:
            for r in range(10):
                dec.decode_device(llr.ptr, bits.ptr, 20, F, early_exit=False, d_iters=it.ptr, d_conv=cv.ptr)
            if clk is None and time.time() - t0 > 0.7:
                clk = smi()
            dec.sync()
            ms.append(dec.last_timing()[0])
        print(algo, ebn0, "ms %.3f" % np.median(ms), "|", clk, flush=True)
