#!/usr/bin/env python3
"""Sweep of the tester pattern (ldpc_b200_ecc_tester_pattern): threads x Eb/N0, for tuning the mailbox.
    python tools/tester_pattern_probe.py [code-name] [threads,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ecc_ldpc_b200 import codes, lib, philox_host
name = sys.argv[1] if len(sys.argv) > 1 else "ldpc/b200-minsum/jpl.4096.4.5/20/4/5"
threads = [int(t) for t in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8,16,32").split(",")]
lib.init(0)
code = codes.load_builtin(name.split("/")[2])
n_tx = {"jpl.4096.4.5": 5120, "jpl.1024.4.5": 1280}.get(code.name, code.N)
for ebn0 in (2.0, 4.0):
    rx = philox_host.awgn_llrs(code.N, ebn0, code.K, n_tx, 7, 0, 4096).astype(np.float64)[:, :n_tx]
    row = []
    for T in threads:
        nf = 4096 if T >= 4 else 1024
        msg, secs, nl, nfr = lib.tester_pattern(codes.builtin_codes_dir(), name, T, rx[:nf])
        row.append("T=%d: %.0f f/s (%.1f us/frame/thread, %.1f frames/launch)" % (T, nf / secs, secs / (nf / T) * 1e6, nfr / max(nl, 1)))
    print("MB_WAIT=%s MB_SPIN=%s ebn0 %.1f | " % (os.environ.get("LDPC_B200_MB_WAIT", "-"), os.environ.get("LDPC_B200_MB_SPIN", "-"), ebn0) + " | ".join(row), flush=True)
