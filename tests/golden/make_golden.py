#!/usr/bin/env python3
"""Generates tests/golden/*.npz with oracle/dense_ref.py, the literal dense transcription of
Reference/Orig.hs:58-98 and Reference/Min.hs:54-104 (and, algo "cm", of Fast/CachedMult.hs:233-264).

The reference itself cannot run here (no GHC; ecc-manifold absent), so these are NOT reference outputs:
they pin the C oracle and the CUDA kernels to a second, independently written statement of the same
algorithm ("parity unpinned", see oracle/ldpc_oracle.c).  Deterministic: numpy PCG64 with fixed seeds.
Usage: python tests/golden/make_golden.py   (minutes; pure Python)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ecc_ldpc_b200 import codes  # noqa: E402
from oracle import arraylet_ref, dense_ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def encode(code, msg):
    g = code.g_dense()
    par = (msg.astype(np.int64) @ g.astype(np.int64)) % 2
    return np.concatenate([msg, par]).astype(np.uint8)


def make(code_name, algo, ebn0_db, max_iters, n_frames, seed, n_tx=None, keep="all"):
    code = codes.load_builtin(code_name)
    H = code.dense().tolist()
    rng = np.random.default_rng(seed)
    n_tx = code.N if n_tx is None else n_tx
    R = code.K / n_tx
    sigma = np.sqrt(1.0 / (2 * R * 10 ** (ebn0_db / 10)))
    llrs, bits, iters, conv, c2v_first, c2v_last, lam_first, lam_last, c2v_all, lam_all = ([] for _ in range(10))
    for f in range(n_frames):
        if code.g_row_ptr is not None:
            msg = rng.integers(0, 2, code.K).astype(np.uint8)
            cw = encode(code, msg)
        else:
            cw = np.zeros(code.N, dtype=np.uint8)
        y = (2.0 * cw - 1.0) + sigma * rng.standard_normal(code.N)
        llr = 2.0 * y / sigma ** 2
        llr[n_tx:] = 0.0                               # unpuncture, Utils.hs:55
        tr = []
        if algo in ("arraylet", "arraylet-min"):
            # third statement: the circulant-vector modules Fast/Arraylet.hs / Fast/ArrayletMin.hs (oracle/arraylet_ref.py)
            b, it, cv = arraylet_ref.ldpc(code.sz, code.shifts, "spa" if algo == "arraylet" else "min", max_iters, llr.tolist(), trace=tr)
        else:
            b, it, cv = dense_ref.ldpc(H, algo, max_iters, llr.tolist(), trace=tr)
        llrs.append(llr); bits.append(b); iters.append(it); conv.append(cv)
        def edge_vec(d):
            if algo in ("arraylet", "arraylet-min"):
                return np.array(arraylet_ref.matrixlet_to_edge_order(d))
            return np.array([d[(m, int(code.col_idx[e]))] for m in range(code.M)
                             for e in range(code.row_ptr[m], code.row_ptr[m + 1])])
        z_e, z_n = np.zeros(code.E), np.zeros(code.N)
        c2v_first.append(edge_vec(tr[0][0]) if tr else z_e)
        c2v_last.append(edge_vec(tr[-1][0]) if tr else z_e)
        lam_first.append(np.array(tr[0][1]) if tr else z_n)
        lam_last.append(np.array(tr[-1][1]) if tr else z_n)
        if keep == "all":
            ca = np.zeros((max_iters, code.E)); la = np.zeros((max_iters, code.N))
            for i, (ne, lam) in enumerate(tr):
                ca[i] = edge_vec(ne); la[i] = lam
            c2v_all.append(ca); lam_all.append(la)
        print(code_name, algo, "frame", f, "iters", it, "conv", cv, flush=True)
    name = "%s_%s_it%d.npz" % (code_name, algo, max_iters)
    extra = {}
    if keep == "all":
        extra = {"c2v_all": np.array(c2v_all), "lam_all": np.array(lam_all)}
    np.savez_compressed(os.path.join(OUT, name), llr=np.array(llrs), bits=np.array(bits, dtype=np.uint8),
                        iters=np.array(iters, dtype=np.int32), converged=np.array(conv, dtype=np.uint8),
                        c2v_first=np.array(c2v_first), c2v_last=np.array(c2v_last),
                        lam_first=np.array(lam_first), lam_last=np.array(lam_last),
                        ebn0_db=ebn0_db, n_tx=n_tx, max_iters=max_iters, seed=seed, **extra)
    print("wrote", name)


if __name__ == "__main__":
    which = sys.argv[1:] or ["moon", "jpl1024", "mackay", "jpl4096", "cm", "arraylet"]
    if "moon" in which:
        make("moon.7.13", "spa", 2.0, 20, 48, 11)
        make("moon.7.13", "min", 2.0, 20, 48, 12)
        make("moon.7.13", "spa", 5.0, 3, 16, 13)      # max_iters small: exercises the orig_lam fallback
    if "jpl1024" in which:
        make("jpl.1024.4.5", "min", 4.2, 12, 4, 21, n_tx=1280, keep="ends")
        make("jpl.1024.4.5", "spa", 3.9, 10, 3, 22, n_tx=1280, keep="ends")
    if "mackay" in which:
        make("1920.1280.3.303", "min", 2.6, 12, 3, 31, keep="ends")
        make("1920.1280.3.303", "spa", 2.0, 10, 2, 32, keep="ends")
    if "jpl4096" in which:
        make("jpl.4096.4.5", "min", 4.0, 10, 2, 41, n_tx=5120, keep="ends")
        make("jpl.4096.4.5", "spa", 3.6, 6, 1, 42, n_tx=5120, keep="ends")
    if "cm" in which:   # the cached-product sum-product of Fast/CachedMult.hs:233-264 (dense_ref.ldpc_cached_mult)
        make("moon.7.13", "cm", 2.0, 20, 24, 14)
        make("jpl.1024.4.5", "cm", 3.9, 8, 2, 23, n_tx=1280, keep="ends")
    if "arraylet" in which:   # `ldpc/arraylet-min` / `ldpc/arraylet` in their own circulant-vector form (arraylet_ref.ldpc)
        make("jpl.1024.4.5", "arraylet-min", 4.0, 12, 4, 51, n_tx=1280, keep="all")
        make("jpl.1024.4.5", "arraylet", 3.8, 10, 3, 52, n_tx=1280, keep="ends")
        make("jpl.4096.4.5", "arraylet-min", 3.8, 10, 1, 53, n_tx=5120, keep="ends")
