"""ctypes view of oracle/libldpc_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libldpc_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("ldpc_oracle.c", "arraylet_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_code_create.restype = C.c_void_p
        L.oracle_code_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_code_destroy.argtypes = [C.c_void_p]
        L.oracle_decode.restype = C.c_int
        L.oracle_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        L.oracle_decode_batch.restype = C.c_int
        L.oracle_decode_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_ecc_decode.restype = C.c_int
        L.oracle_ecc_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_encode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_max_threads.restype = C.c_int
        # arraylet_oracle.c: the circulant-vector statement of Fast/Arraylet.hs / Fast/ArrayletMin.hs
        L.arraylet_code_create.restype = C.c_void_p
        L.arraylet_code_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.arraylet_code_destroy.argtypes = [C.c_void_p]
        L.arraylet_decode.restype = C.c_int
        L.arraylet_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
                                      C.c_void_p, C.c_void_p]
        L.arraylet_decode_batch.restype = C.c_int
        L.arraylet_decode_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int]
        _lib = L
    return _lib


# 3 / 4: the check updates of 0 / 1 with the posterior association of the arraylet modules (edge-list form)
ALGO = {"spa": 0, "min": 1, "minsum": 1, "spa_cm": 2, "cm": 2, "arraylet": 3, "spa_arraylet": 3,
        "arraylet-min": 4, "min_arraylet": 4, 0: 0, 1: 1, 2: 2, 3: 3, 4: 4}


class Oracle:
    def __init__(self, code):
        self.code = code
        self._rp = np.ascontiguousarray(code.row_ptr, dtype=np.int32)
        self._ci = np.ascontiguousarray(code.col_idx, dtype=np.int32)
        self._h = lib().oracle_code_create(code.M, code.N, self._rp.ctypes.data, self._ci.ctypes.data)

    def __del__(self):
        try:
            lib().oracle_code_destroy(self._h)
        except Exception:
            pass

    def decode(self, algo, max_iters, llr, variant=0, trace=False):
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        N, E = self.code.N, self.code.E
        assert llr.shape == (N,)
        bits = np.zeros(N, dtype=np.uint8)
        conv = C.c_int(0)
        tc = tl = None
        if trace:
            tc = np.zeros((max(max_iters, 1), E), dtype=np.float64)
            tl = np.zeros((max(max_iters, 1), N), dtype=np.float64)
        it = lib().oracle_decode(self._h, ALGO[algo], variant, max_iters, llr.ctypes.data, bits.ctypes.data,
                                 C.byref(conv), tc.ctypes.data if trace else None, tl.ctypes.data if trace else None)
        if trace:
            return bits, it, bool(conv.value), tc[:it], tl[:it]
        return bits, it, bool(conv.value)

    def decode_batch(self, algo, max_iters, llr, variant=0, threads=0):
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        F, N = llr.shape
        assert N == self.code.N
        bits = np.zeros((F, N), dtype=np.uint8)
        iters = np.zeros(F, dtype=np.int32)
        conv = np.zeros(F, dtype=np.uint8)
        used = lib().oracle_decode_batch(self._h, ALGO[algo], variant, max_iters, F, llr.ctypes.data,
                                         bits.ctypes.data, iters.ctypes.data, conv.ctypes.data, threads)
        self.threads_used = used
        return bits, iters, conv

    def ecc_decode(self, algo, max_iters, K, c_length, rx, variant=0):
        rx = np.ascontiguousarray(rx, dtype=np.float64)
        out = np.zeros(K, dtype=np.uint8)
        it = C.c_int(0); cv = C.c_int(0)
        flag = lib().oracle_ecc_decode(self._h, ALGO[algo], variant, max_iters, K, c_length, rx.ctypes.data,
                                       len(rx), out.ctypes.data, C.byref(it), C.byref(cv))
        return out, bool(flag), it.value, bool(cv.value)


class ArrayletOracle:
    """oracle/arraylet_oracle.c: `ldpc/arraylet` ("spa") and `ldpc/arraylet-min` ("min") on a quasi-cyclic code."""

    def __init__(self, code):
        assert code.is_qc
        self.code = code
        self._sh = np.ascontiguousarray(code.shifts, dtype=np.int32)
        self._h = lib().arraylet_code_create(code.sz, code.brows, code.bcols, self._sh.ctypes.data)

    def __del__(self):
        try:
            lib().arraylet_code_destroy(self._h)
        except Exception:
            pass

    def decode(self, algo, max_iters, llr, trace=False):
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        N, E = self.code.N, self.code.E
        bits = np.zeros(N, dtype=np.uint8)
        conv = C.c_int(0)
        tc = tl = None
        if trace:
            tc = np.zeros((max(max_iters, 1), E), dtype=np.float64)
            tl = np.zeros((max(max_iters, 1), N), dtype=np.float64)
        it = lib().arraylet_decode(self._h, {"spa": 0, "min": 1}[algo], max_iters, llr.ctypes.data, bits.ctypes.data,
                                   C.byref(conv), tc.ctypes.data if trace else None, tl.ctypes.data if trace else None)
        if trace:
            return bits, it, bool(conv.value), tc[:it], tl[:it]
        return bits, it, bool(conv.value)

    def decode_batch(self, algo, max_iters, llr, threads=0):
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        F, N = llr.shape
        bits = np.zeros((F, N), dtype=np.uint8)
        iters = np.zeros(F, dtype=np.int32)
        conv = np.zeros(F, dtype=np.uint8)
        lib().arraylet_decode_batch(self._h, {"spa": 0, "min": 1}[algo], max_iters, F, llr.ctypes.data, bits.ctypes.data,
                                    iters.ctypes.data, conv.ctypes.data, threads or lib().oracle_max_threads())
        return bits, iters, conv


def encode(code, msg, c_length=None):
    """Orig.hs:25-26 + Utils.hs:61."""
    msg = np.ascontiguousarray(msg, dtype=np.uint8)
    c_length = code.N if c_length is None else c_length
    cw = np.zeros(c_length, dtype=np.uint8)
    rp = np.ascontiguousarray(code.g_row_ptr, dtype=np.int32)
    ci = np.ascontiguousarray(code.g_col_idx, dtype=np.int32)
    lib().oracle_encode(code.K, code.N - code.K, rp.ctypes.data, ci.ctypes.data, c_length, msg.ctypes.data, cw.ctypes.data)
    return cw
