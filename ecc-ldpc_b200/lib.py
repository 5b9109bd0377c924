"""ctypes binding of libldpc_b200.so (include/ldpc_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present, the calls raise.
The binding is deliberately thin -- every method maps 1:1 onto an exported C function.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LDPC_B200_SO") or os.path.join(_HERE, "libldpc_b200.so")   # env override: kernel-variant experiments

ALGO_SPA, ALGO_MINSUM, ALGO_SPA_CM, ALGO_SPA_ARRAYLET, ALGO_MINSUM_ARRAYLET = 0, 1, 2, 3, 4
PREC_FP64, PREC_FP32 = 0, 1
ENGINE_AUTO, ENGINE_STREAM, ENGINE_RESIDENT, ENGINE_ONCHIP = 0, 1, 2, 3
ALGO = {"spa": ALGO_SPA, "min": ALGO_MINSUM, "minsum": ALGO_MINSUM, "spa_cm": ALGO_SPA_CM, "cm": ALGO_SPA_CM,
        "arraylet": ALGO_SPA_ARRAYLET, "arraylet-min": ALGO_MINSUM_ARRAYLET, 0: 0, 1: 1, 2: 2, 3: 3, 4: 4}
PREC = {"fp64": PREC_FP64, "fp32": PREC_FP32, 0: 0, 1: 1}

EXPORTS = [
    "ldpc_b200_abi_version", "ldpc_b200_last_error", "ldpc_b200_init", "ldpc_b200_shutdown", "ldpc_b200_live_handles",
    "ldpc_b200_device_count", "ldpc_b200_code_create_csr", "ldpc_b200_code_create_qc",
    "ldpc_b200_code_destroy", "ldpc_b200_code_info", "ldpc_b200_decoder_create",
    "ldpc_b200_decoder_set_engine", "ldpc_b200_decoder_engine", "ldpc_b200_decoder_destroy",
    "ldpc_b200_decode", "ldpc_b200_decode_punctured", "ldpc_b200_decode_f32", "ldpc_b200_decode_trace", "ldpc_b200_decode_device",
    "ldpc_b200_decoder_sync", "ldpc_b200_decoder_last_timing", "ldpc_b200_decoder_mailbox_stats", "ldpc_b200_generate_llr",
    "ldpc_b200_simulate", "ldpc_b200_simulate_device", "ldpc_b200_simulate_multi", "ldpc_b200_simulate_multi_release",
    "ldpc_b200_count_errors", "ldpc_b200_malloc", "ldpc_b200_free",
    "ldpc_b200_malloc_host", "ldpc_b200_free_host", "ldpc_b200_memcpy_h2d", "ldpc_b200_memcpy_d2h",
    "ldpc_b200_memset", "ldpc_b200_decoder_stream", "ldpc_b200_philox4x32",
]


ECC_EXPORTS = [
    "ldpc_b200_matrix_load", "ldpc_b200_matrix_parse", "ldpc_b200_matrix_destroy", "ldpc_b200_matrix_info",
    "ldpc_b200_matrix_csr", "ldpc_b200_matrix_qc_shifts", "ldpc_b200_ecc_open", "ldpc_b200_ecc_close",
    "ldpc_b200_ecc_name", "ldpc_b200_ecc_message_length", "ldpc_b200_ecc_codeword_length",
    "ldpc_b200_ecc_unpunctured_length", "ldpc_b200_ecc_max_iters", "ldpc_b200_ecc_encode", "ldpc_b200_ecc_decode",
    "ldpc_b200_ecc_run", "ldpc_b200_ber_line", "ldpc_b200_ecc_encode_device", "ldpc_b200_ecc_sync",
    "ldpc_b200_ecc_tester_pattern",
]
ENOMATCH, ENOENT = -6, -7


class Counters(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("frame_errors", C.c_uint64), ("bit_errors", C.c_uint64),
                ("sum_iters", C.c_uint64), ("unconverged", C.c_uint64), ("kernel_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class LdpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libldpc_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load the shared library (no device needed for this). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError("libldpc_b200.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "-- there is no CPU fallback" % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, i32, u64, f64 = C.c_void_p, C.c_int, C.c_uint64, C.c_double
    L.ldpc_b200_last_error.restype = C.c_char_p
    L.ldpc_b200_shutdown.restype = None
    L.ldpc_b200_live_handles.argtypes = [C.POINTER(i32), C.POINTER(i32)]
    L.ldpc_b200_code_create_csr.argtypes = [i32, i32, vp, vp, C.POINTER(vp)]
    L.ldpc_b200_code_create_qc.argtypes = [i32, i32, i32, vp, C.POINTER(vp)]
    L.ldpc_b200_code_destroy.argtypes = [vp]
    L.ldpc_b200_code_destroy.restype = None
    L.ldpc_b200_code_info.argtypes = [vp] + [C.POINTER(i32)] * 4
    L.ldpc_b200_decoder_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
    L.ldpc_b200_decoder_set_engine.argtypes = [vp, i32]
    L.ldpc_b200_decoder_engine.argtypes = [vp]
    L.ldpc_b200_decoder_destroy.argtypes = [vp]
    L.ldpc_b200_decoder_destroy.restype = None
    L.ldpc_b200_decode.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.ldpc_b200_decode_f32.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.ldpc_b200_decode_punctured.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32]
    L.ldpc_b200_decode_trace.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.ldpc_b200_decode_device.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    L.ldpc_b200_decoder_sync.argtypes = [vp]
    L.ldpc_b200_decoder_last_timing.argtypes = [vp, C.POINTER(f64), C.POINTER(C.c_int64)]
    L.ldpc_b200_decoder_mailbox_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.ldpc_b200_generate_llr.argtypes = [vp, f64, i32, i32, u64, u64, i32, vp, vp]
    L.ldpc_b200_simulate.argtypes = [vp, i32, i32, f64, i32, i32, u64, u64, u64, C.POINTER(Counters)]
    L.ldpc_b200_simulate_device.argtypes = [vp, i32, i32, f64, i32, i32, u64, u64, u64, vp]
    L.ldpc_b200_count_errors.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    L.ldpc_b200_simulate_multi.argtypes = [vp, i32, i32, i32, i32, i32, i32, f64, i32, i32, u64, u64, u64, C.POINTER(Counters)]
    L.ldpc_b200_simulate_multi_release.restype = None
    L.ldpc_b200_malloc.argtypes = [C.POINTER(vp), u64]
    L.ldpc_b200_free.argtypes = [vp]
    L.ldpc_b200_malloc_host.argtypes = [C.POINTER(vp), u64]
    L.ldpc_b200_free_host.argtypes = [vp]
    L.ldpc_b200_memcpy_h2d.argtypes = [vp, vp, u64]
    L.ldpc_b200_memcpy_d2h.argtypes = [vp, vp, u64]
    L.ldpc_b200_memset.argtypes = [vp, i32, u64]
    L.ldpc_b200_decoder_stream.argtypes = [vp]
    L.ldpc_b200_decoder_stream.restype = vp
    L.ldpc_b200_philox4x32.argtypes = [vp, vp, vp]
    # host layer (include/ldpc_b200_ecc.h)
    cs = C.c_char_p
    L.ldpc_b200_matrix_load.argtypes = [cs, cs, C.POINTER(vp)]
    L.ldpc_b200_matrix_parse.argtypes = [cs, cs, C.POINTER(vp)]
    L.ldpc_b200_matrix_destroy.argtypes = [vp]
    L.ldpc_b200_matrix_destroy.restype = None
    L.ldpc_b200_matrix_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_int64), C.POINTER(i32)]
    L.ldpc_b200_matrix_csr.argtypes = [vp, vp, vp]
    L.ldpc_b200_matrix_qc_shifts.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), vp]
    L.ldpc_b200_ecc_open.argtypes = [cs, cs, i32, C.POINTER(vp)]
    L.ldpc_b200_ecc_close.argtypes = [vp]
    L.ldpc_b200_ecc_close.restype = None
    L.ldpc_b200_ecc_name.argtypes = [vp]
    L.ldpc_b200_ecc_name.restype = cs
    for fn in ("message_length", "codeword_length", "unpunctured_length", "max_iters"):
        getattr(L, "ldpc_b200_ecc_" + fn).argtypes = [vp]
    L.ldpc_b200_ecc_encode.argtypes = [vp, vp, vp]
    L.ldpc_b200_ecc_decode.argtypes = [vp, i32, vp, i32, vp, vp]
    L.ldpc_b200_ecc_run.argtypes = [vp, f64, u64, u64, u64, i32, C.POINTER(Counters)]
    L.ldpc_b200_ber_line.argtypes = [cs, i32, f64, cs, f64, u64, u64, i32]
    L.ldpc_b200_ecc_encode_device.argtypes = [vp, i32, vp, vp]
    L.ldpc_b200_ecc_sync.argtypes = [vp]
    L.ldpc_b200_ecc_tester_pattern.argtypes = [cs, cs, i32, i32, vp, i32, vp, C.POINTER(f64), C.POINTER(u64), C.POINTER(u64)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LdpcError(rc, load().ldpc_b200_last_error().decode(errors="replace"))


def init(device=-1):
    check(load().ldpc_b200_init(device))


def philox4x32(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    check(load().ldpc_b200_philox4x32(c.ctypes.data, k.ctypes.data, o.ctypes.data))
    return o


class DeviceBuffer:
    """Raw device allocation owned by the library (ldpc_b200_malloc)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(load().ldpc_b200_malloc(C.byref(p), self.nbytes))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(load().ldpc_b200_memcpy_h2d(self.ptr, arr.ctypes.data, arr.nbytes))
        return self

    def download(self, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(load().ldpc_b200_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes))
        return out

    def zero(self):
        check(load().ldpc_b200_memset(self.ptr, 0, self.nbytes))

    def free(self):
        if self.ptr:
            load().ldpc_b200_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceCode:
    """Device-resident H (ldpc_b200_code_t)."""

    def __init__(self, code, force_csr=False):
        L = load()
        self.code = code
        h = C.c_void_p()
        if code.is_qc and not force_csr:
            sh = np.ascontiguousarray(code.shifts, dtype=np.int32)
            check(L.ldpc_b200_code_create_qc(code.sz, code.brows, code.bcols, sh.ctypes.data, C.byref(h)))
        else:
            rp = np.ascontiguousarray(code.row_ptr, dtype=np.int32)
            ci = np.ascontiguousarray(code.col_idx, dtype=np.int32)
            check(L.ldpc_b200_code_create_csr(code.M, code.N, rp.ctypes.data, ci.ctypes.data, C.byref(h)))
        self.h = h.value

    def info(self):
        v = [C.c_int() for _ in range(4)]
        check(load().ldpc_b200_code_info(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def close(self):
        if self.h:
            load().ldpc_b200_code_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Decoder:
    """One ldpc_b200_decoder_t: the state behind one `decoder0 h` call of mkLDPC (Utils.hs:53)."""

    def __init__(self, dcode: DeviceCode, algo="min", precision="fp32", max_batch=1024, engine=ENGINE_AUTO):
        self.dcode = dcode
        self.N = dcode.code.N
        self.E = dcode.code.E
        self.W = (self.N + 31) // 32
        self.algo = ALGO[algo]
        self.precision = PREC[precision]
        self.max_batch = int(max_batch)
        h = C.c_void_p()
        check(load().ldpc_b200_decoder_create(dcode.h, self.algo, self.precision, self.max_batch, C.byref(h)))
        self.h = h.value
        if engine != ENGINE_AUTO:
            self.set_engine(engine)

    @property
    def np_dtype(self):
        return np.float64 if self.precision == PREC_FP64 else np.float32

    def set_engine(self, engine):
        check(load().ldpc_b200_decoder_set_engine(self.h, engine))

    def engine(self):
        return load().ldpc_b200_decoder_engine(self.h)

    def decode(self, llr, max_iters):
        """ldpc_b200_decode: float64 LLRs [F, N] -> (bits [F, N] u8, iters [F], converged [F])."""
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        if llr.ndim == 1:
            llr = llr[None, :]
        F, N = llr.shape
        assert N == self.N
        bits = np.zeros((F, N), dtype=np.uint8)
        iters = np.zeros(F, dtype=np.int32)
        conv = np.zeros(F, dtype=np.uint8)
        check(load().ldpc_b200_decode(self.h, max_iters, F, llr.ctypes.data, bits.ctypes.data, iters.ctypes.data, conv.ctypes.data))
        return bits, iters, conv

    def decode_f32(self, llr, max_iters):
        llr = np.ascontiguousarray(llr, dtype=np.float32)
        F, N = llr.shape
        assert N == self.N
        packed = np.zeros((F, self.W), dtype=np.uint32)
        iters = np.zeros(F, dtype=np.int32)
        conv = np.zeros(F, dtype=np.uint8)
        check(load().ldpc_b200_decode_f32(self.h, max_iters, F, llr.ctypes.data, packed.ctypes.data, iters.ctypes.data, conv.ctypes.data))
        return packed, iters, conv

    def decode_trace(self, llr, max_iters):
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        assert llr.shape == (self.N,)
        rows = max(max_iters, 1)
        c2v = np.zeros((rows, self.E), dtype=np.float64)
        lam = np.zeros((rows, self.N), dtype=np.float64)
        bits = np.zeros(self.N, dtype=np.uint8)
        it = C.c_int32(0)
        cv = C.c_uint8(0)
        check(load().ldpc_b200_decode_trace(self.h, max_iters, llr.ctypes.data, c2v.ctypes.data, lam.ctypes.data,
                                            C.byref(it), bits.ctypes.data, C.byref(cv)))
        return bits, it.value, bool(cv.value), c2v[:it.value], lam[:it.value]

    def decode_device(self, d_llr, d_bits, max_iters, n_frames, early_exit=True, d_iters=None, d_conv=None):
        check(load().ldpc_b200_decode_device(self.h, max_iters, 1 if early_exit else 0, n_frames, d_llr, d_bits, d_iters, d_conv))

    def sync(self):
        check(load().ldpc_b200_decoder_sync(self.h))

    def last_timing(self):
        ms = C.c_double(0)
        n = C.c_int64(0)
        check(load().ldpc_b200_decoder_last_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def mailbox_stats(self):
        """(launches, frames) served so far by the mailbox this handle shares with the other handles of its code."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(load().ldpc_b200_decoder_mailbox_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def generate_llr(self, d_llr, ebn0_db, k_info, n_tx, seed, first_frame, n_frames, d_codewords=None):
        check(load().ldpc_b200_generate_llr(self.h, ebn0_db, k_info, n_tx, seed, first_frame, n_frames, d_codewords, d_llr))

    def simulate(self, max_iters, ebn0_db, k_info, n_tx, seed, first_frame, n_frames, early_exit=True):
        c = Counters()
        check(load().ldpc_b200_simulate(self.h, max_iters, 1 if early_exit else 0, ebn0_db, k_info, n_tx, seed,
                                        first_frame, n_frames, C.byref(c)))
        return c.as_dict()

    def simulate_device(self, d_counters, max_iters, ebn0_db, k_info, n_tx, seed, first_frame, n_frames, early_exit=True):
        check(load().ldpc_b200_simulate_device(self.h, max_iters, 1 if early_exit else 0, ebn0_db, k_info, n_tx, seed,
                                               first_frame, n_frames, d_counters))

    def count_errors(self, d_counters, n_frames, k_info, d_bits, d_iters, d_conv, d_ref=None):
        check(load().ldpc_b200_count_errors(self.h, n_frames, k_info, d_bits, d_ref, d_iters, d_conv, d_counters))

    def stream(self):
        return load().ldpc_b200_decoder_stream(self.h)

    def close(self):
        if self.h:
            load().ldpc_b200_decoder_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def simulate_multi(dcode, n_gpus, algo, precision, max_batch, max_iters, ebn0_db, k_info, n_tx, seed, first_frame, n_frames,
                   early_exit=True):
    """ldpc_b200_simulate_multi: one process, n_gpus devices, NCCL all-reduce of the counters."""
    c = Counters()
    check(load().ldpc_b200_simulate_multi(dcode.h, n_gpus, ALGO[algo], PREC[precision], max_batch, max_iters, 1 if early_exit else 0,
                                          ebn0_db, k_info, n_tx, seed, first_frame, n_frames, C.byref(c)))
    return c.as_dict()


def unpack_bits(packed, N):
    """[F, W] uint32 -> [F, N] uint8 (bit j of a frame = word j//32, bit j%32)."""
    packed = np.ascontiguousarray(packed, dtype=np.uint32)
    b = np.unpackbits(packed.view(np.uint8).reshape(packed.shape[0], -1), axis=1, bitorder="little")
    return b[:, :N]


class HostMatrix:
    """A matrix file read by the C++ loader (ldpc_b200_matrix_*): Loader.hs / QuasiCyclic.hs / Alist.hs / Matlab.hs."""

    def __init__(self, handle):
        self.h = handle
        r, c, sz = C.c_int(), C.c_int(), C.c_int()
        ones = C.c_int64()
        check(load().ldpc_b200_matrix_info(self.h, C.byref(r), C.byref(c), C.byref(ones), C.byref(sz)))
        self.rows, self.cols, self.ones, self.sz = r.value, c.value, ones.value, sz.value

    @classmethod
    def load(cls, codes_root, path_noext):
        h = C.c_void_p()
        check(load().ldpc_b200_matrix_load(codes_root.encode(), path_noext.encode(), C.byref(h)))
        return cls(h.value)

    @classmethod
    def parse(cls, kind, text):
        h = C.c_void_p()
        check(load().ldpc_b200_matrix_parse(kind.encode(), text.encode(), C.byref(h)))
        return cls(h.value)

    def csr(self):
        rp = np.zeros(self.rows + 1, dtype=np.int32)
        ci = np.zeros(max(self.ones, 1), dtype=np.int32)
        check(load().ldpc_b200_matrix_csr(self.h, rp.ctypes.data, ci.ctypes.data))
        return rp, ci[:self.ones]

    def qc_shifts(self):
        br, bc = C.c_int(), C.c_int()
        check(load().ldpc_b200_matrix_qc_shifts(self.h, C.byref(br), C.byref(bc), None))
        sh = np.zeros(br.value * bc.value, dtype=np.int32)
        check(load().ldpc_b200_matrix_qc_shifts(self.h, C.byref(br), C.byref(bc), sh.ctypes.data))
        return sh.reshape(br.value, bc.value)

    def close(self):
        if self.h:
            load().ldpc_b200_matrix_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Ecc:
    """The ECC record of mkLDPC (Utils.hs:59-75) as built by the C++ host layer (ldpc_b200_ecc_*)."""

    def __init__(self, codes_root, name, max_batch=1024):
        h = C.c_void_p()
        rc = load().ldpc_b200_ecc_open(codes_root.encode(), name.encode(), max_batch, C.byref(h))
        check(rc)
        self.h = h.value
        L = load()
        self.name = L.ldpc_b200_ecc_name(self.h).decode()
        self.message_length = L.ldpc_b200_ecc_message_length(self.h)
        self.codeword_length = L.ldpc_b200_ecc_codeword_length(self.h)
        self.unpunctured_length = L.ldpc_b200_ecc_unpunctured_length(self.h)
        self.max_iters = L.ldpc_b200_ecc_max_iters(self.h)

    def encode(self, msg):
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        assert msg.shape == (self.message_length,)
        cw = np.zeros(self.codeword_length, dtype=np.uint8)
        check(load().ldpc_b200_ecc_encode(self.h, msg.ctypes.data, cw.ctypes.data))
        return cw

    def decode_batch(self, rx):
        rx = np.ascontiguousarray(rx, dtype=np.float64)
        if rx.ndim == 1:
            rx = rx[None, :]
        F, rl = rx.shape
        out = np.zeros((F, self.message_length), dtype=np.uint8)
        flags = np.zeros(F, dtype=np.uint8)
        check(load().ldpc_b200_ecc_decode(self.h, F, rx.ctypes.data, rl, out.ctypes.data, flags.ctypes.data))
        return out, flags

    def decode(self, rx):
        out, flags = self.decode_batch(rx)
        return out[0], bool(flags[0])

    def run(self, ebn0_db, n_frames, seed=1, first_frame=0, random_messages=False):
        c = Counters()
        check(load().ldpc_b200_ecc_run(self.h, ebn0_db, seed, first_frame, n_frames, 1 if random_messages else 0, C.byref(c)))
        return c.as_dict()

    def encode_device(self, d_msg, d_cw, n_frames):
        check(load().ldpc_b200_ecc_encode_device(self.h, n_frames, d_msg, d_cw))

    def sync(self):
        check(load().ldpc_b200_ecc_sync(self.h))

    def close(self):
        if self.h:
            load().ldpc_b200_ecc_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tester_pattern(codes_root, name, n_threads, rx):
    """ldpc_b200_ecc_tester_pattern: n_threads tester threads, one ECC.decode call per frame.  Returns (message bits
    [F, K], seconds, mailbox launches, mailbox frames)."""
    rx = np.ascontiguousarray(rx, dtype=np.float64)
    F, rl = rx.shape
    e = Ecc(codes_root, name, 1)
    K = e.message_length
    e.close()
    out = np.zeros((F, K), dtype=np.uint8)
    secs = C.c_double(0)
    nl, nf = C.c_uint64(0), C.c_uint64(0)
    check(load().ldpc_b200_ecc_tester_pattern(codes_root.encode(), name.encode(), n_threads, F, rx.ctypes.data, rl, out.ctypes.data,
                                              C.byref(secs), C.byref(nl), C.byref(nf)))
    return out, secs.value, nl.value, nf.value


def ber_line(seconds, name, ebn0_db, frames, bit_errors, message_length):
    buf = C.create_string_buffer(512)
    load().ldpc_b200_ber_line(buf, 512, seconds, name.encode(), ebn0_db, frames, bit_errors, message_length)
    return buf.value.decode()
