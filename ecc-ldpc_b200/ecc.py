"""Host-side mirror of the reference's plug-in glue for the decode path (src/ECC/Code/LDPC/Utils.hs:26-108).

`mk_ldpc` builds what `mkLDPC` builds -- an ECC record (name, encode, decode, message_length,
codeword_length) -- with the `decoder` slot filled by libldpc_b200.so through its C ABI.  `match` implements the
code-name grammar of `mkLDPC_CodeIO` (Utils.hs:100-108):

    ldpc/<decoder>/<matrix-name>/<max-rounds>[/<num>/<den>]

with the decoder names this library registers: b200-minsum, b200-spa (fp32 production) and b200-minsum-fp64,
b200-spa-fp64 (verification).  Same argument meaning and error behaviour as the reference: unknown names match
nothing (empty list), a G/H size mismatch raises, punctured positions are zero-filled (Utils.hs:55), the first
message_length decoded bits are returned with flag True (Utils.hs:68-72).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import codes as _codes
from . import lib as _lib

DECODERS = {
    "b200-minsum": ("min", "fp32"),
    "b200-spa": ("spa", "fp32"),
    "b200-minsum-fp64": ("min", "fp64"),
    "b200-spa-fp64": ("spa", "fp64"),
}

USAGE = ["ldpc/%s/<matrix-name>/<max-rounds>[/codeword/message]" % n for n in DECODERS]


@dataclass
class ECC:
    """The fields of `ECC IO` as constructed at Utils.hs:59-75."""
    name: str
    encode: Callable[[np.ndarray], np.ndarray]
    decode: Callable[[np.ndarray], Tuple[np.ndarray, bool]]
    message_length: int
    codeword_length: int
    decode_batch: Optional[Callable] = None
    close: Optional[Callable[[], None]] = None


def mk_ldpc(prefix: str, code_name: str, max_iters: int, rate: Optional[Tuple[int, int]] = None,
            codes_root: Optional[str] = None, max_batch: int = 1024, engine: int = _lib.ENGINE_AUTO) -> ECC:
    """Utils.hs:35-75. `rate` = (numerator, denominator) of message/codeword, e.g. (4, 5)."""
    if prefix not in DECODERS:
        raise ValueError("unknown decoder %r" % prefix)
    algo, prec = DECODERS[prefix]
    code = _codes.load_dir(code_name, codes_root) if codes_root else _codes.load_builtin(code_name)
    if code.g_row_ptr is None:
        raise FileNotFoundError("can not find any matrix files in %r" % (code_name + "/G"))   # Loader.hs:73
    m_length = code.K
    unpunc = code.N
    num, den = rate if rate is not None else (m_length, unpunc)
    if rate is None:                                  # m_length % unpunc_c_length, reduced like Data.Ratio does
        from math import gcd
        g = gcd(num, den)
        num, den = num // g, den // g
    c_length = (m_length * den) // num                # Utils.hs:50
    g_dense_t = code.g_dense().T.astype(np.uint8)     # (N-K) x K
    _lib.init(-1)
    dcode = _lib.DeviceCode(code)
    dec = _lib.Decoder(dcode, algo, prec, max_batch=max_batch, engine=engine)

    def encode(inp):
        inp = np.asarray(inp, dtype=np.uint8)
        parity = (g_dense_t.astype(np.int32) @ inp.astype(np.int32)) & 1            # Orig.hs:25-26: v * G over GF(2)
        return np.concatenate([inp, parity[: c_length - m_length].astype(np.uint8)])  # Utils.hs:61

    def unpuncture(rx):
        rx = np.asarray(rx, dtype=np.float64)
        if rx.ndim == 1:
            rx = rx[None, :]
        full = np.zeros((rx.shape[0], unpunc), dtype=np.float64)
        take = min(c_length, rx.shape[1])
        full[:, :take] = rx[:, :take]                 # U.take c_length xs ++ replicate (N - c_length) 0
        return full

    def decode_batch(rx):
        bits, iters, conv = dec.decode(unpuncture(rx), max_iters)
        return bits[:, :m_length], iters, conv

    def decode(rx):
        bits, _, _ = decode_batch(rx)
        return bits[0], True                          # every decoder returns Just (Utils.hs:72)

    def close():
        dec.close()
        dcode.close()

    return ECC(name="ldpc/%s/%s/%d/%d/%d" % (prefix, code_name, max_iters, num, den),   # Utils.hs:60
               encode=encode, decode=decode, message_length=m_length, codeword_length=c_length,
               decode_batch=decode_batch, close=close)


def parse_name(name: str):
    """The matcher of mkLDPC_CodeIO (Utils.hs:100-108): returns (decoder, matrix, max_rounds, rate|None) or None."""
    xs = name.split("/")
    if len(xs) in (4, 6) and xs[0] == "ldpc" and xs[1] in DECODERS and xs[3].isdigit() and xs[3] != "":
        if len(xs) == 4:
            return xs[1], xs[2], int(xs[3]), None
        if xs[4].isdigit() and xs[5].isdigit() and xs[4] and xs[5]:
            return xs[1], xs[2], int(xs[3]), (int(xs[4]), int(xs[5]))
    return None


def match(name: str, **kw) -> List[ECC]:
    p = parse_name(name)
    if p is None:
        return []                                     # `_ -> return []`
    return [mk_ldpc(p[0], p[1], p[2], p[3], **kw)]
