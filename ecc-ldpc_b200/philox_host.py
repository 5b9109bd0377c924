"""Host (numpy) statement of the device frame generator of csrc/philox.cuh: Philox4x32-10 keyed by the seed, counter =
(group of 4 positions, 0, frame_lo, frame_hi), Box-Muller in float32, BPSK + AWGN, LLR = (2 / sigma^2) * y.

The reference gets its frames from ecc-manifold's tester (external); this module lets a HOST-side consumer
(the CPU arm of bench.py, the tests) draw exactly the frames `ldpc_b200_generate_llr` / `ldpc_b200_simulate` draw on the
device for the same (seed, global frame index): the Philox words are identical bit for bit; the float32 logarithm /
sine / cosine of numpy and of CUDA differ by at most a couple of units in the last place, so the LLRs agree to ~1e-6
relative (tests/test_gpu_parity.py::test_host_generator_matches_device).
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): counters as uint64 arrays holding 32-bit values, scalar key."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def awgn_normals(seed: int, first_frame: int, n_frames: int, n: int) -> np.ndarray:
    """[n_frames, n] float32 N(0,1) samples: positions 4g .. 4g+3 of frame f come from the Philox block (g, 0, f_lo, f_hi)."""
    groups = (n + 3) // 4
    f = (np.arange(n_frames, dtype=np.uint64) + np.uint64(first_frame))[:, None]
    g = np.arange(groups, dtype=np.uint64)[None, :]
    r = philox4x32_10(g + np.zeros_like(f), np.zeros((n_frames, groups), dtype=np.uint64), (f & _MASK) + np.zeros_like(g),
                      (f >> np.uint64(32)) + np.zeros_like(g), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((n_frames, groups, 4), dtype=np.float32)
    two_m24 = np.float32(5.9604644775390625e-08)
    for p in range(2):
        u1 = ((r[2 * p] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * two_m24
        u2 = ((r[2 * p + 1] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * two_m24
        rad = np.sqrt(np.float32(-2.0) * np.log(u1))
        ang = (np.float32(2.0) * u2).astype(np.float64) * np.pi            # sincospif(2 u2): evaluated in float64, rounded to float32
        out[:, :, 2 * p] = rad * np.cos(ang).astype(np.float32)
        out[:, :, 2 * p + 1] = rad * np.sin(ang).astype(np.float32)
    return out.reshape(n_frames, groups * 4)[:, :n]


def channel_params(ebn0_db: float, k_info: int, n_tx: int):
    """sigma and 2 / sigma^2 as the float32 values api.cu:channel_params hands the kernels."""
    R = k_info / n_tx
    s2 = 1.0 / (2.0 * R * 10.0 ** (ebn0_db / 10.0))
    return np.float32(np.sqrt(s2)), np.float32(2.0 / s2)


def awgn_llrs(N: int, ebn0_db: float, k_info: int, n_tx: int, seed: int, first_frame: int, n_frames: int) -> np.ndarray:
    """[n_frames, N] float32 LLRs of the all-zero codeword: y = -1 + sigma n (fused multiply-add), llr = scale * y for
    the first n_tx positions, 0 for the punctured rest (Utils.hs:55)."""
    sigma, scale = channel_params(ebn0_db, k_info, n_tx)
    nz = awgn_normals(seed, first_frame, n_frames, N)
    y = (sigma.astype(np.float64) * nz.astype(np.float64) - 1.0).astype(np.float32)      # one rounding, like __fmaf_rn
    llr = (scale * y).astype(np.float32)
    llr[:, n_tx:] = 0.0
    return llr
