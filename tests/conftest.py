import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def builtin_codes():
    from ecc_ldpc_b200 import codes
    return {n: codes.load_builtin(n) for n in codes.BUILTIN}


@pytest.fixture(scope="session")
def oracles(builtin_codes):
    from oracle import oracle
    return {n: oracle.Oracle(c) for n, c in builtin_codes.items()}


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def awgn_llrs(code, ebn0_db, n_frames, seed, n_tx=None, codewords=None):
    """BPSK (bit b -> 2b-1) + AWGN LLRs the way the tester makes them: float64, zeros at punctured."""
    rng = np.random.default_rng(seed)
    n_tx = code.N if n_tx is None else n_tx
    R = code.K / n_tx
    sigma = np.sqrt(1.0 / (2 * R * 10 ** (ebn0_db / 10)))
    cw = np.zeros((n_frames, code.N)) if codewords is None else codewords.astype(np.float64)
    y = (2.0 * cw - 1.0) + sigma * rng.standard_normal((n_frames, code.N))
    llr = 2.0 * y / sigma ** 2
    llr[:, n_tx:] = 0.0
    return llr


@pytest.fixture(scope="session")
def gpu_lib():
    from ecc_ldpc_b200 import lib
    lib.load()
    lib.init(0)
    return lib
