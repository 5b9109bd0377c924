"""CPU suite: the C-ABI library loads, exports every symbol include/ldpc_b200.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header="ldpc_b200.h"):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ldpc_b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from ecc_ldpc_b200 import lib
    L = lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(lib.EXPORTS) == names
    assert L.ldpc_b200_abi_version() == 1
    ecc_names = _declared("ldpc_b200_ecc.h")
    assert len(ecc_names) >= 18
    for n in ecc_names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(lib.ECC_EXPORTS) == ecc_names


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ecc_ldpc_b200 import lib
    L = lib.load()
    assert L.ldpc_b200_device_count() == 0
    with pytest.raises(lib.LdpcError) as ei:
        lib.init(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
    from ecc_ldpc_b200 import codes
    with pytest.raises(lib.LdpcError):
        lib.DeviceCode(codes.load_builtin("moon.7.13"))


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under the package or the shared library may reference it."""
    pkg = os.path.join(ROOT, "ecc-ldpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, fn)


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (host instance of the same inline function)."""
    from ecc_ldpc_b200 import lib
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in lib.philox4x32(ctr, key)) == want
