"""Identity of the shipped kernels for the measurement files: SHA-256 over the sources the RESIDENT kernels are compiled from
and the Makefile (flags).  The library file itself cannot serve: two nvcc builds of the same sources differ in the unique ids
nvcc gives internal symbols, so its hash changes with every rebuild while the SASS does not."""
import hashlib
import os

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
KERNEL_SOURCES = ("resident_engine.cu", "resident_kernel.cuh", "resident_common.cuh", "bp_tree.cuh", "spa_fp32.cuh", "tma.cuh", "qc_tables.inc",
                  "philox.cuh", "common.cuh", "Makefile")


def kernel_source_sha256() -> str:
    h = hashlib.sha256()
    for fn in KERNEL_SOURCES:
        h.update(fn.encode() + b"\0" + open(os.path.join(_CSRC, fn), "rb").read() + b"\0")
    return h.hexdigest()
