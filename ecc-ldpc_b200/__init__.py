"""ldpc-b200: B200-native flooding belief-propagation LDPC decoder behind ku-fpg/ecc-ldpc's ECC record.

Python here is only the ctypes view of libldpc_b200.so (the C-ABI in include/ldpc_b200.h) plus the code
fixtures; the product is the shared library.  Submodules:
  codes  -- H/G of the reference's codes and readers for its file formats (no GPU needed)
  lib    -- ctypes binding; raises if the CUDA library is missing (there is no CPU fallback)
  ecc    -- mirror of mkLDPC / the ECC record (Utils.hs:26-108) over the C-ABI
"""
__version__ = "0.1.0"
