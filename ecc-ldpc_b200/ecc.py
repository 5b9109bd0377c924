"""Python view of the C++ host layer (include/ldpc_b200_ecc.h), which mirrors the reference's plug-in glue for the
decode path (src/ECC/Code/LDPC/Utils.hs:26-108): the code-name grammar of mkLDPC_CodeIO

    ldpc/<decoder>/<matrix-name>/<max-rounds>[/<num>/<den>]        decoders: b200-minsum, b200-spa, *-fp64

and the ECC record mkLDPC builds (name, encode, decode, message_length, codeword_length).  Same argument meaning
and error behaviour as the reference: a name that is not ours matches nothing (empty list), a missing matrix file
or a G/H size mismatch raises, punctured positions are zero-filled (Utils.hs:55), the first message_length decoded
bits come back with flag True (Utils.hs:68-72).  All logic lives in libldpc_b200.so; this file only forwards.
"""
from __future__ import annotations

from typing import List, Optional

from . import codes as _codes
from . import lib as _lib

DECODERS = ("b200-minsum", "b200-spa", "b200-minsum-fp64", "b200-spa-fp64", "b200-spa-cm-fp64", "b200-arraylet-fp64",
            "b200-arraylet-min-fp64")
USAGE = ["ldpc/%s/<matrix-name>/<max-rounds>[/codeword/message]" % n for n in DECODERS]


def match(name: str, codes_root: Optional[str] = None, max_batch: int = 1024) -> List[_lib.Ecc]:
    """What a Code's matcher returns for `name`: [] when the name is not in our grammar (Utils.hs:108)."""
    root = codes_root or _codes.builtin_codes_dir()
    try:
        return [_lib.Ecc(root, name, max_batch)]
    except _lib.LdpcError as e:
        if e.code == _lib.ENOMATCH:
            return []
        raise
