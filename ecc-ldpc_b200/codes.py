"""Parity-check / generator matrices of the reference's codes, and readers for its file formats.

Mirrors the reference's matrix layer for the decode path only:
  * loadMatrix search order and formats      -- src/Data/BitMatrix/Loader.hs:58-81
  * .q quasi-cyclic semantics                  -- src/Data/Matrix/QuasiCyclic.hs:19-25, :52-56
  * .alist (reference orientation)             -- src/Data/BitMatrix/Alist.hs:30-46
  * .m 0/1 matrix                              -- src/Data/BitMatrix/Matlab.hs:19-26
The four codes of /root/reference/codes travel with this repo as JSON (tools/import_codes.py), because
the GPU box has no /root/reference.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_CODES_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "codes")

BUILTIN = ("moon.7.13", "jpl.1024.4.5", "jpl.4096.4.5", "1920.1280.3.303")


@dataclass
class Code:
    """H in the reference's edge order (row-major, ascending column) plus optional QC metadata and G."""
    name: str
    M: int
    N: int
    row_ptr: np.ndarray            # int32 [M+1]
    col_idx: np.ndarray            # int32 [E]
    K: int                         # message length = rows of G (Utils.hs:39); N - rank(H) when no G exists
    # quasi-cyclic metadata (None for general codes)
    sz: Optional[int] = None
    brows: Optional[int] = None
    bcols: Optional[int] = None
    shifts: Optional[np.ndarray] = None   # int32 [brows, bcols], -1 = zero block (Arraylet2.hs:325-331)
    # generator parity part, K x (N-K), as CSR (None when the reference has no G for the code)
    g_row_ptr: Optional[np.ndarray] = None
    g_col_idx: Optional[np.ndarray] = None
    g_words: Optional[list] = field(default=None, repr=False)   # QC: python ints [g_brows][g_bcols]
    sha256: dict = field(default_factory=dict)

    @property
    def E(self) -> int:
        return int(self.row_ptr[-1])

    @property
    def is_qc(self) -> bool:
        return self.sz is not None

    def dense(self) -> np.ndarray:
        h = np.zeros((self.M, self.N), dtype=np.uint8)
        for m in range(self.M):
            h[m, self.col_idx[self.row_ptr[m]:self.row_ptr[m + 1]]] = 1
        return h

    def g_dense(self) -> np.ndarray:
        P = self.N - self.K
        g = np.zeros((self.K, P), dtype=np.uint8)
        for k in range(self.K):
            g[k, self.g_col_idx[self.g_row_ptr[k]:self.g_row_ptr[k + 1]]] = 1
        return g

    def c_length(self, num: Optional[int] = None, den: Optional[int] = None) -> int:
        """Transmitted length, Utils.hs:46-50: rate defaults to K/N; c_length = K*den `div` num."""
        if num is None:
            return self.N
        return (self.K * den) // num


def _csr_from_rows(rows: List[List[int]]):
    row_ptr = np.zeros(len(rows) + 1, dtype=np.int32)
    for i, r in enumerate(rows):
        row_ptr[i + 1] = row_ptr[i] + len(r)
    col_idx = np.array([c for r in rows for c in sorted(r)], dtype=np.int32)
    return row_ptr, col_idx


def expand_qc_rows(sz: int, words: List[List[int]]) -> List[List[int]]:
    """QuasiCyclic.hs:19-25: entry (m*sz+i, n*sz+j) is set iff bit ((j-i) mod sz) of words[m][n] is set."""
    rows = []
    for brow in words:
        for i in range(sz):
            r = []
            for n, w in enumerate(brow):
                k = 0
                while w >> k:
                    if (w >> k) & 1:
                        r.append(n * sz + (i + k) % sz)
                    k += 1
            rows.append(sorted(r))
    return rows


def from_qc(name: str, sz: int, shifts, g_words=None, sha256=None) -> Code:
    shifts = np.asarray(shifts, dtype=np.int32)
    brows, bcols = shifts.shape
    words = [[(1 << int(s)) if s >= 0 else 0 for s in row] for row in shifts]
    rows = expand_qc_rows(sz, words)
    row_ptr, col_idx = _csr_from_rows(rows)
    code = Code(name=name, M=brows * sz, N=bcols * sz, row_ptr=row_ptr, col_idx=col_idx, K=0,
                sz=sz, brows=brows, bcols=bcols, shifts=shifts, sha256=sha256 or {})
    if g_words is not None:
        grows = expand_qc_rows(sz, g_words)
        code.K = len(grows)
        code.g_row_ptr, code.g_col_idx = _csr_from_rows(grows)
        code.g_words = g_words
    return code


def from_rows(name: str, M: int, N: int, h_rows, g_rows=None, K: Optional[int] = None, sha256=None) -> Code:
    assert len(h_rows) == M
    row_ptr, col_idx = _csr_from_rows(h_rows)
    code = Code(name=name, M=M, N=N, row_ptr=row_ptr, col_idx=col_idx, K=K or 0, sha256=sha256 or {})
    if g_rows:
        code.K = len(g_rows)
        code.g_row_ptr, code.g_col_idx = _csr_from_rows(g_rows)
    return code


def gf2_rank(h: np.ndarray) -> int:
    a = np.packbits(h.astype(np.uint8), axis=1, bitorder="little")
    a = [int.from_bytes(r.tobytes(), "little") for r in a]
    rank = 0
    for bit in range(h.shape[1]):
        piv = None
        for i in range(rank, len(a)):
            if (a[i] >> bit) & 1:
                piv = i
                break
        if piv is None:
            continue
        a[rank], a[piv] = a[piv], a[rank]
        for i in range(len(a)):
            if i != rank and (a[i] >> bit) & 1:
                a[i] ^= a[rank]
        rank += 1
    return rank


def load_builtin(name: str) -> Code:
    """The codes of /root/reference/codes, from this repo's JSON fixtures."""
    with open(os.path.join(_CODES_DIR, name + ".json")) as f:
        d = json.load(f)
    if d["kind"] == "qc":
        g_words = [[int(x, 16) for x in row] for row in d["g_words_hex"]]
        return from_qc(d["name"], d["sz"], d["h_shifts"], g_words, d.get("source_sha256"))
    code = from_rows(d["name"], d["M"], d["N"], d["h_rows"], d.get("g_rows") or None, sha256=d.get("source_sha256"))
    if code.K == 0:
        # no generator in the reference (1920.1280.3.303): K = N - rank(H); only the all-zero codeword is used
        code.K = {"1920.1280.3.303": 640}.get(name) or (code.N - gf2_rank(code.dense()))
    return code


# ---------------------------------------------------------------------------------------------
# Readers for the reference's own file formats (row f-2 of SURVEY.md section 8)
# ---------------------------------------------------------------------------------------------

def parse_q(text: str):
    """QuasiCyclic.hs:52-56: first token = circulant size, rest = Matlab-style integer matrix."""
    first, rest = text.split(None, 1)
    rows = [[int(t) for t in line.split()] for line in rest.splitlines() if line.strip()]
    return int(first), rows


def parse_alist(text: str):
    """Alist.hs:30-46, reference orientation: zeros dropped, 'nrows ncols', two ignored numbers,
    row weights, column weights, row lists (1-based); column lists are read and discarded."""
    toks = [int(t) for t in text.split()]
    toks = [t for t in toks if t != 0]
    it = iter(toks)
    n = next(it); m = next(it); next(it); next(it)
    rw = [next(it) for _ in range(n)]
    cw = [next(it) for _ in range(m)]
    rows = [sorted(next(it) - 1 for _ in range(c)) for c in rw]
    for c in cw:                      # consumed like the reference does; content ignored
        for _ in range(c):
            next(it)
    return n, m, rows


def parse_m(text: str):
    """BitMatrix/Matlab.hs:19-26: lines of space separated 0/1."""
    rows = []
    for line in text.splitlines():
        t = line.split()
        if t:
            rows.append([j for j, v in enumerate(t) if v == "1"])
    ncols = max(len(line.split()) for line in text.splitlines() if line.split())
    return len(rows), ncols, rows


def _load_matrix(path_noext: str, prefixes):
    """Loader.hs:58-81: suffixes q, alist, m (in that order) x prefixes; first existing file wins."""
    for suffix, kind in (("q", "q"), ("alist", "alist"), ("m", "m")):
        for prefix in prefixes:
            p = os.path.join(prefix, path_noext + "." + suffix)
            if os.path.isfile(p):
                return kind, open(p).read()
    raise FileNotFoundError("can not find any matrix files in %r" % path_noext)


def load_dir(name: str, codes_root: str) -> Code:
    """Equivalent of `loadMatrix (name ++ "/G")`, `loadMatrix (name ++ "/H")` (Utils.hs:36,40)."""
    hk, ht = _load_matrix(os.path.join(name, "H"), [codes_root])
    gk, gt = _load_matrix(os.path.join(name, "G"), [codes_root])
    if hk == "q":
        sz, hw = parse_q(ht)
        if all(bin(v).count("1") <= 1 for r in hw for v in r):
            shifts = [[(v.bit_length() - 1) if v else -1 for v in r] for r in hw]
            code = from_qc(name, sz, shifts)
        else:
            rows = expand_qc_rows(sz, hw)
            code = from_rows(name, len(rows), sz * len(hw[0]), rows)
    elif hk == "alist":
        n, m, rows = parse_alist(ht)
        code = from_rows(name, n, m, rows)
    else:
        n, m, rows = parse_m(ht)
        code = from_rows(name, n, m, rows)
    if gk == "q":
        gsz, gw = parse_q(gt)
        grows = expand_qc_rows(gsz, gw)
        code.g_words = gw
    elif gk == "alist":
        _, _, grows = parse_alist(gt)
    else:
        _, _, grows = parse_m(gt)
    code.K = len(grows)
    code.g_row_ptr, code.g_col_idx = _csr_from_rows(grows)
    p = code.N - code.K
    gcols = (max((max(r) for r in grows if r), default=-1) + 1)
    if gcols > p:
        raise ValueError("bad code size match %r" % ((code.K + gcols, code.N),))   # Utils.hs:43-44
    return code


# ---------------------------------------------------------------------------------------------
# Writers: the builtin fixtures back into the reference's file formats (what `codes/` looks like to the C++ loader)
# ---------------------------------------------------------------------------------------------

def _alist_text(nrows, ncols, rows):
    """Reference orientation (Alist.hs:12-24 `show`): 'nrows ncols', max weights, row weights, column weights,
    row lists, column lists (1-based)."""
    cols = [[] for _ in range(ncols)]
    for r, cs in enumerate(rows):
        for c in cs:
            cols[c].append(r)
    out = ["%d %d" % (nrows, ncols), "%d %d" % (max(len(r) for r in rows), max(len(c) for c in cols)),
           " ".join(str(len(r)) for r in rows), " ".join(str(len(c)) for c in cols)]
    out += [" ".join(str(c + 1) for c in r) for r in rows]
    out += [" ".join(str(r + 1) for r in c) for c in cols]
    return "\n".join(out) + "\n"


def _q_text(sz, words):
    return "%d\n" % sz + "".join(" ".join(str(w) for w in row) + " \n" for row in words)


def _rows_of(row_ptr, col_idx):
    return [list(map(int, col_idx[row_ptr[i]:row_ptr[i + 1]])) for i in range(len(row_ptr) - 1)]


def write_reference_dir(root, names=("moon.7.13", "jpl.1024.4.5", "jpl.4096.4.5")):
    """Writes <root>/<name>/{H,G}.{q|alist} for the builtin codes that have a generator. Returns root."""
    for name in names:
        code = load_builtin(name)
        d = os.path.join(root, name)
        os.makedirs(d, exist_ok=True)
        if code.is_qc:
            hw = [[(1 << int(s)) if s >= 0 else 0 for s in row] for row in code.shifts]
            open(os.path.join(d, "H.q"), "w").write(_q_text(code.sz, hw))
            open(os.path.join(d, "G.q"), "w").write(_q_text(code.sz, code.g_words))
        else:
            open(os.path.join(d, "H.alist"), "w").write(_alist_text(code.M, code.N, _rows_of(code.row_ptr, code.col_idx)))
            open(os.path.join(d, "G.alist"), "w").write(_alist_text(code.K, code.N - code.K, _rows_of(code.g_row_ptr, code.g_col_idx)))
    return root


def builtin_codes_dir():
    """A `codes/` directory in the reference's formats, generated once from the JSON fixtures."""
    root = os.path.join(_CODES_DIR, "_ref_format")
    if not os.path.isfile(os.path.join(root, "jpl.4096.4.5", "G.q")):
        # several ranks may arrive here at once (torchrun): build in a private directory, publish with one rename
        import shutil
        tmp = "%s.tmp.%d" % (root, os.getpid())
        write_reference_dir(tmp)
        try:
            os.rename(tmp, root)
        except OSError:                      # another process published first
            shutil.rmtree(tmp, ignore_errors=True)
    return root
