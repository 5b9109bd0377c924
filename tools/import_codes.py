#!/usr/bin/env python3
"""Import the reference's parity-check / generator matrices into this repo's own fixture format.

Run in the dev container only (needs /root/reference).  The GPU box has no /root/reference, so the
matrices travel as JSON under ecc-ldpc_b200/codes/.  The JSON is a *derived* description (circulant shift
tables, CSR row lists, generator circulants as hex words), not a copy of the reference files; the
SHA-256 of every source file is recorded so tests can prove provenance when the reference is present.

File formats understood here are the ones the reference reads:
  .q     src/Data/Matrix/QuasiCyclic.hs:48-56   first token = circulant size, then rows of decimal integers
  .alist src/Data/BitMatrix/Alist.hs:30-46      "nrows ncols", 2 ignored, row weights, col weights, row lists
  MacKay alist (codes/1920.1280.3.303)          "N M" header, column lists first -> stored as true H (M x N)
"""
import hashlib, json, os, sys

REF = os.environ.get("LDPC_REFERENCE_CODES", "/root/reference/codes")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ecc-ldpc_b200", "codes")


def sha(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest()


def read_q(p):
    txt = open(p).read()
    first, rest = txt.split(None, 1)
    rows = [[int(t) for t in line.split()] for line in rest.splitlines() if line.strip()]
    return int(first), rows


def read_alist_ref(p):
    """Reference orientation (Alist.hs:30-46): zeros are dropped from the token stream first."""
    toks = [int(t) for t in open(p).read().split()]
    toks = [t for t in toks if t != 0]
    it = iter(toks)
    n = next(it); m = next(it); next(it); next(it)
    rw = [next(it) for _ in range(n)]
    cw = [next(it) for _ in range(m)]
    rows = [[next(it) - 1 for _ in range(c)] for c in rw]
    return n, m, rows


def read_alist_mackay(p):
    """MacKay orientation: 'N M', max col weight, max row weight, N col weights, M row weights,
    N column lists (zero padded), M row lists (zero padded).  Returns true H as M rows of column indices."""
    toks = [int(t) for t in open(p).read().split()]
    it = iter(toks)
    N = next(it); M = next(it); maxc = next(it); maxr = next(it)
    cw = [next(it) for _ in range(N)]
    rw = [next(it) for _ in range(M)]
    cols = [[x - 1 for x in (next(it) for _ in range(maxc)) if x != 0] for _ in range(N)]
    rows = [[x - 1 for x in (next(it) for _ in range(maxr)) if x != 0] for _ in range(M)]
    chk = [[] for _ in range(M)]
    for j, cl in enumerate(cols):
        assert len(cl) == cw[j]
        for r in cl:
            chk[r].append(j)
    assert [sorted(r) for r in rows] == [sorted(r) for r in chk], "row/column lists disagree"
    assert [len(r) for r in rows] == rw
    return M, N, [sorted(r) for r in rows]


def qc_code(name):
    d = os.path.join(REF, name)
    sz, h = read_q(os.path.join(d, "H.q"))
    gsz, g = read_q(os.path.join(d, "G.q"))
    assert gsz == sz
    shifts = []
    for row in h:
        srow = []
        for v in row:
            if v == 0:
                srow.append(-1)
            else:
                assert bin(v).count("1") == 1, "multi-weight circulant"
                srow.append(v.bit_length() - 1)
        shifts.append(srow)
    out = {
        "name": name, "kind": "qc", "sz": sz, "brows": len(h), "bcols": len(h[0]),
        "h_shifts": shifts,
        "g_brows": len(g), "g_bcols": len(g[0]),
        "g_words_hex": [["%x" % v for v in row] for row in g],
        "source_sha256": {"H.q": sha(os.path.join(d, "H.q")), "G.q": sha(os.path.join(d, "G.q"))},
        "note": "h_shifts[r][c] = k means ones at (r*sz+i, c*sz+(i+k)%sz); -1 = zero block. "
                "g_words_hex[r][c] bit k set means ones at (r*sz+i, c*sz+(i+k)%sz) of the K x (N-K) parity part of G.",
    }
    return out


def moon():
    d = os.path.join(REF, "moon.7.13")
    n, m, hrows = read_alist_ref(os.path.join(d, "H.alist"))
    gn, gm, grows = read_alist_ref(os.path.join(d, "G.alist"))
    return {
        "name": "moon.7.13", "kind": "csr", "M": n, "N": m, "h_rows": hrows,
        "g_rows_n": gn, "g_cols_n": gm, "g_rows": grows,
        "source_sha256": {"H.alist": sha(os.path.join(d, "H.alist")), "G.alist": sha(os.path.join(d, "G.alist"))},
        "note": "h_rows[m] = ascending column indices (0-based) of check m; g_rows[k] = columns of the K x (N-K) parity part.",
    }


def mackay():
    p = os.path.join(REF, "1920.1280.3.303")
    M, N, rows = read_alist_mackay(p)
    return {
        "name": "1920.1280.3.303", "kind": "csr", "M": M, "N": N, "h_rows": rows,
        "g_rows_n": 0, "g_cols_n": 0, "g_rows": [],
        "source_sha256": {"1920.1280.3.303": sha(p)},
        "note": "MacKay orientation file (header 'N M'); stored here as the true H (M=1280 checks x N=1920 variables). "
                "rank(H)=1280 so K=640, rate 1/3 (SURVEY.md Appendix B/D). No generator exists in the reference.",
    }


def main():
    os.makedirs(OUT, exist_ok=True)
    for c in [moon(), qc_code("jpl.1024.4.5"), qc_code("jpl.4096.4.5"), mackay()]:
        p = os.path.join(OUT, c["name"] + ".json")
        with open(p, "w") as f:
            json.dump(c, f, separators=(",", ":"))
            f.write("\n")
        print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    sys.exit(main())
