#!/usr/bin/env python3
"""Generates ecc-ldpc_b200/csrc/qc_tables.inc: compile-time circulant tables of the quasi-cyclic codes the
RESIDENT engine is specialised for (the reference's two QC codes, jpl.1024.4.5 and jpl.4096.4.5).

Input is this repo's JSON fixtures (ecc-ldpc_b200/codes/*.json), so it runs anywhere.  For each code the
12 block rows are split into NP parts (one per thread of a check position; NP = 2 and 4 are emitted) such that
the check-node work is equal and the per-stage variable-node work is as balanced as possible.
"""
import itertools
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODES = os.path.join(ROOT, "ecc-ldpc_b200", "codes")
OUT = os.path.join(ROOT, "ecc-ldpc_b200", "csrc", "qc_tables.inc")


def arr(vals):
    return "{" + ", ".join(str(v) for v in vals) + "}"


def arr2(rows):
    return "{" + ",\n        ".join(arr(r) for r in rows) + "}"


def split_rows(shifts, NP):
    """Partition the block rows into NP parts with equal edge counts, minimising sum over VN stages of the
    largest per-part stage work (every stage ends in a barrier, so the slowest part sets its length)."""
    BR, BC = len(shifts), len(shifts[0])
    deg = [sum(1 for v in row if v >= 0) for row in shifts]
    cols = [sorted((r for r in range(BR) if shifts[r][c] >= 0), reverse=True) for c in range(BC)]
    maxcd = max(len(c) for c in cols)
    total = sum(deg)
    target, per = total // NP, BR // NP
    assert total % NP == 0 and BR % NP == 0
    best = [None]

    def cost(parts):
        owner = {}
        for pi, rows in enumerate(parts):
            for r in rows:
                owner[r] = pi
        tot = 0
        for q in range(maxcd):
            cnt = [0] * NP
            for c in cols:
                if len(c) > q:
                    cnt[owner[c[q]]] += 1
            tot += max(cnt)
        return tot

    def rec(remaining, parts):
        if not remaining:
            c = cost(parts)
            if best[0] is None or c < best[0][0]:
                best[0] = (c, [list(p) for p in parts])
            return
        first = remaining[0]
        rest = remaining[1:]
        for comb in itertools.combinations(rest, per - 1):
            rows = (first,) + comb
            if sum(deg[r] for r in rows) != target:
                continue
            rec([r for r in rest if r not in comb], parts + [rows])

    rec(list(range(BR)), [])
    assert best[0] is not None, "no balanced split"
    return best[0][1], best[0][0]


def schedule_stages(cols_desc, owner, NP, T, seed=1):
    """Variable-node schedule: column c must apply its chain steps q = 0..deg-1 (block rows in descending order) in
    order, at most one step per stage; inside a stage every address is then touched by one thread only.  Choose
    the stage of every step so that the per-stage maximum over parts (the stage ends in a barrier) summed over the T
    stages is minimal.  Local search over 'which stages does column c use', deterministic."""
    import random
    rng = random.Random(seed)
    BC = len(cols_desc)
    choice = [list(range(len(c))) for c in cols_desc]            # start: step q in stage q

    def loads(ch):
        L = [[0] * NP for _ in range(T)]
        for c, st in enumerate(ch):
            for q, t in enumerate(st):
                L[t][owner[cols_desc[c][q]]] += 1
        return L

    def cost(ch):
        L = loads(ch)
        return sum(max(row) for row in L) * 1000 + sum(max(row) - min(row) for row in L)

    best = cost(choice)
    for it in range(60000):
        c = rng.randrange(BC)
        d = len(cols_desc[c])
        if d == T:
            continue
        new = sorted(rng.sample(range(T), d))
        old = choice[c]
        choice[c] = new
        v = cost(choice)
        if v <= best:
            best = v
        else:
            choice[c] = old
    table = [[-1] * T for _ in range(BC)]
    for c, st in enumerate(choice):
        for q, t in enumerate(st):
            table[c][t] = q
    return table, best // 1000


def emit(name, ident, out, NP=2):
    d = json.load(open(os.path.join(CODES, name + ".json")))
    sh = d["h_shifts"]
    sz, BR, BC = d["sz"], d["brows"], d["bcols"]
    deg = [sum(1 for v in row if v >= 0) for row in sh]
    maxd = max(deg)
    row_col = [[c for c in range(BC) if sh[r][c] >= 0] for r in range(BR)]
    row_shift = [[sh[r][c] for c in row_col[r]] for r in range(BR)]
    pad = lambda l, n: l + [-1] * (n - len(l))
    cols_desc = [sorted((r for r in range(BR) if sh[r][c] >= 0), reverse=True) for c in range(BC)]
    maxcd = max(len(c) for c in cols_desc)
    col_k = [[row_col[r].index(c) for r in cols_desc[c]] for c in range(BC)]
    parts, cost = split_rows(sh, NP)
    half_of_row = [next(pi for pi, rows in enumerate(parts) if r in rows) for r in range(BR)]
    col_stage, cost2 = schedule_stages(cols_desc, half_of_row, NP, maxcd)
    out.append("// %s: sz=%d, %dx%d blocks, %d circulants; %d parts %s; VN stage cost %d in chain order, %d scheduled (ideal %d)" %
               (name, sz, BR, BC, sum(deg), NP, parts, cost, cost2, sum(deg) // NP))
    out.append("struct %s {" % ident)
    out.append("    static constexpr const char *name = \"%s\";" % name)
    out.append("    static constexpr int SZ = %d, BR = %d, BC = %d, NEDGE = %d, MAXD = %d, MAXCD = %d, NP = %d, HR = %d;" %
               (sz, BR, BC, sum(deg), maxd, maxcd, NP, BR // NP))
    out.append("    static constexpr int row_deg[BR] = %s;" % arr(deg))
    out.append("    static constexpr int row_col[BR][MAXD] = %s;" % arr2([pad(r, maxd) for r in row_col]))
    out.append("    static constexpr int row_shift[BR][MAXD] = %s;" % arr2([pad(r, maxd) for r in row_shift]))
    out.append("    static constexpr int half_of_row[BR] = %s;" % arr(half_of_row))
    out.append("    static constexpr int half_rows[NP][HR] = %s;" % arr2(parts))
    out.append("    static constexpr int col_deg[BC] = %s;" % arr([len(c) for c in cols_desc]))
    out.append("    static constexpr int col_row_desc[BC][MAXCD] = %s;" % arr2([pad(c, maxcd) for c in cols_desc]))
    out.append("    static constexpr int col_k_desc[BC][MAXCD] = %s;" % arr2([pad(c, maxcd) for c in col_k]))
    out.append("    // col_stage[c][t] = chain step (index into col_row_desc[c]) that column c applies in VN stage t, -1 = idle")
    out.append("    static constexpr int col_stage[BC][MAXCD] = %s;" % arr2(col_stage))
    out.append("};")
    out.append("")


def main():
    out = ["// GENERATED by tools/gen_qc_tables.py from ecc-ldpc_b200/codes/*.json -- do not edit.",
           "// Circulant (r, k): block column row_col[r][k], shift s = row_shift[r][k]:",
           "// check r*SZ+i is connected to variable row_col[r][k]*SZ + (i+s)%SZ   (QuasiCyclic.hs:19-25).",
           ""]
    emit("jpl.1024.4.5", "QcJpl1024P2", out, 2)
    emit("jpl.4096.4.5", "QcJpl4096P2", out, 2)
    emit("jpl.1024.4.5", "QcJpl1024P4", out, 4)
    emit("jpl.4096.4.5", "QcJpl4096P4", out, 4)
    open(OUT, "w").write("\n".join(out))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
