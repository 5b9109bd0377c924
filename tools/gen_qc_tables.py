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
    """Partition the block rows into NP parts with equal edge counts: first maximise the number of edges the column
    executors can fold in registers (see cost), then minimise the sum over the original VN stages of the largest per-part
    stage work (every stage ends in a barrier, so the slowest part sets its length)."""
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
        # edges of one variable that fall to the SAME thread (same part, same shift) are folded in registers by the column's
        # executor: each saves a slot store, a slot load and a posterior load per iteration (plan_short_chains) -- worth far
        # more than stage balance, so it leads the objective
        merged = 0
        for c in range(BC):
            groups = {}
            for r in cols[c]:
                key = (owner[r], shifts[r][c])
                groups[key] = groups.get(key, 0) + 1
            merged += max(groups.values()) - 1
        return tot - 1000 * merged

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


def plan_short_chains(cols_desc, owner, NP, S, cn_single, max_slots, seed=1, exec_rule="any", shifts=None, long_chain=False, balance_cn=False, iters=200000):
    """Variable-node plan with S stages (S < the maximum column degree).  Column c of degree d applies the FIRST
    l <= min(d-1, S-1) links of its chain (innermost first) one per stage as before; of the remaining steps one is the
    column's EXECUTOR and the others are SINGLETONS: a singleton's thread only stores its raw message into a slot of its
    own (no load), in a stage before the executor's -- or, with cn_single, already in the check phase (stage -1); the
    executor loads the chain value (or negL) and the singleton slots and folds them together with its own message IN CHAIN
    ORDER, so the arithmetic -- operands, order, roundings -- is exactly that of the one-link-per-stage schedule
    (Orig.hs:95-98); only who performs it changes.  Shared-memory operations are the same under every plan (d loads + d
    stores per column); barriers per iteration drop from 1 + max degree to 1 + S.  Local search over (l, executor,
    stages) per column, minimising the sum over stages of the busiest part, with at most max_slots singleton slots.
    Returns per-step tables stage / kind (0 chain, 1 singleton, 2 executor) / slot, the number of slots and the loads."""
    import random
    rng = random.Random(seed)
    BC = len(cols_desc)
    degs = [len(c) for c in cols_desc]
    lo_single = -1 if cn_single else 0
    # Same-thread groups: steps of a column whose edges belong to one thread (same part, same shift).  If a column has
    # such a group of two or more, that thread is the executor and ALL the group's messages are its own registers.
    exec_group = []
    for c, rows in enumerate(cols_desc):
        groups = {}
        for q, r in enumerate(rows):
            groups.setdefault((owner[r], shifts[r][c]), []).append(q)
        best_g = max(groups.values(), key=lambda g: (len(g), shifts[rows[g[0]]][c] == 0, max(g)))
        exec_group.append(best_g if len(best_g) > 1 else None)

    def random_assign(d, c):
        G = exec_group[c]
        first = min(G) if G else d - 1                         # chain links come before the executor's first own step
        lmax = min(first, S - 1)
        l = lmax if long_chain else rng.randint(0, lmax)
        st = sorted(rng.sample(range(S), l + 1))               # chain stages..., executor stage
        tf = st[-1]
        if G:
            ex = list(G)
        elif exec_rule == "top":                               # the column's first block row, as in the original schedule
            ex = [d - 1]
        elif exec_rule == "zero":                              # an identity circulant if there is one behind the chain (its thread
            z = [q for q in range(l, d) if shifts[cols_desc[c][q]][c] == 0]        # needs no address arithmetic)
            ex = [rng.choice(z) if z else rng.randint(l, d - 1)]
        elif exec_rule == "nonzero":
            z = [q for q in range(l, d) if shifts[cols_desc[c][q]][c] != 0]
            ex = [rng.choice(z) if z else rng.randint(l, d - 1)]
        else:
            ex = [rng.randint(l, d - 1)]
        ns = d - l - len(ex)
        if ns > 0 and tf - 1 < lo_single:
            return None
        stage = []
        for q in range(d):
            if q < l:
                stage.append(st[q])
            elif q in ex:
                stage.append(tf)
            else:
                stage.append(-1 if cn_single else rng.randint(lo_single, tf - 1))
        return (l, tuple(ex), stage)

    def initial(d, c):
        # the plan with the fewest slots: longest chain
        for _ in range(100000):
            a = random_assign(d, c)
            if a is not None and a[0] == min(min(exec_group[c]) if exec_group[c] else d - 1, S - 1):
                return a
        raise SystemExit("no feasible plan: S=%d cn_single=%s degree %d" % (S, cn_single, d))

    choice = [initial(d, c) for c, d in enumerate(degs)]

    def slots(ch):
        return sum(degs[c] - a[0] - len(a[1]) for c, a in enumerate(ch))

    def contrib(c, a, L, sign):
        l, ex, st = a
        d = degs[c]
        for q, t in enumerate(st):
            if q < l:
                w = 2
            elif q in ex:
                w = 2 + (d - l - len(ex)) if q == ex[-1] else 0    # the executor: chain value + slots in, posterior out
            else:
                w = 1
            L[t if t >= 0 else S][owner[cols_desc[c][q]]] += sign * w
            if balance_cn and q not in ex:
                # the posterior load of a non-executor edge in the check phase, and its address rotation (two ALU-pipe
                # instructions in an ALU-bound phase) when the circulant is not an identity
                L[S][owner[cols_desc[c][q]]] += sign * (1 + (2 if shifts[cols_desc[c][q]][c] != 0 else 0))

    def loads(ch):
        L = [[0] * NP for _ in range(S + 1)]                   # index S: the check phase (stage -1)
        for c, a in enumerate(ch):
            contrib(c, a, L, +1)
        return L

    def cost_of(L, nslots):
        over = max(0, nslots - max_slots)
        # (a singleton's store overlaps the trees of the check phase, a chain link is two operations in a stage: at equal
        # stage cost prefer the plan with more singletons)
        if balance_cn:        # the check phase ends in a barrier too (the vote): its busiest part counts like a stage's
            return (over * 100000 - nslots * 400 + sum(max(row) for row in L) * 1000 + sum(max(row) - min(row) for row in L) * 10)
        return (over * 100000 - nslots * 400 + sum(max(row) for row in L[:S]) * 1000 + sum(max(row) - min(row) for row in L[:S]) * 10 +
                (max(L[S]) - min(L[S])) * 10 + (max(max(row) for row in L[:S]) - min(max(row) for row in L[:S])))

    L = loads(choice)
    ns = slots(choice)
    best = cost_of(L, ns)
    for it in range(iters):
        c = rng.randrange(BC)
        new = random_assign(degs[c], c)
        if new is None:
            continue
        old = choice[c]
        contrib(c, old, L, -1)
        contrib(c, new, L, +1)
        ns2 = ns - (degs[c] - old[0] - len(old[1])) + (degs[c] - new[0] - len(new[1]))
        v = cost_of(L, ns2)
        if v <= best:
            best = v
            ns = ns2
            choice[c] = new
        else:
            contrib(c, new, L, -1)
            contrib(c, old, L, +1)
    assert slots(choice) <= max_slots, "slot budget not met"
    maxcd = max(degs)
    stage = [[-9] * maxcd for _ in range(BC)]
    kind = [[-1] * maxcd for _ in range(BC)]
    slot = [[-1] * maxcd for _ in range(BC)]
    nslot = 0
    for c, (l, ex, st) in enumerate(choice):
        for q, t in enumerate(st):
            stage[c][q] = t
            kind[c][q] = 0 if q < l else (2 if q in ex else 1)
            if kind[c][q] == 1:
                slot[c][q] = nslot
                nslot += 1
    validate_plan(cols_desc, owner, shifts, S, cn_single, stage, kind, slot, nslot)
    return stage, kind, slot, nslot, loads(choice)


def validate_plan(cols_desc, owner, shifts, S, cn_single, stage, kind, slot, nslot):
    """Structural proof obligations of a plan (what makes it equal to the one-link-per-stage schedule bit for bit):
    chain links are a prefix of the chain in strictly increasing stages; the executor steps of a column all belong to ONE
    thread (same part, same shift), lie behind the chain and run in a later stage than every chain link; every other step
    is a singleton with a slot of its own, stored before the executor's stage (in the check phase, stage -1, only if the
    plan allows it)."""
    seen = set()
    for c, rows in enumerate(cols_desc):
        d = len(rows)
        ks = [kind[c][q] for q in range(d)]
        l = 0
        while l < d and ks[l] == 0:
            l += 1
        assert all(k in (1, 2) for k in ks[l:]) and ks.count(2) >= 1, (c, ks)
        ex = [q for q in range(d) if ks[q] == 2]
        assert len({(owner[rows[q]], shifts[rows[q]][c]) for q in ex}) == 1, (c, ex)
        assert len({stage[c][q] for q in ex}) == 1
        chain_st = [stage[c][q] for q in range(l)]
        assert all(0 <= t < S for t in chain_st) and chain_st == sorted(set(chain_st)), (c, chain_st)
        tf = stage[c][ex[0]]
        assert 0 <= tf < S and all(t < tf for t in chain_st), (c, tf, chain_st)
        for q in range(d):
            if ks[q] == 1:
                assert (-1 if cn_single else 0) <= stage[c][q] < tf, (c, q, stage[c][q], tf)
                assert 0 <= slot[c][q] < nslot and slot[c][q] not in seen
                seen.add(slot[c][q])
            else:
                assert slot[c][q] == -1
        assert all(kind[c][q] == -1 for q in range(d, len(kind[c])))
    assert len(seen) == nslot


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
    merged = (-cost + 999) // 1000
    out.append("// %s: sz=%d, %dx%d blocks, %d circulants; %d parts %s: %d edges share their variable's executor thread; VN stage cost %d in chain order, %d scheduled (ideal %d)" %
               (name, sz, BR, BC, sum(deg), NP, parts, merged, cost + 1000 * merged, cost2, sum(deg) // NP))
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
    # short-chain plans (see plan_short_chains): VN<S> with singleton stores in the variable stages, VNC<S> with the
    # singleton stores in the check phase
    for S, cn_single, max_slots, rule, long_chain, seed, tag in PLANS:
        stage, kind, slot, nslot, L = plan_short_chains(cols_desc, half_of_row, NP, S, cn_single, max_slots, seed=seed, exec_rule=rule.split("+")[0], shifts=sh,
                                                        long_chain=long_chain, balance_cn=rule.endswith("+cn"))
        out.append("    // %s: %d stages%s, %d singleton slots; shared-memory operations per stage and part %s%s" %
                   (tag, S, ", singletons stored in the check phase" if cn_single else "", nslot, L[:S],
                    ", check phase %s" % L[S] if cn_single else ""))
        out.append("    struct %s {" % tag)
        out.append("        static constexpr int S = %d, NSLOT = %d;" % (S, nslot))
        out.append("        static constexpr int stage[BC][MAXCD] = %s;" % arr2(stage).replace("\n        ", "\n            "))
        out.append("        static constexpr int kind[BC][MAXCD] = %s;" % arr2(kind).replace("\n        ", "\n            "))
        out.append("        static constexpr int slot[BC][MAXCD] = %s;" % arr2(slot).replace("\n        ", "\n            "))
        out.append("    };")
    out.append("};")
    out.append("")


# (stages, singleton stores in the check phase, slot budget, executor rule, longest chains only, seed, name)
# VNC2 is the production plan (measured best of the family on B200, profiles/r02_vn_plans.md): two stages, 92 slots -- what
# fits two CTAs per SM next to the staging row --, executor = the column's first block row (mostly identity circulants);
# VNC2T: longest chains (72 slots); VNC1: no chains at all, one stage, 112 slots (no room for a staging row).
PLANS = [(2, True, 92, "top", False, 3, "VNC2"), (2, True, 72, "top", True, 1, "VNC2T"), (1, True, 112, "zero", True, 1, "VNC1")]
import os as _os
if _os.environ.get("QC_EXTRA_SEEDS"):        # plan lottery for A/B builds: QC_EXTRA_SEEDS="5 6 7" adds VNC2S5, VNC2S6, ...
    PLANS = PLANS + [(2, True, 92, "top", False, int(x), "VNC2S" + x) for x in _os.environ["QC_EXTRA_SEEDS"].split()]
if _os.environ.get("QC_EXTRA_CN"):           # ... QC_EXTRA_CN="1 2": VNC2B1, VNC2B2 balance the check phase too; VNC2A1...: any executor
    PLANS = PLANS + [(2, True, 92, "top+cn", False, int(x), "VNC2B" + x) for x in _os.environ["QC_EXTRA_CN"].split()]
    PLANS = PLANS + [(2, True, 92, "any+cn", False, int(x), "VNC2A" + x) for x in _os.environ["QC_EXTRA_CN"].split()]


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
