#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions and stall samples per SASS opcode."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
st = collections.Counter()
for r in rows[2:]:
    if len(r) <= iex: continue
    try: ex = int(r[iex]); sm = int(r[isamp] or 0)
    except ValueError: continue
    src = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:12]
    op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "FMNMX", "LOP3", "LDS", "STS", "BAR", "ISETP", "FSETP")) else op.split(".")[0]
    ops[op] += ex; samp[op] += sm; tot += ex; tots += sm
    for i, h in stall_cols:
        try: st[h] += int(r[i] or 0)
        except ValueError: pass
print("total warp-instructions %d, samples %d" % (tot, tots))
for op, n in ops.most_common(28):
    print("%-16s %12d  %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(tots, 1)))
print("stall samples:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tots, 1)) for h, v in st.most_common(10)))
