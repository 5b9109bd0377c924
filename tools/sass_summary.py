#!/usr/bin/env python3
"""Static SASS summary of a kernel of libldpc_b200.so: opcode histogram and the lines that show the sm_100a features the
design relies on (UBLKCP = cp.async.bulk / TMA, SYNCS = mbarrier, FMNMX.XORSIGN = min', BAR.RED = vote, MUFU, LDS/STS).

    python tools/sass_summary.py <kernel-name-substring> [n excerpt lines per pattern]  > profiles/r02_sass_<engine>.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ecc-ldpc_b200", "libldpc_b200.so")
want = sys.argv[1]
nex = int(sys.argv[2]) if len(sys.argv) > 2 else 3
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
arch = re.search(r"arch = (sm_\w+)", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    if want not in dem and want not in name:
        continue
    ops = collections.Counter()
    lines = []
    for ln in b.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m:
            ops[m.group(2)] += 1
            lines.append(ln.rstrip())
    total = sum(ops.values())
    print("== %s\n   (%s; %d SASS instructions)" % (dem, arch.group(1) if arch else "?", total))
    fam = collections.Counter()
    for k, v in ops.items():
        fam[k.split(".")[0] if not k.startswith(("FMNMX.XORSIGN", "BAR.RED", "MUFU.")) else ".".join(k.split(".")[:2])] += v
    print("   " + ", ".join("%s %d" % kv for kv in fam.most_common(24)))
    for pat in ("UBLKCP", "SYNCS", "FMNMX.XORSIGN", "BAR.RED", "MUFU", "LDS", "STS", "ATOM", "RED."):
        hits = [l for l in lines if pat in l]
        if hits:
            print("   -- %s: %d" % (pat, len(hits)))
            for l in hits[:nex]:
                print("      " + re.sub(r"\s+", " ", l.strip()))
    print()
