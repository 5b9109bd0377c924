#!/usr/bin/env python3
"""Markdown summary of one kernel of an .ncu-rep (run where `ncu` is installed):
    python tools/ncu_report_md.py <file.ncu-rep> "<title>" "<command that was profiled>" > profiles/<name>.md
Key metrics from `--page raw`, opcode mix and stall samples from `--page source`."""
import collections
import csv
import io
import re
import subprocess
import sys

rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
KEYS = """Kernel Name
Grid Size
Block Size
gpu__time_duration.sum
dram__bytes_read.sum
dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct
launch__registers_per_thread
launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__inst_executed.sum
sm__icc_request_hit_rate.pct
smsp__warps_eligible.avg.per_cycle_active
smsp__warps_active.avg.per_cycle_active""".split("\n")
print("# %s\n" % title)
print("Command: `%s`\n" % cmd)
print("| metric | value | unit |\n|---|---|---|")
for k in KEYS:
    if k in get:
        print("| %s | %s | %s |" % (k, get[k][0], get[k][1]))
for k in sorted(get):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            if float(get[k][0]) >= 0.05:
                print("| %s | %.3f | per issue |" % (k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", ""), float(get[k][0])))
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
isrc, iex, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
ops = collections.Counter(); tot = 0; tots = 0; st = collections.Counter()
for r in rows[2:]:
    try:
        ex = int(r[iex]); sm = int(r[isamp] or 0)
    except (ValueError, IndexError):
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc].strip())
    op = m.group(2) if m else r[isrc].strip()[:12]
    op = ".".join(op.split(".")[:2]) if op.startswith(("FMNMX", "MUFU", "BAR", "UBLKCP", "SYNCS")) else op.split(".")[0]
    ops[op] += ex; tot += ex; tots += sm
    for i, c in stall_cols:
        try:
            st[c] += int(r[i] or 0)
        except ValueError:
            pass
print("\nExecuted warp-instructions by opcode (total %d):\n" % tot)
print("| opcode | share |\n|---|---|")
for op, n in ops.most_common(14):
    print("| %s | %.1f %% |" % (op, 100.0 * n / max(tot, 1)))
print("\nStall samples: " + ", ".join("%s %.1f %%" % (c[6:], 100.0 * v / max(tots, 1)) for c, v in st.most_common(8)))
