#!/usr/bin/env python3
"""Turns gpurun_out/ev/ (written by tools/evidence.sh on a B200 box) into the committed files under profiles/:
ncu full-capture summaries, launch lists, bench JSON lines, traffic.json.  Run where `ncu` is installed."""
import collections, csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV, PR = os.path.join(ROOT, "gpurun_out", "ev"), os.path.join(ROOT, "profiles")
MD = os.path.join(ROOT, "tools", "ncu_report_md.py")

REPORTS = [
    ("prof_resident_min", "r01_ncu_resident_minsum_final.md",
     "ncu --set full: RESIDENT min-sum kernel k_resident<QcJpl4096P2, float, 1, MINSUM> (round 1 final: 32768 frames, all 20 iterations executed, Eb/N0 2.0 dB)",
     "ncu --set full --clock-control none --import-source on -k regex:k_resident -s 2 -c 1 python bench.py --steps 2 --warmup 1 --no-extra --no-cpu"),
    ("prof_resident_spa", "r01_ncu_resident_spa_final.md",
     "ncu --set full: RESIDENT sum-product kernel k_resident<QcJpl4096P2, float, 1, SPA> (round 1 final: 16384 frames, 20 iterations, Eb/N0 2.0 dB)",
     "ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 python tools/one_decode.py spa 2.0 16384 1"),
    ("prof_stream_cn", "r01_ncu_stream_cn_final.md",
     "ncu --set full: STREAM check-node kernel k_cn_tma<float, MINSUM, 18, 2> (jpl.4096.4.5, 8192 frames, one iteration)",
     "ncu --set full --clock-control none --import-source on -k regex:k_cn_tma -s 3 -c 1 python tools/stream_probe.py 8192"),
    ("prof_stream_vn", "r01_ncu_stream_vn_final.md",
     "ncu --set full: STREAM variable-node kernel k_vn_tma<float, 4> (jpl.4096.4.5, 8192 frames, one iteration)",
     "ncu --set full --clock-control none --import-source on -k regex:k_vn_tma -s 3 -c 1 python tools/stream_probe.py 8192"),
    ("prof_onchip", "r01_ncu_onchip_final.md",
     "ncu --set full: ONCHIP kernel k_onchip<float, MINSUM, 4, regular> (1920.1280.3.303, 65536 frames, 20 iterations, Eb/N0 0 dB)",
     "ncu --set full --clock-control none --import-source on -k regex:k_onchip -s 1 -c 1 python tools/stream_probe.py 65536 1920.1280.3.303 min auto"),
]


def launch_md(csvpath, title, cmd, out):
    lines = open(csvpath).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines[i:]):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v, u = float(r["Metric Value"]), r["Metric Unit"]
        us = v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)
        a = agg.setdefault(r["Kernel Name"], [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# %s\n\nCommand: `%s` (after the same command exited 0 without ncu). Times are cold-cache and serialised: compare shares.\n\n" % (title, cmd))
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.3f |\n" % (k[:110], a[0], a[1], a[1] / tot))


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return {h: rows[2][i] for i, h in enumerate(rows[0])}


def main():
    for stem, out, title, cmd in REPORTS:
        rep = os.path.join(EV, stem + ".ncu-rep")
        md = subprocess.run([sys.executable, MD, rep, title, cmd], capture_output=True, text=True).stdout
        open(os.path.join(PR, out), "w").write(md)
    launch_md(os.path.join(EV, "launches_bench.csv"), "ncu launch list, round 1 final kernels: bench.py",
              "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu",
              os.path.join(PR, "r01_launches_bench_final.md"))
    launch_md(os.path.join(EV, "launches_stream.csv"), "ncu launch list, round 1 final kernels: STREAM engine, one decode of 8192 frames jpl.4096.4.5 min-sum",
              "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv python tools/stream_probe.py 8192",
              os.path.join(PR, "r01_launches_stream_final.md"))
    for a, b in (("launches_bench.csv", "r01_launches_bench_final.csv"), ("launches_stream.csv", "r01_launches_stream_final.csv"),
                 ("bench_1gpu.json", "r01_bench_1gpu_final.json"), ("bench_reference.json", "r01_bench_reference_final.json"),
                 ("plain_bench_short.json", "r01_bench_short_for_launch_list.json")):
        shutil.copy(os.path.join(EV, a), os.path.join(PR, b))
    m = raw_metrics(os.path.join(EV, "prof_resident_min.ncu-rep"))
    f = lambda k: float(m[k])
    def to_bytes(key):
        # ncu prints these in Mbyte in the raw page of our captures
        return f(key) * 1e6
    traffic = {
        "dram_bytes_per_launch_per_frame": (to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")) / 32768,
        "ncu": {"smem_wavefronts_pct_of_peak": round(f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), 2),
                "issue_active_pct": round(f("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
                "lsu_pipe_pct": round(f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), 2),
                "alu_pipe_pct": round(f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 2),
                "fma_pipe_pct": round(f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2),
                "dram_throughput_pct": round(f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 2),
                "l2_sector_hit_pct": round(f("lts__t_sector_hit_rate.pct"), 2),
                "warps_per_scheduler": round(f("smsp__warps_active.avg.per_cycle_active"), 2)},
        "source": "profiles/r01_ncu_resident_minsum_final.md (one launch of 32768 frames under ncu --set full; dram = dram__bytes_read.sum + dram__bytes_write.sum)",
    }
    json.dump(traffic, open(os.path.join(PR, "traffic.json"), "w"), indent=1)
    print(json.dumps(traffic["ncu"]), traffic["dram_bytes_per_launch_per_frame"])


if __name__ == "__main__":
    main()
