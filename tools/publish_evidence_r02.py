#!/usr/bin/env python3
"""Turns gpurun_out/ev2/ (written by tools/evidence_r02.sh on a B200 box) into the committed files under profiles/:
ncu full-capture summaries, launch lists, bench JSON lines, traffic.json (with the hash of the profiled library), SASS
summaries of the shipped cubins.  Run where `ncu` and `cuobjdump` are installed."""
import os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from publish_evidence import launch_md  # noqa: E402
EV, PR = os.path.join(ROOT, "gpurun_out", "ev2"), os.path.join(ROOT, "profiles")
MD = os.path.join(ROOT, "tools", "ncu_report_md.py")
NCU = "ncu --set full --clock-control none --import-source on"

REPORTS = [
    ("prof_resident_min", "r02_ncu_resident_minsum.md",
     "ncu --set full: RESIDENT min-sum kernel k_resident<WithPlan<QcJpl4096P2, VNC2>, float, 1, MINSUM, 2, false> (round 2: fused normalisation, two-stage variable phase with singleton slots; 32768 frames, all 20 iterations executed, Eb/N0 2.0 dB)",
     NCU + " -k regex:k_resident -s 2 -c 1 python bench.py --steps 2 --warmup 1 --no-extra --no-cpu"),
    ("prof_resident_spa", "r02_ncu_resident_spa.md",
     "ncu --set full: RESIDENT sum-product kernel k_resident<WithPlan<QcJpl4096P2, VNC2>, float, 1, SPA, 2, false> (round 2; 16384 frames, 20 iterations, Eb/N0 2.0 dB)",
     NCU + " -k regex:k_resident -s 1 -c 1 python tools/one_decode.py spa 2.0 16384 1"),
    ("prof_resident_min1024", "r02_ncu_resident_minsum_jpl1024.md",
     "ncu --set full: RESIDENT min-sum kernel k_resident<WithPlan<QcJpl1024P2, VNC2>, float, 4, MINSUM, 2, false> (BASELINE config 2: jpl.1024.4.5, 16384 frames, four frames per CTA, all 20 iterations, Eb/N0 2.0 dB)",
     NCU + " -k regex:k_resident -s 3 -c 1 python tools/quick_bench.py 16384 jpl.1024.4.5 min"),
    ("prof_latency_min", "r02_ncu_latency_minsum.md",
     "ncu --set full: latency instance k_resident<WithPlan<QcJpl4096P4, VNC1>, float, 1, MINSUM, 1, false> behind the mailbox (ONE frame = one CTA of 512 threads, 20 iterations, LLRs read from mapped host memory)",
     NCU + " -k regex:k_resident -s 4 -c 1 python tools/one_mailbox.py min 2.0 12"),
    ("prof_stream_cn", "r02_ncu_stream_cn.md",
     "ncu --set full: STREAM check-node kernel k_cn_tma<float, MINSUM, 18, 2> (jpl.4096.4.5, 8192 frames, one iteration)",
     NCU + " -k regex:k_cn_tma -s 3 -c 1 python tools/stream_probe.py 8192"),
    ("prof_stream_vn", "r02_ncu_stream_vn.md",
     "ncu --set full: STREAM variable-node kernel k_vn_tma<float, 4> (jpl.4096.4.5, 8192 frames, one iteration)",
     NCU + " -k regex:k_vn_tma -s 3 -c 1 python tools/stream_probe.py 8192"),
    ("prof_onchip", "r02_ncu_onchip.md",
     "ncu --set full: ONCHIP kernel k_onchip<float, MINSUM, 4, regular> (1920.1280.3.303, 65536 frames, 20 iterations, Eb/N0 0 dB)",
     NCU + " -k regex:k_onchip -s 1 -c 1 python tools/stream_probe.py 65536 1920.1280.3.303 min auto"),
]


def main():
    for stem, out, title, cmd in REPORTS:
        rep = os.path.join(EV, stem + ".ncu-rep")
        if not os.path.exists(rep):
            print("absent (other part of the evidence run), kept:", out)
            continue
        md = subprocess.run([sys.executable, MD, rep, title, cmd], capture_output=True, text=True)
        if md.returncode:
            print("FAILED", stem, md.stderr[-500:])
            continue
        open(os.path.join(PR, out), "w").write(md.stdout)
    launch_md(os.path.join(EV, "launches_bench.csv"), "ncu launch list, round 2 kernels: bench.py",
              "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu",
              os.path.join(PR, "r02_launches_bench.md"))
    if os.path.exists(os.path.join(EV, "launches_stream.csv")):
      launch_md(os.path.join(EV, "launches_stream.csv"), "ncu launch list, round 2 kernels: STREAM engine, decodes of 8192 frames jpl.4096.4.5 min-sum",
                "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv python tools/stream_probe.py 8192",
                os.path.join(PR, "r02_launches_stream.md"))
    for a, b in (("launches_bench.csv", "r02_launches_bench.csv"), ("launches_stream.csv", "r02_launches_stream.csv"),
                 ("bench_1gpu.json", "r02_bench_1gpu.json"), ("bench_reference.json", "r02_bench_reference.json"),
                 ("plain_bench_short.json", "r02_bench_short_for_launch_list.json")):
        if os.path.exists(os.path.join(EV, a)):
            shutil.copy(os.path.join(EV, a), os.path.join(PR, b))
    # traffic.json: only if the capture was taken from the library that is in the tree now
    import hashlib
    sys.path.insert(0, ROOT)
    from ecc_ldpc_b200 import buildinfo
    want = open(os.path.join(EV, "kernel_sources.sha256")).read().split()[0]
    have = buildinfo.kernel_source_sha256()
    raw = subprocess.run(["ncu", "-i", os.path.join(EV, "prof_resident_min.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(EV, "prof_resident_min_raw.csv"), "w").write(raw)
    if want == have:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), os.path.join(EV, "prof_resident_min_raw.csv"), "32768",
                               "profiles/r02_ncu_resident_minsum.md"])
    else:
        print("traffic.json NOT written: the capture is of other kernel sources (%s... vs %s...)" % (want[:12], have[:12]))
    for eng, pats in (("resident", ["QcJpl4096P2ENS2_4VNC2EEEfLi1ELi1ELi2ELb0", "QcJpl4096P2ENS2_4VNC2EEEfLi1ELi0ELi2ELb0", "QcJpl4096P4ENS2_4VNC1EEEfLi1ELi1ELi1ELb0"]),
                      ("stream", ["k_cn_tmaIfLi1ELi18ELi2", "k_vn_tmaIfLi4"]), ("onchip", ["k_onchipIfLi1ELi4ELb1ELi4"])):
        txt = "".join(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_summary.py"), p, "3"], capture_output=True, text=True).stdout for p in pats)
        open(os.path.join(PR, "r02_sass_%s.txt" % eng), "w").write(
            "# cuobjdump -sass of ecc-ldpc_b200/libldpc_b200.so (kernel sources sha256 %s), tools/sass_summary.py\n\n%s" % (have, txt))


if __name__ == "__main__":
    main()
