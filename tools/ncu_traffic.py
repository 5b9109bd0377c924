#!/usr/bin/env python3
"""profiles/traffic.json from an ncu capture of the CURRENT build: dram bytes of one launch of the headline kernel.

    ncu --set full --clock-control none -k regex:k_resident -c 1 -o gpurun_out/r02_resident python tools/quick_bench.py ...
    ncu -i gpurun_out/r02_resident.ncu-rep --page raw --csv > gpurun_out/r02_resident_raw.csv
    python tools/ncu_traffic.py gpurun_out/r02_resident_raw.csv <frames in the profiled launch> [capture name]

Records the SHA-256 of the kernel's SOURCES (bench.kernel_source_sha256: the files the RESIDENT kernels are compiled from plus
the Makefile) next to the numbers; bench.py reports roofline.traffic only when that hash is the hash of the sources in the
tree (a capture of other sources is stale by definition; the library file's own hash changes with every nvcc run)."""
import csv, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "shared_wavefronts_pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__inst_executed.sum": "warp_instructions", "launch__registers_per_thread": "registers_per_thread",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    path, frames = sys.argv[1], int(sys.argv[2])
    capture = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(path)
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    kernel = vals[names.index("Kernel Name")]
    got = {}
    for k, short in WANT.items():
        if k in names:
            i = names.index(k)
            got[short] = float(vals[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    so = os.path.join(ROOT, "ecc-ldpc_b200", "libldpc_b200.so")
    sys.path.insert(0, ROOT)
    from ecc_ldpc_b200 import buildinfo
    out = {"kernel_source_sha256": buildinfo.kernel_source_sha256(), "so_sha256_at_capture": hashlib.sha256(open(so, "rb").read()).hexdigest(), "capture": capture, "kernel": kernel, "frames_in_launch": frames,
           "dram_bytes_per_launch": got["dram_read"] + got["dram_write"],
           "dram_bytes_per_launch_per_frame": (got["dram_read"] + got["dram_write"]) / frames,
           "ncu": {k: v for k, v in got.items() if k not in ("dram_read", "dram_write")}}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
