#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path: jpl.4096.4.5 normalised min-sum, 20 iterations.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU restatement of the reference on the host cores)

One "step" = one pass of the decoder over one batch of B frames per GPU whose LLRs are already resident
in HBM (float32, generated on the device before the timed region).  The workload is the WORST CASE of the
metric: Eb/N0 = 2.0 dB, where no frame of this code converges, so the reference algorithm itself (syndrome
test before every iteration, Orig.hs:68-70) executes all 20 iterations on every frame -- nothing is skipped
and `avg_iters` in the JSON line proves it.  Early-exit operating points are reported under "extra".

value      decoded info Gbit/s = frames * K_info / time, whole job (all ranks), device-timed (CUDA events on the
           decoder's stream, barrier + synchronize on both sides, max over ranks)
e2e        same metric through ldpc_b200_decode with HOST float64 LLRs (pinned) in and host decisions out
roofline   HBM roofline of the decode kernel with the reference's algorithmic bytes (3E+3N words per frame
           and iteration); the RESIDENT engine keeps that state on chip, so frac may exceed 1
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODE = "jpl.4096.4.5"
MAX_ITERS = 20
N_TX = 5120
EBN0_WORST = 2.0
SEED = 20261018
METRIC = "decoded_info_gbit_per_s"
UNIT = "Gbit/s"


def workload_config(frames_per_step_per_gpu, n_values_per_frame):
    """The `config` object of the JSON line: the SAME dictionary on both arms (the CPU arm times a bounded sample of every
    step of this workload and says so in `cpu_baseline.sample`)."""
    return {"workload": "%s min-sum alpha=3/4, max %d iterations, rate 4/5 (n_tx=%d), all-zero codeword, BPSK+AWGN Eb/N0=%.1f dB (no frame converges: all %d iterations run)" % (CODE, MAX_ITERS, N_TX, EBN0_WORST, MAX_ITERS),
            "frames_per_step_per_gpu": frames_per_step_per_gpu,
            "frames": "Philox4x32-10 keyed by seed %d, stream = global frame index; rank r owns frames [r*B, (r+1)*B)" % SEED,
            "l2": "inputs larger than L2 (%.0f MB of LLRs per step per GPU)" % (frames_per_step_per_gpu * n_values_per_frame * 4 / 1e6),
            "sharding": "contiguous global frame ranges per rank; one NCCL all-reduce of 5 x u64 error counters per step"}


def kernel_source_sha256():
    """Identity of the headline kernel for roofline.traffic (ecc-ldpc_b200/buildinfo.py)."""
    from ecc_ldpc_b200 import buildinfo
    return buildinfo.kernel_source_sha256()


def alg_bytes_per_frame(code, iters, w=4):
    """B_alg = iters * w * (3E + 3N) + w * N_tx + K/8."""
    return iters * w * (3 * code.E + 3 * code.N) + w * N_TX + code.K // 8


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def reference_arm(args):
    """CPU arm: the C restatement of the reference decoder (oracle/, kind 'port' -- the Haskell reference cannot be
    built in this image) on all host cores, on the GPU arm's workload: same code / algorithm / iteration cap / Eb/N0 and
    the SAME frames (the Philox stream of the device generator, drawn on the host by ecc-ldpc_b200/philox_host.py).
    Each step = a bounded sample of the step's frames_per_step_per_gpu frames (the first `--cpu-frames`, default 64 per
    core), so that the run ends within minutes; frames/s do not depend on the sample size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ecc_ldpc_b200 import codes, philox_host
    from oracle import oracle
    code = codes.load_builtin(CODE)
    orc = oracle.Oracle(code)
    cores = oracle.lib().oracle_max_threads()
    frames = args.cpu_frames or 64 * cores
    llr = philox_host.awgn_llrs(code.N, EBN0_WORST, code.K, N_TX, SEED, 0, frames).astype(np.float64)
    for _ in range(args.warmup):
        orc.decode_batch("min", MAX_ITERS, llr[: max(4 * cores, 8)], variant=1, threads=cores)
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.steps):
        bits, iters, conv = orc.decode_batch("min", MAX_ITERS, llr, variant=1, threads=cores)
        its += int(iters.sum())
    dt = time.perf_counter() - t0
    fps = frames * args.steps / dt
    val = fps * code.K / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.frames, code.N),
        "avg_iters": its / (frames * args.steps),
        "edge_updates_per_s": fps * (its / (frames * args.steps)) * code.E,
        "frames_per_s": fps,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "every step decodes the first %d of the step's %d frames (same Philox frames as the GPU arm, float64), %d steps after %d warm-up passes; two-minimum O(E) check update, one frame per thread, %d threads" % (frames, args.frames, args.steps, args.warmup, cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def bind_to_gpu_numa_node(torch, local):
    """Run this rank (and allocate its pinned host buffers) on the NUMA node its GPU hangs off, so that the e2e leg's
    H2D/D2H copies of 8 ranks do not all cross the socket interconnect. Best effort; returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints (e.g. NCCL's version banner) was
    redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=32768, help="frames per step per GPU (LLRs resident in HBM)")
    ap.add_argument("--e2e-frames", type=int, default=8192, help="frames per step per GPU for the host-buffer e2e leg")
    ap.add_argument("--cpu-frames", type=int, default=0)
    ap.add_argument("--engine", default="auto", choices=["auto", "stream", "resident"])
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from ecc_ldpc_b200 import codes, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libldpc_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")       # host-side rendezvous that leaves the GPUs idle (see the simulate_multi leg)
    lib.init(local)

    code = codes.load_builtin(CODE)
    dcode = lib.DeviceCode(code)
    B, W = args.frames, code.N // 32
    engine = {"auto": lib.ENGINE_AUTO, "stream": lib.ENGINE_STREAM, "resident": lib.ENGINE_RESIDENT}[args.engine]
    dec = lib.Decoder(dcode, "min", "fp32", max_batch=B, engine=engine)
    ext = torch.cuda.ExternalStream(dec.stream(), device=torch.device("cuda", local))

    # device buffers as torch tensors (torch = memory + streams + NCCL plumbing only)
    d_llr = torch.empty((B, code.N), dtype=torch.float32, device="cuda")
    d_bits = torch.zeros((B, W), dtype=torch.int32, device="cuda")
    d_iters = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_conv = torch.zeros(B, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(5, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def gen(ebn0, first):
        dec.generate_llr(d_llr.data_ptr(), ebn0, code.K, N_TX, SEED, first, B)
        dec.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """K steps between barrier+synchronize, CUDA events on the decoder's stream, MAX over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    kernel_ms, launches = [], []

    def step_device(early_exit=True):
        def fn():
            dec.decode_device(d_llr.data_ptr(), d_bits.data_ptr(), MAX_ITERS, B, early_exit=early_exit,
                              d_iters=d_iters.data_ptr(), d_conv=d_conv.data_ptr())
            dec.count_errors(d_cnt.data_ptr(), B, code.K, d_bits.data_ptr(), d_iters.data_ptr(), d_conv.data_ptr())
            if world > 1:
                with torch.cuda.stream(ext):      # the path's one exchange step: 5 x u64 error counters over NCCL
                    dist.all_reduce(d_cnt)
        return fn

    # ---------------- headline: worst case, all 20 iterations executed ----------------
    gen(EBN0_WORST, rank * B)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    d_cnt.zero_()
    total_ms = timed(step_device(True), args.steps, args.warmup)
    # kernel-only duration of one launch, measured live by the library's own events on the launching stream
    for _ in range(3):
        dec.decode_device(d_llr.data_ptr(), d_bits.data_ptr(), MAX_ITERS, B, early_exit=True,
                          d_iters=d_iters.data_ptr(), d_conv=d_conv.data_ptr())
        dec.sync()
        ms, nl = dec.last_timing()
        kernel_ms.append(ms); launches.append(nl)
    clocks = sampler.stop()
    avg_iters = float(d_iters.float().mean().item())
    frames_total = B * world * args.steps
    fps = frames_total / (total_ms * 1e-3)
    value = fps * code.K / 1e9
    k_ms = float(np.median(kernel_ms))
    b_alg = alg_bytes_per_frame(code, avg_iters)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = B * b_alg / (k_ms * 1e-3) / 1e9
    # roofline.traffic: dram bytes of ONE launch of the decode kernel from an `ncu --set full` capture, written to
    # profiles/traffic.json by tools/ncu_traffic.py together with the SHA-256 of the kernel's sources at capture time.
    # A capture of other sources is NOT reported: traffic = null and the reason is printed.
    traffic, ncu_static, traffic_note = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("kernel_source_sha256") != kernel_source_sha256():
            traffic_note = "profiles/traffic.json was captured from other kernel sources (sha256 %s...); re-run tools/evidence_r02.sh + tools/publish_evidence_r02.py" % str(tj.get("kernel_source_sha256"))[:12]
            print("bench.py: " + traffic_note, file=sys.stderr)
        else:
            traffic = tj["dram_bytes_per_launch_per_frame"] * B
            ncu_static = tj.get("ncu")
            traffic_note = "ncu capture %s of kernel %s, same kernel sources (sha256 %s...)" % (tj.get("capture"), tj.get("kernel"), tj["kernel_source_sha256"][:12])
    except Exception as ex:
        traffic_note = "no usable profiles/traffic.json (%r)" % (ex,)

    # ---------------- e2e: the reference-facing plugin call, host buffers in and out ----------------
    # ECC record `decode` (Utils.hs:62-72) of "ldpc/b200-minsum/jpl.4096.4.5/20/4/5": c_length = 5120 received float64
    # LLRs per frame in (pinned host memory), the first message_length = 4096 decided bits per frame out.
    from ecc_ldpc_b200 import ecc as ecc_mod
    Be = args.e2e_frames
    rec = ecc_mod.match("ldpc/b200-minsum/%s/%d/4/5" % (CODE, MAX_ITERS), max_batch=min(Be, 4096))[0]
    assert rec.codeword_length == N_TX and rec.message_length == code.K, (rec.codeword_length, rec.message_length)
    h_rx = torch.empty((Be, N_TX), dtype=torch.float64).pin_memory()
    h_rx.copy_(d_llr[:Be, :N_TX].double().cpu())
    h_msg = torch.empty((Be, code.K), dtype=torch.uint8).pin_memory()
    h_flags = torch.empty(Be, dtype=torch.uint8).pin_memory()
    h_iters = torch.empty(Be, dtype=torch.int32).pin_memory()
    h_conv = torch.empty(Be, dtype=torch.uint8).pin_memory()
    L = lib.load()

    def step_e2e():
        lib.check(L.ldpc_b200_ecc_decode(rec.h, Be, h_rx.data_ptr(), N_TX, h_msg.data_ptr(), h_flags.data_ptr()))

    for _ in range(max(1, args.warmup)):
        step_e2e()
    barrier()
    e2e_steps = max(5, args.steps)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_fps = Be * world * e2e_steps / float(dt.item())
    # the message bits of the e2e call must be the ones the device-resident path decided for the same frames
    dec.decode_device(d_llr.data_ptr(), d_bits.data_ptr(), MAX_ITERS, B, early_exit=True, d_iters=d_iters.data_ptr(), d_conv=d_conv.data_ptr())
    dec.sync()
    ref_bits = lib.unpack_bits(d_bits[:256].cpu().numpy().view(np.uint32), code.N)[:, :code.K]
    e2e_same = bool((h_msg[:256].numpy() == ref_bits).all())
    e2e = {"value": e2e_fps * code.K / 1e9, "unit": UNIT, "h2d_bytes_per_step": Be * N_TX * 8,
           "d2h_bytes_per_step": Be * code.K, "frames_per_step": Be, "matches_device_path": e2e_same,
           "api": "ldpc_b200_ecc_decode = ECC.decode of ldpc/b200-minsum/%s/%d/4/5 (float64 received LLRs in pinned host memory; message bits back)" % (CODE, MAX_ITERS)}

    # ---------------- what the platform can move: concurrent pinned H2D copies on all ranks ----------------
    # the e2e leg is bound by the host -> device hop (41 KB of float64 per frame); this is its ceiling at N ranks
    extra = {}
    cp_bytes = 256 << 20
    h_big = torch.empty(cp_bytes, dtype=torch.uint8).pin_memory()
    d_big = torch.empty(cp_bytes, dtype=torch.uint8, device="cuda")
    d_big.copy_(h_big, non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(8):
        d_big.copy_(h_big, non_blocking=True)
    c1.record()
    barrier()
    cms = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device="cuda")
    own_gbs = 8 * cp_bytes / (float(cms.item()) * 1e-3) / 1e9
    if world > 1:
        dist.all_reduce(cms, op=dist.ReduceOp.MAX)
    agg_gbs = world * 8 * cp_bytes / (float(cms.item()) * 1e-3) / 1e9
    e2e_h2d_gbs = e2e_fps * N_TX * 8 / 1e9
    e2e["h2d_gb_per_s"] = e2e_h2d_gbs
    e2e["pinned_h2d_ceiling_gb_per_s"] = agg_gbs
    e2e["fraction_of_h2d_ceiling"] = e2e_h2d_gbs / agg_gbs
    extra["pinned_h2d_ceiling"] = {"ranks": world, "aggregate_gb_per_s": agg_gbs, "rank0_gb_per_s": own_gbs, "bytes_per_copy": cp_bytes, "copies": 8,
                                   "note": "all ranks copy 256 MiB of pinned host memory to their GPU at the same time, 8 times; aggregate = N x bytes / slowest rank"}
    del h_big, d_big

    # ---------------- single-process multi-GPU route (what a one-process host such as the Haskell tester uses) ----------------
    # ldpc_b200_simulate_multi with n_gpus = N from rank 0 alone, after the timed regions, against the per-rank path's counters
    if world > 1:
        barrier()
    sm_frames = 4096 * world
    d_cnt.zero_()
    torch.cuda.synchronize()
    first, count = rank * 4096, 4096
    dec.simulate_device(d_cnt.data_ptr(), MAX_ITERS, 3.5, code.K, N_TX, SEED, first, count)
    dec.sync()
    if world > 1:
        dist.all_reduce(d_cnt)
    per_rank_counters = [int(v) for v in d_cnt.cpu()]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(group=cpu_group)       # every GPU is idle from here until rank 0 is done driving all of them
    if rank == 0:
        try:
            t0 = time.perf_counter()
            got = lib.simulate_multi(dcode, world, "min", "fp32", 4096, MAX_ITERS, 3.5, code.K, N_TX, SEED, 0, sm_frames)
            dt_sm = time.perf_counter() - t0
            got_l = [int(got[k]) for k in ("frames", "frame_errors", "bit_errors", "sum_iters", "unconverged")]
            extra["simulate_multi_single_process"] = {"n_gpus": world, "frames": sm_frames, "counters": got_l, "per_rank_path_counters": per_rank_counters,
                                                      "identical": got_l == per_rank_counters, "first_call_s": dt_sm, "kernel_ms_max_over_gpus": got["kernel_ms"]}
            if got_l != per_rank_counters:
                print("bench.py: simulate_multi counters differ from the per-rank path: %r vs %r" % (got_l, per_rank_counters), file=sys.stderr)
            lib.load().ldpc_b200_simulate_multi_release()
        except Exception as ex:      # noqa: BLE001
            extra["simulate_multi_single_process"] = {"n_gpus": world, "error": str(ex)}
    if world > 1:
        dist.barrier(group=cpu_group)

    # ---------------- extra operating points with early exit (reference semantics) ----------------
    if not args.no_extra:
        # float32 LLRs in / packed decisions out: half the PCIe bytes of the reference-typed call
        h32 = torch.empty((Be, code.N), dtype=torch.float32).pin_memory()
        h32.copy_(d_llr[:Be].cpu())
        hpk = torch.empty((Be, W), dtype=torch.int32).pin_memory()

        def step_e2e32():
            lib.check(L.ldpc_b200_decode_f32(dec.h, MAX_ITERS, Be, h32.data_ptr(), hpk.data_ptr(), h_iters.data_ptr(), h_conv.data_ptr()))
        step_e2e32()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e32()
        torch.cuda.synchronize()
        dt32 = time.perf_counter() - t0
        extra["e2e_f32_packed"] = {"info_gbit_per_s": Be * e2e_steps / dt32 * code.K / 1e9, "h2d_bytes_per_step": Be * code.N * 4,
                                   "d2h_bytes_per_step": Be * (W * 4 + 5), "api": "ldpc_b200_decode_f32", "ranks_timed": 1}
        for ebn0 in (3.5, 4.0, 4.5):
            gen(ebn0, rank * B)
            d_cnt.zero_()
            ms = timed(step_device(True), max(3, args.steps // 2), 1)
            n = max(3, args.steps // 2)
            it = float(d_iters.float().mean().item())
            f = B * world * n / (ms * 1e-3)
            cnt = d_cnt.clone()
            extra["ebn0_%.1f" % ebn0] = {"info_gbit_per_s": f * code.K / 1e9, "frames_per_s": f, "avg_iters": it,
                                          "edge_updates_per_s": f * it * code.E,
                                          "fer": float((d_conv == 0).float().mean().item())}
        gen(4.0, rank * B)
        ms = timed(step_device(False), 3, 1)
        f = B * world * 3 / (ms * 1e-3)
        extra["ebn0_4.0_no_early_exit"] = {"info_gbit_per_s": f * code.K / 1e9, "frames_per_s": f, "avg_iters": float(MAX_ITERS),
                                           "edge_updates_per_s": f * MAX_ITERS * code.E}

        # ---- the other BASELINE.json configurations and engines, device-resident, kernel-timed (rank 0 only) ----
        if rank == 0:
            def guarded(fn):
                # the legs below are reported beside the headline; none of them may take the headline line down with it
                try:
                    return fn()
                except Exception as ex:      # noqa: BLE001
                    print("bench.py: extra leg failed: %r" % (ex,), file=sys.stderr)
                    return {"error": repr(ex)}

            def run_cfg(code_name, algo, engine_id, frames, max_it, ebn0, n_tx, reps=3):
                c2 = code if code_name == CODE else codes.load_builtin(code_name)
                dc2 = dcode if code_name == CODE else lib.DeviceCode(c2)
                de2 = lib.Decoder(dc2, algo, "fp32", max_batch=frames, engine=engine_id)
                W2 = (c2.N + 31) // 32
                tl = torch.empty((frames, c2.N), dtype=torch.float32, device="cuda")
                tb = torch.zeros((frames, W2), dtype=torch.int32, device="cuda")
                ti = torch.zeros(frames, dtype=torch.int32, device="cuda")
                tc = torch.zeros(frames, dtype=torch.uint8, device="cuda")
                de2.generate_llr(tl.data_ptr(), ebn0, c2.K, n_tx, SEED, 0, frames); de2.sync()
                ms = []
                for _ in range(reps + 1):
                    de2.decode_device(tl.data_ptr(), tb.data_ptr(), max_it, frames, early_exit=True, d_iters=ti.data_ptr(), d_conv=tc.data_ptr())
                    de2.sync()
                    ms.append(de2.last_timing()[0])
                m = float(np.median(ms[1:]))
                it = float(ti.float().mean().item())
                f = frames / (m * 1e-3)
                out = {"engine": {1: "stream", 2: "resident", 3: "onchip"}[de2.engine()], "frames": frames, "max_iters": max_it, "ebn0_db": ebn0,
                       "kernel_ms": m, "frames_per_s": f, "info_gbit_per_s": f * c2.K / 1e9, "avg_iters": it,
                       "edge_updates_per_s": f * it * c2.E, "fer": float((tc == 0).float().mean().item())}
                ba = it * 4 * (3 * c2.E + 3 * c2.N) + 4 * n_tx + c2.K // 8
                out["alg_gb_per_s"] = f * ba / 1e9
                out["hbm_frac"] = out["alg_gb_per_s"] / peak
                de2.close()
                return out
            extra["spa_jpl4096_20it_worst"] = guarded(lambda: run_cfg(CODE, "spa", lib.ENGINE_AUTO, 16384, 20, EBN0_WORST, N_TX))
            extra["spa_jpl4096_50it_3.5dB"] = guarded(lambda: run_cfg(CODE, "spa", lib.ENGINE_AUTO, 16384, 50, 3.5, N_TX))
            extra["minsum_jpl4096_50it_3.5dB"] = guarded(lambda: run_cfg(CODE, "min", lib.ENGINE_AUTO, 32768, 50, 3.5, N_TX))
            extra["minsum_jpl1024_20it_16k_worst"] = guarded(lambda: run_cfg("jpl.1024.4.5", "min", lib.ENGINE_AUTO, 16384, 20, EBN0_WORST, 1280))
            extra["minsum_jpl1024_20it_16k_4.0dB"] = guarded(lambda: run_cfg("jpl.1024.4.5", "min", lib.ENGINE_AUTO, 16384, 20, 4.0, 1280))
            # the HBM-streaming engine (any H): the kernel whose roofline IS the HBM roof
            # (at the headline's batch size; 8192 frames: 0.76 / 0.69 -- the load pass and the clears weigh more)
            extra["stream_minsum_jpl4096_20it_worst"] = guarded(lambda: run_cfg(CODE, "min", lib.ENGINE_STREAM, 32768, 20, EBN0_WORST, N_TX, reps=2))
            extra["stream_spa_jpl4096_20it_worst"] = guarded(lambda: run_cfg(CODE, "spa", lib.ENGINE_STREAM, 32768, 20, EBN0_WORST, N_TX, reps=2))
            extra["stream_minsum_1920.1280.3.303_20it_worst"] = guarded(lambda: run_cfg("1920.1280.3.303", "min", lib.ENGINE_STREAM, 32768, 20, 0.0, 1920))
            # the run-time table on-chip engine (general codes whose frame state fits shared memory)
            extra["onchip_minsum_1920.1280.3.303_20it_worst"] = guarded(lambda: run_cfg("1920.1280.3.303", "min", lib.ENGINE_AUTO, 65536, 20, 0.0, 1920))
            extra["onchip_minsum_1920.1280.3.303_20it_2.5dB"] = guarded(lambda: run_cfg("1920.1280.3.303", "min", lib.ENGINE_AUTO, 65536, 20, 2.5, 1920))
            extra["onchip_spa_moon.7.13_20it_3.0dB"] = guarded(lambda: run_cfg("moon.7.13", "spa", lib.ENGINE_AUTO, 262144, 20, 3.0, 20))

            def tester_legs():
                # the reference tester's own calling pattern: ECC.decode, ONE frame per call (Utils.hs:68-69), host buffers
                one_rx = h_rx[:1].clone().pin_memory(); one_out = torch.empty((1, code.K), dtype=torch.uint8).pin_memory()
                one_fl = torch.empty(1, dtype=torch.uint8).pin_memory()
                for _ in range(20):
                    lib.check(L.ldpc_b200_ecc_decode(rec.h, 1, one_rx.data_ptr(), N_TX, one_out.data_ptr(), one_fl.data_ptr()))
                t0 = time.perf_counter()
                for _ in range(200):
                    lib.check(L.ldpc_b200_ecc_decode(rec.h, 1, one_rx.data_ptr(), N_TX, one_out.data_ptr(), one_fl.data_ptr()))
                us = (time.perf_counter() - t0) / 200 * 1e6
                extra["single_frame_call"] = {"api": "ldpc_b200_ecc_decode, n_frames = 1 (all 20 iterations run), one caller: mailbox, latency kernel", "us_per_call": us,
                                              "frames_per_s": 1e6 / us, "info_gbit_per_s": code.K / us / 1e3}

                # the tester's pattern at its own concurrency (Utils.hs:53: maxThreadCount decoders; B200.hs registers 16): T host
                # threads, one ECC record each, one frame per ECC.decode call -- measured inside the library's C++ harness
                # (ldpc_b200_ecc_tester_pattern), so no Python is in the loop; frames coalesce in the shared mailbox
                name = "ldpc/b200-minsum/%s/%d/4/5" % (CODE, MAX_ITERS)
                tp = {}
                for label, ebn0 in (("worst_2.0dB_all_20_iterations", EBN0_WORST), ("4.0dB_early_exit", 4.0)):
                    gen(ebn0, 0)
                    nfp = 8192
                    rxp = d_llr[:nfp, :N_TX].double().cpu().numpy()
                    ref_bits = None
                    for T in (1, 4, 16, 32):
                        nft = nfp if T >= 4 else 2048
                        msg, secs, nl, nfr = lib.tester_pattern(codes.builtin_codes_dir(), name, T, rxp[:nft])
                        if ref_bits is None:
                            dec.decode_device(d_llr.data_ptr(), d_bits.data_ptr(), MAX_ITERS, B, early_exit=True, d_iters=d_iters.data_ptr(), d_conv=d_conv.data_ptr())
                            dec.sync()
                            ref_bits = lib.unpack_bits(d_bits[:nfp].cpu().numpy().view(np.uint32), code.N)[:, :code.K]
                        tp["%s_threads_%d" % (label, T)] = {"threads": T, "frames": nft, "frames_per_s": nft / secs, "us_per_frame_per_thread": secs / (nft / T) * 1e6,
                                                             "info_gbit_per_s": nft / secs * code.K / 1e9, "kernel_launches": nl,
                                                             "frames_per_launch": (nfr / nl) if nl else None,
                                                             "decisions_identical_to_batched_path": bool((msg == ref_bits[:nft]).all())}
                tp["round1_single_caller_frames_per_s"] = 12500
                extra["tester_pattern"] = tp
                gen(EBN0_WORST, rank * B)
            guarded(tester_legs)

    # ---------------- CPU baseline beside it (rank 0, bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import oracle
            from ecc_ldpc_b200 import philox_host
            orc = oracle.Oracle(code)
            cores = oracle.lib().oracle_max_threads()
            gen(EBN0_WORST, 0)
            nf = args.cpu_frames or 64 * cores
            sample = d_llr[:nf].double().cpu().numpy()
            # the host statement of the generator draws the same frames (the CPU arm of --impl reference relies on it)
            host = philox_host.awgn_llrs(code.N, EBN0_WORST, code.K, N_TX, SEED, 0, 64)
            gen_err = float(np.abs(host - sample[:64]).max() / np.abs(sample[:64]).max())
            orc.decode_batch("min", MAX_ITERS, sample[: 4 * cores], variant=1, threads=cores)          # warm-up: threads, caches, page faults
            reps = []
            for _ in range(3):
                t0 = time.perf_counter()
                rb, ri, rc = orc.decode_batch("min", MAX_ITERS, sample, variant=1, threads=cores)
                reps.append(time.perf_counter() - t0)
            dtc = float(np.median(reps))
            dec.decode_device(d_llr.data_ptr(), d_bits.data_ptr(), MAX_ITERS, B, early_exit=True, d_iters=d_iters.data_ptr(), d_conv=d_conv.data_ptr())
            dec.sync()
            gb = lib.unpack_bits(d_bits[:nf].cpu().numpy().view(np.uint32), code.N)
            same = float((gb == rb).all(axis=1).mean())
            cpu = {"value": nf / dtc * code.K / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d frames of the same LLRs (downloaded from the GPU), C restatement oracle/ldpc_oracle.c, two-minimum check update, %d threads, avg %.1f iterations; median of 3 passes after a warm-up pass (%s s)" % (
                       nf, cores, float(ri.mean()), ", ".join("%.3f" % r for r in reps)),
                   "frames_per_s": nf / dtc, "decisions_identical_to_gpu": same,
                   "host_generator_max_rel_diff_vs_device": gen_err}
        except Exception as ex:      # noqa: BLE001
            print("bench.py: cpu_baseline leg failed: %r" % (ex,), file=sys.stderr)
            cpu = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, code.N),
            "avg_iters": avg_iters, "engine": {1: "stream", 2: "resident", 3: "onchip"}[dec.engine()], "numa_node_rank0": numa,
            "frames_per_s": fps, "edge_updates_per_s": fps * avg_iters * code.E,
            "gpu_launches": int(sum(launches) / len(launches) + 1) * args.steps,
            "clocks": clocks,
            "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_note, "kernel_ms": k_ms, "alg_bytes_per_frame": b_alg,
                         # the frame state is on chip, so beside the algorithmic figure report what HBM really
                         # carries (ncu dram bytes / live kernel time) and the resources that do bind (from the committed capture)
                         "measured_dram_gbs": (traffic / (k_ms * 1e-3) / 1e9) if traffic else None,
                         "binding_resources_ncu": ncu_static,
                         "note": "algorithmic bytes = reference state model 3E+3N words/frame/iteration (SURVEY.md 8d); peak = MEASURED_PEAKS.json hbm_gbs%s" % ("" if peaks else " (fallback)")},
            "cpu_baseline": cpu,
            "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
