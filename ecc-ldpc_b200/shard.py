"""Multi-GPU sharding of the path: frames are independent, so rank r of W decodes a contiguous range of the global
frame index (Philox streams are keyed by that index, hence results do not depend on W), and the only exchange step is
a sum all-reduce of the five error counters {frames, frame_errors, bit_errors, sum_iters, unconverged}.
One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

COUNTER_NAMES = ("frames", "frame_errors", "bit_errors", "sum_iters", "unconverged")


def frame_range(rank: int, world: int, first_frame: int, n_frames: int):
    """Contiguous, balanced split of [first_frame, first_frame + n_frames): returns (first, count) of `rank`."""
    if world <= 0 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_frames, world)
    count = base + (1 if rank < rem else 0)
    first = first_frame + rank * base + min(rank, rem)
    return first, count


def allreduce_counters(t):
    """Sum all-reduce of a 5-element int64 tensor over the default process group (no-op without one)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def simulate_sharded(dec, max_iters, ebn0_db, k_info, n_tx, seed, first_frame, n_frames, early_exit=True, device=None):
    """Every rank simulates its frame range on its GPU (ldpc_b200_simulate_device accumulates straight into the
    all-reduce buffer), then one NCCL all-reduce; returns the global counters as a dict on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    first, count = frame_range(rank, world, first_frame, n_frames)
    ext = torch.cuda.ExternalStream(dec.stream())
    with torch.cuda.stream(ext):      # the zero-fill is ordered on the decoder's own stream, in front of the kernel's atomics
        cnt = torch.zeros(5, dtype=torch.int64, device=device or "cuda")
    if count:
        dec.simulate_device(cnt.data_ptr(), max_iters, ebn0_db, k_info, n_tx, seed, first, count, early_exit=early_exit)
    with torch.cuda.stream(ext):
        allreduce_counters(cnt)
    dec.sync()
    return dict(zip(COUNTER_NAMES, (int(v) for v in cnt.cpu())))
