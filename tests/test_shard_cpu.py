"""World-size-2 gloo test of the N>1 host logic (frame sharding + counter all-reduce) on CPU.
The decoder behind it here is the CPU oracle -- this checks the plumbing, not the kernels."""
import os
import socket

import numpy as np
import pytest


def test_frame_range_partitions():
    from ecc_ldpc_b200.shard import frame_range
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 1000, 1001):
            got = [frame_range(r, world, 50, n) for r in range(world)]
            assert sum(c for _, c in got) == n
            nxt = 50
            for first, c in got:
                assert first == nxt
                nxt += c
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        frame_range(2, 2, 0, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import awgn_llrs
    from ecc_ldpc_b200 import codes
    from ecc_ldpc_b200.shard import allreduce_counters, frame_range
    from oracle import oracle
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    code = codes.load_builtin("moon.7.13")
    orc = oracle.Oracle(code)
    first, count = frame_range(rank, world, 100, total)
    # frame-index keyed inputs: every frame's LLRs depend only on its global index
    llr = np.stack([awgn_llrs(code, 2.0, 1, 1000 + f)[0] for f in range(first, first + count)]) if count else np.zeros((0, code.N))
    bits, iters, conv = orc.decode_batch("spa", 20, llr, threads=1)
    k = code.K
    cnt = torch.tensor([count, int(bits[:, :k].any(axis=1).sum()), int(bits[:, :k].sum()), int(iters.sum()), int((conv == 0).sum())],
                       dtype=torch.int64)
    allreduce_counters(cnt)
    np.save(os.path.join(out_dir, "cnt_%d_%d.npy" % (world, rank)), cnt.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    total = 301
    res = {}
    for world in (1, 2):
        port = _free_port()
        mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
        res[world] = [np.load(tmp_path / ("cnt_%d_%d.npy" % (world, r))) for r in range(world)]
    assert res[1][0][0] == total
    assert np.array_equal(res[2][0], res[2][1])              # every rank holds the global sum
    assert np.array_equal(res[2][0], res[1][0])              # and it does not depend on the number of ranks
