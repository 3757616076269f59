"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: contiguous env slices with global Philox ids
reproduce the single-process trajectories, and the statistics all-reduce sums to the global statistics.  The
stepping engine here is the oracle (test infrastructure); on GPUs the same helpers shard the CUDA engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from colosseumrl_b200.sharding import shard_range, all_reduce_stats


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (f0, n0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + n0 == f1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total, K, seed, q):
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard_range(total, rank, world)
    tb = orc.TronBatch(n, 9, 4)
    tb.rollout(seed, first, 0, K, fresh=True, nthreads=1)
    tt = orc.TTTBatch(n, 4)
    tt.rollout(seed, first, 0, K, fresh=True, nthreads=1)
    red_tron = all_reduce_stats(torch.from_numpy(tb.stats))
    red_ttt = all_reduce_stats(torch.from_numpy(tt.stats))
    q.put((rank, first, n, tb.board.copy(), tb.deaths.copy(), tt.board.copy(), red_tron.numpy(), red_ttt.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rollout_equals_global(world):
    from oracle import oracle as orc
    total, K, seed = 101, 30, 13
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, K, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = orc.TronBatch(total, 9, 4); g.rollout(seed, 0, 0, K, fresh=True, nthreads=1)
    gt = orc.TTTBatch(total, 4); gt.rollout(seed, 0, 0, K, fresh=True, nthreads=1)
    for rank, first, n, board, deaths, tboard, red_tron, red_ttt in res:
        assert (board == g.board[first:first + n]).all() and (deaths == g.deaths[first:first + n]).all()
        assert (tboard == gt.board[first:first + n]).all()
        assert (red_tron == g.stats).all() and (red_ttt == gt.stats).all()
