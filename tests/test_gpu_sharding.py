"""Multi-rank parity on the CUDA engine: the same GLOBAL environment ids give identical
per-environment results for any GPU count.

W NCCL ranks (one process per GPU) each own the `shard_range` slice of one global batch and step it through the C ABI
(Tron: policy kernel + crl_tron_step; Blokus: legal + pick + step; Tic Tac Toe 4p: fused rollout); the end states are
compared bit-exactly with the 1-GPU run of the whole batch and with the CPU oracle, and the NCCL all-reduced episode
statistics with both.  Needs >= 2 GPUs (`gpurun --gpus 2` / `--gpus 4`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TRON_TOTAL, TRON_K = 10007, 48            # odd totals: uneven shards
BLOKUS_TOTAL, BLOKUS_K = 301, 72
TTT_TOTAL, TTT_K = 50021, 40
SEED = 11


def _run_shard(device, first, n):
    """Steps the slice [first, first + n) of the three global batches on `device`; returns numpy end states + the
    environment objects' (local) statistics tensors."""
    import torch
    from colosseumrl_b200 import BatchedTronGridEnvironment, BatchedBlokusEnvironment, BatchedTicTacToe4PlayerEnv
    out = {}
    f, m = first["tron"], n["tron"]
    env = BatchedTronGridEnvironment("", batch=m, device=device, seed=SEED, auto_reset=True, first_env_id=f)
    st, _ = env.new_state()
    for t in range(TRON_K):
        st = env.step_(st, env.random_actions(t), out=st)
    o = env.state_to_observation(st, -1)
    out["tron"] = (o["board"].cpu().numpy(), o["heads"].cpu().numpy(), o["deaths"].cpu().numpy(), st.result.cpu().numpy())
    tron_env = env
    f, m = first["blokus"], n["blokus"]
    env = BatchedBlokusEnvironment("", batch=m, device=device, seed=SEED, auto_reset=True, first_env_id=f)
    st, _ = env.new_state()
    for t in range(BLOKUS_K):
        valid = env.valid_actions(st)
        st = env.step_(st, env.random_actions(valid, t), out=st)
    o = env.state_to_observation(st, -1)
    out["blokus"] = (o["board"].cpu().numpy(), o["pieces"].cpu().numpy(), o["score"].cpu().numpy(), st.result.cpu().numpy())
    blokus_env = env
    f, m = first["ttt"], n["ttt"]
    env = BatchedTicTacToe4PlayerEnv("", batch=m, device=device, seed=SEED, auto_reset=True, first_env_id=f)
    st, _ = env.new_state()
    env.rollout(st, 0, TTT_K)
    board, winner, mover = env.state_arrays(st)
    out["ttt"] = (board.cpu().numpy(), winner.cpu().numpy(), mover.cpu().numpy())
    torch.cuda.synchronize(device)
    return out, {"tron": tron_env, "blokus": blokus_env, "ttt": env}


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from colosseumrl_b200.sharding import shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    totals = {"tron": TRON_TOTAL, "blokus": BLOKUS_TOTAL, "ttt": TTT_TOTAL}
    first, n = {}, {}
    for k, tot in totals.items():
        first[k], n[k] = shard_range(tot, rank, world)
    out, envs = _run_shard(dev, first, n)
    red = {k: e.all_reduce_stats().cpu().numpy() for k, e in envs.items()}      # NCCL all-reduce of the 32-slot vectors
    q.put((rank, first, n, out, red))
    dist.barrier()
    dist.destroy_process_group()


def _worlds():
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4) if w <= n]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_cuda_engine_equals_one_gpu_and_oracle(world):
    import torch
    import torch.multiprocessing as mp
    if world not in _worlds():
        pytest.skip("needs %d GPUs" % world)
    from oracle import oracle as orc
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # the whole batch on ONE GPU
    zero = {"tron": 0, "blokus": 0, "ttt": 0}
    one, envs = _run_shard(torch.device("cuda", 0), zero, {"tron": TRON_TOTAL, "blokus": BLOKUS_TOTAL, "ttt": TTT_TOTAL})
    one_stats = {k: e.stats.cpu().numpy() for k, e in envs.items()}
    # ... and on the CPU oracle
    tb = orc.TronBatch(TRON_TOTAL, 19, 4); tb.rollout(SEED, 0, 0, TRON_K, fresh=True)
    bb = orc.BlokusBatch(BLOKUS_TOTAL); bb.rollout(SEED, 0, 0, BLOKUS_K, fresh=True)
    tt = orc.TTTBatch(TTT_TOTAL, 4); tt.rollout(SEED, 0, 0, TTT_K, fresh=True)
    assert (one["tron"][0] == tb.board).all() and (one["tron"][2] == tb.deaths).all()
    assert (one["blokus"][0] == bb.board).all() and (one["blokus"][2] == bb.scores).all()
    assert (one["ttt"][0] == tt.board).all() and (one["ttt"][1] == tt.winner).all()
    assert (one_stats["tron"] == tb.stats).all() and (one_stats["blokus"] == bb.stats).all() and (one_stats["ttt"] == tt.stats).all()
    covered = {k: 0 for k in zero}
    for rank, first, n, out, red in res:
        for k in ("tron", "blokus", "ttt"):
            lo, hi = first[k], first[k] + n[k]
            covered[k] += n[k]
            for a, b in zip(out[k], one[k]):
                assert (a == b[lo:hi]).all(), (k, rank)
            assert (red[k] == one_stats[k]).all(), (k, rank)                   # all-reduced statistics == global run
    assert covered == {"tron": TRON_TOTAL, "blokus": BLOKUS_TOTAL, "ttt": TTT_TOTAL}


def test_two_environments_on_two_devices_in_one_process():
    """The C ABI launches on the device that owns the stream; environments on different GPUs coexist in one process and
    the caller's current device is never changed (advisor finding, round 1)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import oracle as orc
    from colosseumrl_b200 import BatchedTronGridEnvironment
    torch.cuda.set_device(0)
    envs = [BatchedTronGridEnvironment("", batch=1000, device="cuda:%d" % d, seed=3, auto_reset=True, first_env_id=1000 * d)
            for d in (0, 1)]
    assert torch.cuda.current_device() == 0
    states = [e.new_state()[0] for e in envs]
    for t in range(30):
        for e, st in zip(envs, states):
            e.step_(st, e.random_actions(t), out=st)
    assert torch.cuda.current_device() == 0
    ob = orc.TronBatch(2000, 19, 4); ob.rollout(3, 0, 0, 30, fresh=True)
    for d, (e, st) in enumerate(zip(envs, states)):
        o = e.state_to_observation(st, -1)
        assert o["board"].device.index == d
        assert (o["board"].cpu().numpy() == ob.board[1000 * d:1000 * d + 1000]).all()
    stepper = envs[1].host_stepper(states[1])
    rec = stepper()
    assert rec.shape == (1000, 8) and torch.cuda.current_device() == 0
