"""The N>1 host path on CPU: two gloo ranks shard a job and all-reduce episode statistics
exactly the way the NCCL ranks do on the GPU box (same functions, CPU tensors)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total_envs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rl_env_b200 import all_reduce_stats, shard_range
    start, count = shard_range(total_envs, rank, world)
    # rank-local statistics vector as plantos_stats lays it out: one finished episode per env
    ids = torch.arange(start, start + count, dtype=torch.float64)
    local = torch.stack([torch.tensor(float(count), dtype=torch.float64), ids.sum(), 1000.0 * ids.numel() + 0 * ids.sum(),
                         (ids % 7).sum(), (ids % 3).sum(), (ids % 5).sum(),
                         torch.tensor(0.0, dtype=torch.float64), torch.tensor(float(count), dtype=torch.float64)])
    out = all_reduce_stats(local)
    assert out is not local  # the rank-local vector is left untouched
    pending = all_reduce_stats(local, async_op=True)      # the overlapped form bench.py uses: same sums after wait()
    assert torch.equal(pending.wait(), out) and pending.wait() is pending.tensor
    q.put((rank, start, count, out.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total_envs", [16, 1001])
def test_two_rank_sharding_and_stat_allreduce(total_envs):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, c0, v0), (r1, s1, c1, v1) = results
    assert s0 == 0 and s0 + c0 == s1 and c0 + c1 == total_envs
    assert v0 == v1  # every rank holds the global sums
    ids = torch.arange(total_envs, dtype=torch.float64)
    want = [float(total_envs), ids.sum().item(), 1000.0 * total_envs, (ids % 7).sum().item(),
            (ids % 3).sum().item(), (ids % 5).sum().item(), 0.0, float(total_envs)]
    assert v0 == want


def test_all_reduce_is_identity_without_process_group():
    from rl_env_b200 import all_reduce_stats
    v = torch.arange(8, dtype=torch.float64)
    assert all_reduce_stats(v) is v
    assert all_reduce_stats(v, async_op=True).wait() is v
