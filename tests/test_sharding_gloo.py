"""CPU, world_size 2 over gloo: the N>1 host logic -- contiguous sharding by global env id and the
statistics all-reduce -- checked with the CPU oracle standing in for the per-rank device step."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, STEPS = 101, 60
KW = dict(size=10, n_snakes=2, rules="classic", seed=17)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import c_oracle
    from snakes_b200.sharding import all_reduce_stats, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, count = shard_range(TOTAL, rank, world)
    co = c_oracle.COracle(count, env_id_base=base, **KW)
    co.reset()
    rewards = []
    for t in range(STEPS):
        _, r, _, _ = co.step(c_oracle.gen_actions(co.cfg, t, 3))
        rewards.append(r.copy())
    stats = all_reduce_stats(torch.tensor(co.stats(), dtype=torch.float64))
    gathered = [None] * world
    dist.all_gather_object(gathered, (base, count, np.stack(rewards, 1)))
    if rank == 0:
        out.put((stats.numpy(), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    import c_oracle
    from snakes_b200.sharding import shard_range
    assert [shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 3), (7, 3)]
    assert sum(shard_range(1048576, r, 8)[1] for r in range(8)) == 1048576 and shard_range(1048576, 3, 8) == (393216, 131072)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    stats, gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = c_oracle.COracle(TOTAL, env_id_base=0, **KW)
    whole.reset()
    rewards = []
    for t in range(STEPS):
        _, r, _, _ = whole.step(c_oracle.gen_actions(whole.cfg, t, 3))
        rewards.append(r.copy())
    rewards = np.stack(rewards, 1)
    assert np.allclose(stats, whole.stats())
    covered = 0
    for base, count, rew in sorted(gathered):
        assert base == covered
        assert np.array_equal(rew, rewards[base:base + count])
        covered += count
    assert covered == TOTAL


def _learner_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from snakes_b200 import selfplay
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # every rank initialises ITS network differently ...
    model = selfplay.Model((12, 12, 3), 5, device="cpu")
    before = torch.cat([p.detach().flatten() for p in model.net.parameters()]).clone()
    selfplay.sync_model_across_ranks(model)             # ... and must start from rank 0's weights
    g = torch.Generator().manual_seed(7 + rank)         # different data per rank (sharded envs)
    for _ in range(3):
        n = 64
        obs = torch.randint(0, 256, (n, 12, 12, 3), dtype=torch.uint8, generator=g)
        ret, val, nlp = torch.randn(n, generator=g), torch.randn(n, generator=g), torch.rand(n, generator=g) + 1.0
        act = torch.randint(0, 5, (n,), generator=g)
        model.train(2.5e-4, 0.2, obs, ret, None, act, val, nlp)
    after = torch.cat([p.detach().flatten() for p in model.net.parameters()])
    gathered = [None] * world
    dist.all_gather_object(gathered, (before.numpy(), after.numpy()))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_learner_replicas_stay_identical():
    """the data-parallel learner averaged gradients but never synchronised the initial weights.  Two
    ranks with different seeds and different data: after sync_model_across_ranks + 3 updates the replicas are equal."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_learner_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (b0, a0), (b1, a1) = gathered
    assert not np.array_equal(b0, b1)          # they did start different
    assert np.array_equal(a0, a1)              # and ended identical, bit for bit
    assert not np.array_equal(a0, b0)          # after real updates
