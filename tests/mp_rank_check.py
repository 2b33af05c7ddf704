"""Worker of tests/test_gpu_multirank.py: one rank of a torchrun job (gloo rendezvous), all ranks on the GPUs available
(rank % device_count, so two ranks can share one GPU).  Every rank steps ITS shard of the global env range on the CUDA
path, checks it bit-for-bit against C oracles keyed by the same global ids, runs the per-step statistics exchange in the
requested form ('p2p' peer-memory pushes fused into the step kernel, or 'nccl'), and compares the global sums it
delivers with a gloo all-reduce of the local ones.  Exit code 0 = all good."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import c_oracle
import snakes_b200


def main():
    mode = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    total, T = 4096 * world, 24
    base, count = snakes_b200.shard_range(total, rank, world)
    kw = dict(size=19, n_snakes=2, rules="classic", seed=11)
    env = snakes_b200.SnakeVecEnv(count, device=dev, env_id_base=base, **kw)
    co = c_oracle.COracle(count, env_id_base=base, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    assert env.init_comm(mode=mode) == world
    # eager steps, every one against the oracle
    for t in range(8):
        a = env.gen_actions(t, 3)
        obs, rew, done, _ = env.step(a)
        cobs, crew, cdone, _ = co.step(a.cpu().numpy())
        assert np.array_equal(obs.cpu().numpy(), cobs) and np.array_equal(rew.cpu().numpy(), crew) and np.array_equal(done.cpu().numpy(), cdone), t
    # then T steps as one graph launch (the bench's step loop), checked at the end
    acts = torch.stack([env.gen_actions(8 + t, 3).clone() for t in range(T)])
    g = env.make_graph(acts)
    g.launch()
    an = acts.cpu().numpy()
    for t in range(T):
        cobs, _, _, _ = co.step(an[t])
    assert np.array_equal(env.obs.cpu().numpy(), cobs)
    dev_state, cpu_state = env.dump_state(), co.state()
    for k in cpu_state:
        assert np.array_equal(dev_state[k], cpu_state[k]), k
    torch.cuda.synchronize()
    dist.barrier()   # every rank's last push / all-reduce has landed
    local = env.stats(reduce=False)
    want = torch.tensor([local[k] for k in snakes_b200._lib.STAT_NAMES], dtype=torch.float64)
    dist.all_reduce(want)
    got = env.stats_global()
    assert [got[k] for k in snakes_b200._lib.STAT_NAMES] == want.tolist(), (rank, got, want.tolist())
    assert got["env_steps"] == float(total) * (8 + T)
    info = env.comm_info()
    assert info["ranks"] == world and info["allreduces"] == 8 + T, info
    dist.barrier()   # nobody tears its inbox down while a peer may still read / write it
    g.close(); env.close()
    dist.destroy_process_group()
    print("rank %d/%d ok (%s, device %d)" % (rank, world, mode, dev), flush=True)


if __name__ == "__main__":
    main()
