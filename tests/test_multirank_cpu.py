"""World-size-2 gloo test (CPU) of the N>1 host logic: shard bookkeeping, per-rank seeds, episode-statistics
reduction and the max-over-ranks timing convention. The env data path itself needs no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isaacgymdyros_b200.sharding import Shard, max_over_ranks, reduce_episode_stats, shard_from_env


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = shard_from_env(8)
    assert (sh.rank, sh.world, sh.first_env, sh.global_envs) == (rank, world, rank * 8, 16)
    g = torch.Generator().manual_seed(sh.seed(42))
    epi = torch.arange(8, dtype=torch.float32) + 100 * rank
    crm = torch.full((8,), 0.1 * (rank + 1))
    crm[0] = float("nan")  # an env that never finished an episode (0/0, T:654)
    stats = reduce_episode_stats({"epi_len_log": epi, "contact_reward_mean": crm})
    t = max_over_ranks(1.0 + rank, "cpu")
    out[rank] = (stats, t, int(torch.randint(0, 1 << 30, (1,), generator=g)))
    dist.destroy_process_group()


def test_world_size_two_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    s0, t0, r0 = out[0]
    s1, t1, r1 = out[1]
    assert s0 == s1 and t0 == t1 == 2.0 and r0 != r1
    want_epi = (sum(range(8)) + sum(range(100, 108))) / 16
    assert abs(s0["epi_len_log"] - want_epi) < 1e-9
    assert abs(s0["contact_reward_mean"] - (7 * 0.1 + 7 * 0.2) / 16) < 1e-6


def test_single_process_paths():
    sh = Shard(rank=0, world=1, local_rank=0, envs_per_rank=4096)
    assert sh.device() == "cuda:0" and sh.global_envs == 4096 and sh.seed(42) == 42
    assert max_over_ranks(3.5, "cpu") == 3.5
    assert reduce_episode_stats({"x": torch.tensor([1.0, 3.0])}) == {"x": 2.0}
