"""Env sharding across ranks (one process per GPU; SURVEY section 8e, reference: utils/rlgames_utils.py:71-81 maps
rank -> cuda:{rank} and builds an independent env of numEnvs there). Envs are independent, so the data path needs no
collective; the only reductions are the episode statistics / curriculum-gate means and (in the caller) PPO gradients."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.distributed as dist


@dataclass
class Shard:
    rank: int
    world: int
    local_rank: int
    envs_per_rank: int

    @property
    def global_envs(self) -> int:
        return self.envs_per_rank * self.world

    @property
    def first_env(self) -> int:
        """Global index of this rank's first env: rank r owns [r * n, (r + 1) * n)."""
        return self.rank * self.envs_per_rank

    def device(self) -> str:
        return f"cuda:{self.local_rank}"

    def seed(self, base: int) -> int:
        """Per-rank RNG key (reference seeds every rank identically, utils/utils.py:43-68; distinct keys avoid
        replaying the same randomisation on every GPU)."""
        return base + self.rank


def shard_from_env(envs_per_rank: int) -> Shard:
    return Shard(rank=int(os.environ.get("RANK", "0")), world=int(os.environ.get("WORLD_SIZE", "1")),
                 local_rank=int(os.environ.get("LOCAL_RANK", "0")), envs_per_rank=envs_per_rank)


def reduce_episode_stats(local: Dict[str, torch.Tensor], group: Optional[dist.ProcessGroup] = None) -> Dict[str, float]:
    """All-reduce of per-rank sums into global means. `local` maps a name to a 1-D per-env tensor; NaNs (envs that
    never finished an episode: 0/0 in T:654) count as zero, as torch.mean would poison the reference's gate otherwise.
    One small collective (NCCL on GPUs, gloo in the CPU tests)."""
    names = sorted(local)
    sums = torch.stack([local[n].double().nan_to_num().sum() for n in names] +
                       [torch.tensor(float(local[names[0]].numel()), dtype=torch.float64, device=local[names[0]].device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    n = float(sums[-1].item())
    return {k: float(sums[i].item()) / n for i, k in enumerate(names)}


def max_over_ranks(x: float, device: str, group: Optional[dist.ProcessGroup] = None) -> float:
    """Timing convention of bench.py: every multi-GPU number is the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
