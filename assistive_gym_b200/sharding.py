"""Multi-GPU layout of the env batch: envs are independent, so each rank owns a
contiguous block of N/G envs; nothing is exchanged inside the physics step.  The one collective of the
path is an all-gather of the per-env reward tensor, issued only when a global view is requested
(the reference's data-parallel mode is one process per core with no exchange at all, learn.py:26)."""
import numpy as np


def shard_range(rank, world, n_global):
    """Contiguous env block [lo, hi) of `rank`."""
    assert n_global % world == 0, 'global batch must be divisible by the number of ranks'
    per = n_global // world
    return rank * per, (rank + 1) * per


def env_rng(global_env_id, base_seed=1001):
    """Per-env generator derived from the GLOBAL env id, so results do not depend on the partition
    (reference default seed 1001, envs/env.py:21)."""
    return np.random.default_rng([base_seed, int(global_env_id)])


def sample_block(fb, lo, hi, base_seed=1001):
    """Reset-time randomisation for envs [lo, hi): every env draws from its own generator."""
    rows = [fb.sample(1, env_rng(g, base_seed)) for g in range(lo, hi)]
    return {k: np.concatenate([r[k] for r in rows], axis=0) for k in rows[0]}


def all_gather_rewards(reward, group=None):
    """reward: torch tensor [N/G] on this rank -> [N] on every rank (NCCL on GPUs, gloo in tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty(world * reward.numel(), dtype=reward.dtype, device=reward.device)
    dist.all_gather_into_tensor(out, reward.contiguous(), group=group)
    return out
