"""Multi-GPU plumbing: one process per GPU, envs sharded ``i -> rank i mod G``.

Every (env, ear) is independent, so the data path needs NO collective: in DD-PPO each rank's
policy consumes its own shard -- as in the reference, whose collectives only touch gradients and
statistics (ss_baselines/av_nav/ddppo/ddppo.py:35-39, ddppo_trainer.py:315-324).
``gather_observations`` is the one optional collective: an all-gather that reassembles the
observation batch for a centralised policy (NCCL on CUDA tensors; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .planning import shard_envs  # noqa: F401


def gather_observations(local: torch.Tensor, n_envs: int, rank: int, world: int) -> torch.Tensor:
    """All-gather the per-rank rows (rank r holds envs r, r+G, r+2G, ...) into the
    (n_envs, ...) batch in env order.  Ragged shards are padded to the largest shard."""
    if world == 1:
        return local
    per = -(-n_envs // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    # rank r, slot k  ->  env k*world + r
    out = out.view(world, per, *local.shape[1:]).transpose(0, 1).reshape(world * per, *local.shape[1:])
    return out[:n_envs]
