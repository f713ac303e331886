"""Data-parallel plumbing of the policy update: trajectories are sharded along N,
parameters replicated, ONE all-reduce of the flat trainable gradient per optimiser step
(inside ``ILTrainerStep.update_agent``).  The reference's IL trainers are single-GPU; its
DD-PPO trainer averages gradients through a DDP reducer (``ddppo_waypoint_trainer.py:370``)."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl"):
    """One process per GPU under torchrun: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def shard_rows(x: torch.Tensor, t: int, n: int, rank: int, world: int) -> torch.Tensor:
    """Slice the trajectory axis of a time-major ``(T*N, ...)`` tensor: all T rows of a
    trajectory stay on one rank (the GRU recurrence runs along T)."""
    assert n % world == 0, "trajectories must divide evenly across ranks"
    per = n // world
    v = x.view(t, n, *x.shape[1:])[:, rank * per:(rank + 1) * per]
    return v.reshape(t * per, *x.shape[1:]).contiguous()


def shard_batch(batch: Tuple, rank: int, world: int):
    """``collate_fn``-layout batch (obs, prev_actions (T*N,1), masks (T*N,1), actions (T,N),
    weights (T,N)) -> this rank's shard."""
    obs, prev, masks, actions, weights = batch
    t, n = actions.shape
    per = n // world
    o: Dict[str, torch.Tensor] = {k: shard_rows(v, t, n, rank, world) for k, v in obs.items()}
    sl = slice(rank * per, (rank + 1) * per)
    return (o, shard_rows(prev, t, n, rank, world), shard_rows(masks, t, n, rank, world),
            actions[:, sl].contiguous(), weights[:, sl].contiguous())
