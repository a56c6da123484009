"""Multi-GPU plumbing: one process per GPU, envs sharded by rank, one gradient exchange per optimizer step.

The reference's CleanRL path is single-process; its rl_games / skrl front-ends shard envs
across processes and all-reduce the policy gradient (`U/skrl/ppo.py:126-131,534-537`).  This module is the
equivalent for the B200-native trainer: `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests)
carries exactly two kinds of traffic -- the initial weight broadcast and a sum-allreduce of the flat fp32
gradient bucket (377 241 floats = 1.509 MB) per optimizer step; the 1/world factor is folded into the
clip + Adam kernels (`catb200_adam_step(grad_scale=...)`), so every rank clips with the same global norm.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise the default process group from torchrun's env vars. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local_rank


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def broadcast_params(flat: torch.Tensor, src: int = 0) -> None:
    """Rank `src` seeds every rank's flat parameter vector (done once, before the first rollout)."""
    if world_size() > 1:
        dist.broadcast(flat, src=src)


def allreduce_grads(flat_grads: torch.Tensor) -> float:
    """Sum-allreduce the flat gradient bucket in place; returns the scale (1/world) the optimizer applies."""
    w = world_size()
    if w > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / w


def shard_seed(base_seed: int) -> int:
    """Per-rank seed, like the reference's distributed front-ends (`scripts/skrl/train.py:116-117`)."""
    return base_seed + rank()
