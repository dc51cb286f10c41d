"""Multi-GPU plumbing: the physics needs NO data-path collective — environments are independent, so each rank
owns a contiguous range of GLOBAL env ids (Philox streams are keyed by global id, so any sharding reproduces
the same per-env trajectories).  The only exchange is a sum all-reduce of the 8-double episode-statistics
vector once per iteration (the reference concatenates worker results on the host instead,
environment/controller/ppo.py:371-382).  One process per GPU, torch.distributed for the plumbing."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

STAT_NAMES = ["sum_return", "sum_length", "n_episodes", "n_solved", "n_broken", "n_timeout", "sum_effort", "n_steps"]


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous split of n_total global env ids; returns (n_local, env_id_offset).  The first
    n_total % world ranks get one extra env."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_total), world)
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return n_local, offset


def init_distributed(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).
    Returns (rank, world, local_rank).  Single-process runs return (0, 1, 0) without initialising anything."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def allreduce_stats(stats: torch.Tensor, async_op: bool = False):
    """In-place sum all-reduce of an (8,) float64 statistics vector over all ranks (no-op for world size 1)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(stats, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


def stats_dict(stats: torch.Tensor) -> dict:
    vals = stats.detach().cpu().tolist()
    out = dict(zip(STAT_NAMES, vals))
    ne = max(out["n_episodes"], 1.0)
    out["mean_return"] = out["sum_return"] / ne
    out["mean_length"] = out["sum_length"] / ne
    out["solved_frac"] = out["n_solved"] / ne
    return out
