"""Replicas across GPUs: the only multi-device axis of this path (SURVEY §8e — "replicas only").

GPU g of G owns the seeded solver replicas {r : r mod G == g}; instance data is replicated, no
candidate/state/aggregate ever crosses devices. The single exchange is a best-score sync: each
rank contributes its best committed score packed into an order-preserving int64 key, reduced with
MAX (NCCL on GPUs, gloo in the CPU tests); the owner rank is recovered with a second tiny MIN.

Reference analogue: up to 16 concurrent jobs of one SolverManager, each with its own
`SolverConfig::random_seed` (solverforge-solver/src/manager/solver_manager/manager.rs:22,93-146).
"""
from __future__ import annotations

HARD_BIAS = 1 << 22
SOFT_BIAS = 1 << 39


def partition_replicas(total: int, world: int, rank: int) -> list[int]:
    """Replica ids owned by `rank` (round-robin so seeds 1000+r spread evenly)."""
    return [r for r in range(total) if r % world == rank]


def replica_seed(replica: int, base: int = 1000) -> int:
    return base + replica


def key_saturates(hard: int, soft: int) -> bool:
    """True when (hard, soft) lies outside the packed key's field ranges: distinct scores would then collapse
    onto one key. sync_best_scores (and the C ABI's sfgpu_sync_best) have no such limit."""
    return not (-HARD_BIAS <= hard < HARD_BIAS and -SOFT_BIAS <= soft < SOFT_BIAS)


def pack_score_key(hard: int, soft: int) -> int:
    """((hard + 2^22) << 40) | (soft + 2^39); levels saturate at the field range (same as the device
    kernel behind sfgpu_pack_best_keys) — check key_saturates() or use sync_best_scores when scores may leave
    that range. Ordering of keys == lexicographic (hard, soft) ordering."""
    h = min(max(hard, -HARD_BIAS), HARD_BIAS - 1) + HARD_BIAS
    s = min(max(soft, -SOFT_BIAS), SOFT_BIAS - 1) + SOFT_BIAS
    return (h << 40) | s


def unpack_score_key(key: int) -> tuple[int, int]:
    return (key >> 40) - HARD_BIAS, (key & ((1 << 40) - 1)) - SOFT_BIAS


def sync_best(local_best_key, group=None):
    """All ranks learn (global best key, owner rank). `local_best_key` is a 1-element int64 tensor on
    the backend's device (cuda for NCCL, cpu for gloo). Two collectives of 8 bytes each."""
    import torch
    import torch.distributed as dist
    best = local_best_key.clone()
    dist.all_reduce(best, op=dist.ReduceOp.MAX, group=group)
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    mine = torch.where(local_best_key == best, torch.full_like(best, rank), torch.full_like(best, world))
    dist.all_reduce(mine, op=dist.ReduceOp.MIN, group=group)
    return int(best.item()), int(mine.item())


def sync_best_scores(scores, group=None):
    """Exact best-score sync over (hard, soft) pairs: `scores` is an int64 tensor [R, 2] (or [2]) on the backend's
    device holding this rank's replica scores. Every rank contributes its lexicographic best with one all_gather of
    16 bytes per rank and reduces locally — no packed key, so no saturation. Returns (hard, soft, owner_rank); ties
    go to the lowest rank. Same contract as sfgpu_sync_best of the C ABI."""
    import torch
    import torch.distributed as dist
    s = scores.reshape(-1, 2)
    # lexicographic max on the device: best hard level, then the best soft level among its holders
    h = s[:, 0].max()
    soft = torch.where(s[:, 0] == h, s[:, 1], torch.full_like(s[:, 1], torch.iinfo(torch.int64).min)).max()
    mine = torch.stack([h, soft])
    if not (dist.is_available() and dist.is_initialized()):
        v = mine.cpu()
        return int(v[0]), int(v[1]), 0
    world = dist.get_world_size(group)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    allv = torch.stack(out).cpu().tolist()
    best = max(range(world), key=lambda g: (allv[g][0], allv[g][1], -g))
    return int(allv[best][0]), int(allv[best][1]), best
