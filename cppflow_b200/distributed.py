"""Multi-GPU plumbing: candidate paths (and planning problems) are independent, so they are sharded across ranks with
no collective in the data path; NCCL is used once, after refinement, to gather the per-path costs and pick the argmin
(BASELINE.json north_star).  The reference has no multi-GPU code at all."""
from typing import Tuple

import torch

INVALID_COST = 1.0e9


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_items over `world` ranks (first n_items % world ranks get one extra)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def path_costs(metrics: torch.Tensor, constraints) -> torch.Tensor:
    """Per-path scalar cost from the [P, 8] metrics of ops.path_metrics: the trajectory length TL
    (optimization.py:173-175) for valid paths, TL + 1e9 for paths that break a threshold of x_is_valid
    (evaluation_utils.py:41-58) or whose capsules collide."""
    valid = (
        (metrics[:, 0] < constraints.max_allowed_position_error_cm)
        & (metrics[:, 1] < constraints.max_allowed_rotation_error_deg)
        & (metrics[:, 2] < constraints.max_allowed_mjac_deg)
        & (metrics[:, 3] < constraints.max_allowed_mjac_cm)
        & (metrics[:, 5] >= 0)
        & (metrics[:, 6] >= 0)
    )
    return torch.where(valid, metrics[:, 4], metrics[:, 4] + INVALID_COST)


def gather_costs_and_argmin(metrics: torch.Tensor, constraints, rank: int, world: int) -> Tuple[float, int, int]:
    """All-gather the per-path costs of every rank and return (best cost, owning rank, index within that rank's shard).
    Ties resolve to the lowest (rank, index), so the answer does not depend on the shard count."""
    costs = path_costs(metrics, constraints).contiguous()
    if world == 1:
        idx = int(torch.argmin(costs))
        return float(costs[idx]), 0, idx
    import torch.distributed as dist

    n_local = torch.tensor([costs.numel()], device=costs.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    n_max = int(max(int(s) for s in sizes))
    padded = torch.full((n_max,), float("inf"), device=costs.device, dtype=costs.dtype)
    padded[: costs.numel()] = costs
    gathered = torch.empty((world, n_max), device=costs.device, dtype=costs.dtype)
    dist.all_gather(list(gathered.unbind(0)), padded)  # P fp32 per rank (32 KB at P = 8192): pure latency over NVLink
    flat = int(torch.argmin(gathered.reshape(-1)))  # first minimum = lowest (rank, index)
    r, i = divmod(flat, n_max)
    return float(gathered[r, i]), r, i
