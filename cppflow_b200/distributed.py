"""Multi-GPU plumbing: candidate paths (and planning problems) are independent, so they are sharded across ranks with
no collective in the data path; NCCL is used once, after refinement, to gather every rank's best path and pick the
argmin (BASELINE.json north_star).  The reference has no multi-GPU code at all.

Ranking key.  A path is ranked by (invalid, trajectory length TL, global path index), lexicographically: valid paths
before invalid ones, shorter before longer, lowest index on ties - so the answer does not depend on the shard count.
The three fields are packed into ONE non-negative int64

    key = invalid << 62  |  bits(float32 TL) << 31  |  global index          (TL >= 0: its bit pattern orders like the value)

and every rank reduces its shard to a single key on the device.  The collective is then ONE all-gather of three int64
per rank (key, number of valid paths, first global index of the shard): no size exchange, no host synchronisation
before the result is wanted.  (Round 1 ranked by the float32 sum TL + 1e9: at 1e9 one ulp is 64, so among invalid
paths the trajectory length vanished and the winner was a rounding artefact.)"""
from typing import NamedTuple, Tuple

import torch

INVALID_COST = 1.0e9
_IDX_BITS = 31
_TL_SHIFT = _IDX_BITS
_INVALID_SHIFT = 62


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_items over `world` ranks (first n_items % world ranks get one extra)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def path_validity(metrics: torch.Tensor, constraints) -> torch.Tensor:
    """bool [P]: the thresholds of x_is_valid (evaluation_utils.py:41-58, strict '<') and no capsule collision."""
    return (
        (metrics[:, 0] < constraints.max_allowed_position_error_cm)
        & (metrics[:, 1] < constraints.max_allowed_rotation_error_deg)
        & (metrics[:, 2] < constraints.max_allowed_mjac_deg)
        & (metrics[:, 3] < constraints.max_allowed_mjac_cm)
        & (metrics[:, 5] >= 0)
        & (metrics[:, 6] >= 0)
    )


def path_costs(metrics: torch.Tensor, constraints) -> torch.Tensor:
    """Per-path scalar cost (float64 [P]) from the [P, 8] metrics of ops.path_metrics: the trajectory length TL
    (optimization.py:173-175) for valid paths, TL + 1e9 for the others.  For reporting; the argmin uses path_keys."""
    tl = metrics[:, 4].double()
    return torch.where(path_validity(metrics, constraints), tl, tl + INVALID_COST)


def path_keys(metrics: torch.Tensor, constraints, first_index: int = 0) -> torch.Tensor:
    """int64 [P] ranking keys (module docstring).  A NaN trajectory length counts as invalid and sorts last."""
    tl = metrics[:, 4].float().contiguous()
    finite = torch.isfinite(tl) & (tl >= 0)
    invalid = ~(path_validity(metrics, constraints) & finite)
    tl_bits = torch.where(finite, tl, torch.full_like(tl, float("inf"))).view(torch.int32).to(torch.int64)
    idx = torch.arange(first_index, first_index + tl.numel(), device=tl.device, dtype=torch.int64)
    return (invalid.to(torch.int64) << _INVALID_SHIFT) | (tl_bits << _TL_SHIFT) | idx


class Best(NamedTuple):
    cost: float        # TL of the best path, + 1e9 if it is invalid
    rank: int          # owning rank
    index: int         # index within that rank's shard
    n_valid: int       # valid paths over all ranks
    valid: bool
    trajectory_length: float
    global_index: int


def decode_key(key: int):
    """-> (valid, trajectory length, global index)"""
    import struct

    invalid = (key >> _INVALID_SHIFT) & 1
    tl = struct.unpack("<f", struct.pack("<I", (key >> _TL_SHIFT) & 0x7FFFFFFF))[0]
    return (not invalid), tl, key & ((1 << _IDX_BITS) - 1)


class PendingArgmin:
    """Device-side state of an argmin in flight: nothing has been synchronised with the host yet."""

    def __init__(self, gathered: torch.Tensor, world: int):
        self.gathered, self.world = gathered, world  # int64 [world, 3]: key, n_valid, first index

    def result(self) -> Best:
        g = self.gathered.cpu().tolist()  # the one host synchronisation
        best_key = min(row[0] for row in g)
        valid, tl, gidx = decode_key(best_key)
        owner = max(r for r in range(self.world) if g[r][2] <= gidx)
        return Best(tl if valid else tl + INVALID_COST, owner, gidx - g[owner][2], sum(row[1] for row in g), valid, tl, gidx)


def enqueue_argmin(metrics: torch.Tensor, constraints, first_index: int, world: int) -> PendingArgmin:
    """Reduce this rank's [P, 8] metrics to (best key, #valid, first index) on the device and all-gather the three
    int64 of every rank - all enqueued on the current stream / the process group's stream, no host synchronisation."""
    assert first_index + metrics.shape[0] < (1 << _IDX_BITS), "global path index must fit 31 bits"
    dev = metrics.device
    if metrics.shape[0] > 0 and metrics.is_cuda:
        from . import ops  # one launch of the library's key + argmin kernel (the torch arithmetic below: ~25 launches)

        local = ops.path_key_argmin(metrics.contiguous(), constraints, first_index)
    elif metrics.shape[0] > 0:
        keys = path_keys(metrics, constraints, first_index)
        local = torch.stack([keys.min(), (keys >> _INVALID_SHIFT == 0).sum(),
                             torch.tensor(first_index, device=dev, dtype=torch.int64)])
    else:  # an empty shard never wins
        local = torch.tensor([(1 << 63) - 1, 0, first_index], device=dev, dtype=torch.int64)
    if world == 1:
        return PendingArgmin(local.reshape(1, 3), 1)
    import torch.distributed as dist

    gathered = torch.empty((world, 3), device=dev, dtype=torch.int64)
    dist.all_gather_into_tensor(gathered.reshape(-1), local)  # 24 B per rank: pure latency over NVLink
    return PendingArgmin(gathered, world)


def gather_costs_and_argmin(metrics: torch.Tensor, constraints, rank: int, world: int, first_index=None) -> Tuple[float, int, int]:
    """-> (best cost, owning rank, index within that rank's shard); ties resolve to the lowest global index.
    `first_index`: global index of this rank's first path (default: equal shards, rank * P)."""
    if first_index is None:
        first_index = rank * metrics.shape[0]
    b = enqueue_argmin(metrics, constraints, first_index, world).result()
    return b.cost, b.rank, b.index
