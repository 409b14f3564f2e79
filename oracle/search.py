"""`dp_search` and friends (TEST INFRASTRUCTURE - see oracle/__init__.py).

Literal CPU-torch transcription of the reference, op for op, so that the int32 back-pointer table `memo`
is bit-exact:
  search.py:14-21    cost constants, joint-limit paddings (1.5 deg / 3 cm)
  search.py:25-52    joint_limit_almost_violations_3d
  search.py:100-125  _get_mjacs  (prismatic x5 BEFORE the wrap)
  search.py:128-173  dp_search   (bottleneck DP: min_j max(mjac, cost) + ext; first-index argmin; backtrack)
  search.py:55-97    dp_search_slow (second, loop-based implementation: differential check)
  collision_detection.py:27-69  batched capsule collision flags
"""
from typing import Tuple

import numpy as np
import torch

from .robots import RobotModel
from . import geometry as G

K_JLIM_COST = 100
K_COLLISION_COST = 1000
DEFAULT_JLIM_SAFETY_PADDING_REVOLUTE = np.deg2rad(1.5)
DEFAULT_JLIM_SAFETY_PADDING_PRISMATIC = 3 / 100.0  # cm_to_m(3)


def joint_limit_almost_violations_3d(model: RobotModel, qs: torch.Tensor,
                                     eps_revolute: float = DEFAULT_JLIM_SAFETY_PADDING_REVOLUTE,
                                     eps_prismatic: float = DEFAULT_JLIM_SAFETY_PADDING_PRISMATIC) -> torch.Tensor:
    assert len(qs.shape) == 3
    l_lim = torch.tensor([l for l, _ in model.actuated_joints_limits], dtype=qs.dtype)
    u_lim = torch.tensor([u for _, u in model.actuated_joints_limits], dtype=qs.dtype)
    l_lim[model.prismatic_joint_idxs] += eps_prismatic
    l_lim[model.revolute_joint_idxs] += eps_revolute
    u_lim[model.prismatic_joint_idxs] -= eps_prismatic
    u_lim[model.revolute_joint_idxs] -= eps_revolute
    return torch.logical_or((qs < l_lim).any(dim=2), (qs > u_lim).any(dim=2)).type(torch.float32)


def get_mjacs(q: torch.Tensor, model: RobotModel, prismatic_joint_scaling: float = 5.0) -> torch.Tensor:
    k, ntimesteps, ndof = q.shape
    dqs = q[:, 1:, :].unsqueeze(1).expand(k, k, ntimesteps - 1, ndof) - q[:, :-1, :].unsqueeze(0).expand(
        k, k, ntimesteps - 1, ndof
    )
    if model.has_prismatic_joints:
        dqs[:, :, :, model.prismatic_joint_idxs] *= prismatic_joint_scaling
    abs_dqs = torch.abs(torch.remainder(dqs + torch.pi, 2 * torch.pi) - torch.pi)
    mjacs, _ = torch.max(abs_dqs, 3)
    return mjacs


def dp_search(model: RobotModel, q: torch.Tensor, self_collision_violations: torch.Tensor,
              env_collision_violations: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (best_path [T,D], memo int32 [k,T], costs [k,T], chosen candidate index per timestep int64 [T])."""
    k, ntimesteps, ndof = q.shape
    jlimit_violations = joint_limit_almost_violations_3d(model, q)
    costs = torch.zeros((k, ntimesteps), dtype=q.dtype)
    q_costs_external = (
        K_JLIM_COST * jlimit_violations
        + K_COLLISION_COST * env_collision_violations
        + K_COLLISION_COST * self_collision_violations
    )
    costs[:, 0] = q_costs_external[:, 0]
    mjacs = get_mjacs(q, model)
    memo = torch.zeros((k, ntimesteps), dtype=torch.int32)
    for t in range(1, ntimesteps):
        t_next_cost = torch.maximum(mjacs[:, :, t - 1], costs[:, t - 1])
        t_next_cost += q_costs_external[:, t].unsqueeze(0).expand(k, k).transpose(0, 1)
        costs[:, t], memo[:, t] = torch.min(t_next_cost, 1)
    best_path = torch.zeros((ntimesteps, ndof), dtype=q.dtype)
    chosen = torch.zeros(ntimesteps, dtype=torch.int64)
    _, i = torch.min(costs[:, -1], 0)
    for t in range(ntimesteps - 1, -1, -1):
        best_path[t, :] = q[i, t, :]
        chosen[t] = i
        i = memo[i, t]
    return best_path, memo, costs, chosen


def dp_search_slow(model: RobotModel, q: torch.Tensor, self_collision_violations: torch.Tensor,
                   env_collision_violations: torch.Tensor) -> torch.Tensor:
    """search.py:55-97 - note: no prismatic x5 scaling in this variant."""
    k, ntimesteps, ndof = q.shape
    memo = torch.zeros((k, ntimesteps), dtype=torch.int32)
    jlimit_violations = joint_limit_almost_violations_3d(model, q)
    costs = torch.zeros((k, ntimesteps), dtype=q.dtype)
    q_costs_external = (
        K_JLIM_COST * jlimit_violations
        + K_COLLISION_COST * env_collision_violations
        + K_COLLISION_COST * self_collision_violations
    )
    costs[:, 0] = q_costs_external[:, 0]
    for t in range(1, ntimesteps):
        for ki in range(k):
            dqs = q[ki, t, :] - q[:, t - 1, :]
            absdqs = torch.abs(torch.remainder(dqs + torch.pi, 2 * torch.pi) - torch.pi)
            maxdqs, _ = torch.max(absdqs, 1)
            t_next_cost = torch.maximum(maxdqs, costs[:, t - 1])
            t_next_cost += q_costs_external[ki, t]
            costs[ki, t], memo[ki, t] = torch.min(t_next_cost, 0)
    best_path = torch.zeros((ntimesteps, ndof), dtype=q.dtype)
    _, i = torch.min(costs[:, -1], 0)
    for t in range(ntimesteps - 1, -1, -1):
        best_path[t, :] = q[i, t, :]
        i = memo[i, t]
    return best_path


def qpaths_batched_self_collisions(model: RobotModel, q: torch.Tensor) -> torch.Tensor:
    """collision_detection.py:52-69"""
    k, n, ndof = q.shape
    dists = G.self_collision_distances(model, q.reshape((k * n, ndof)))
    min_dists, _ = torch.min(dists, dim=1)
    return (min_dists < 0).reshape((k, n))


def qpaths_batched_env_collisions(model: RobotModel, q: torch.Tensor, cuboids, Tcuboids) -> torch.Tensor:
    """collision_detection.py:27-49"""
    k, n, ndof = q.shape
    colliding = torch.zeros((k, n), dtype=torch.bool)
    q_2d = q.reshape((k * n, ndof))
    for cuboid, Tcuboid in zip(cuboids, Tcuboids):
        dists = G.env_collision_distances(model, q_2d, cuboid, Tcuboid)
        min_dists, _ = torch.min(dists, dim=1)
        colliding = torch.logical_or(colliding, (min_dists < 0).reshape((k, n)))
    return colliding
