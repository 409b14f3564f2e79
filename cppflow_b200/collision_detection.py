"""Batched capsule collision checks with the reference's names and shapes (cppflow/collision_detection.py:9-86).
One fused CUDA kernel evaluates all-link FK, every capsule pair / capsule-cuboid distance and the `min < 0` test;
nothing of size [k*T, S] is materialised.  The klampt mesh variants (:89-120) are out of scope."""
from typing import List

import torch

from . import ops
from .data_types import Problem


def get_only_non_colliding_qpaths(qpaths: List[torch.Tensor], self_colliding: torch.Tensor, env_colliding: torch.Tensor):
    """collision_detection.py:9-24"""
    assert len(qpaths) == self_colliding.shape[0] == env_colliding.shape[0]
    colliding_idxs = torch.logical_or(self_colliding, env_colliding)
    to_keep = torch.sum(colliding_idxs, dim=1) == 0
    return [qpaths[i] for i in to_keep.nonzero()[:, 0]]


def qpaths_batched_env_collisions(problem: Problem, q: torch.Tensor) -> torch.Tensor:
    """[k, T, ndof] -> bool [k, T]: config collides with any cuboid obstacle (collision_detection.py:27-49)."""
    k, n, ndof = q.shape
    robot = problem.robot
    if problem.obstacle_tables.n == 0:
        return torch.zeros((k, n), dtype=torch.bool, device=q.device)
    _, e = ops.collision_flags(robot.robot_id, robot.ndof, q.reshape((k * n, ndof)), problem.obstacle_tables,
                               want_self=False, want_env=True)
    assert e.numel() == n * k
    return e.bool().reshape((k, n))


def qpaths_batched_self_collisions(problem: Problem, q: torch.Tensor) -> torch.Tensor:
    """[k, T, ndof] -> bool [k, T]: any capsule pair overlaps (collision_detection.py:52-69)."""
    k, n, ndof = q.shape
    robot = problem.robot
    s, _ = ops.collision_flags(robot.robot_id, robot.ndof, q.reshape((k * n, ndof)), None, want_self=True, want_env=False)
    return s.bool().reshape((k, n))


def qpaths_batched_collisions(problem: Problem, q: torch.Tensor):
    """Both flag sets from ONE launch (the planner needs both, planners.py:235,245)."""
    k, n, ndof = q.shape
    robot = problem.robot
    s, e = ops.collision_flags(robot.robot_id, robot.ndof, q.reshape((k * n, ndof)), problem.obstacle_tables)
    return s.bool().reshape((k, n)), e.bool().reshape((k, n))


def self_colliding_configs_capsule(problem: Problem, qpath: torch.Tensor) -> torch.Tensor:
    """collision_detection.py:72-74"""
    return qpaths_batched_self_collisions(problem, qpath[None])[0]


def env_colliding_configs_capsule(problem: Problem, qpath: torch.Tensor) -> torch.Tensor:
    """collision_detection.py:77-86"""
    return qpaths_batched_env_collisions(problem, qpath[None])[0]


def env_colliding_links_capsule(problem: Problem, q: torch.Tensor) -> List[str]:
    """collision_detection.py:134-143"""
    links = []
    ordered_links = list(problem.robot._collision_capsules_by_link.keys())
    for cuboid, Tcuboid in zip(problem.obstacles_cuboids, problem.obstacles_Tcuboids):
        dists = problem.robot.env_collision_distances(q.unsqueeze(0), cuboid, Tcuboid)
        for i in range(dists.shape[1]):
            if dists[0, i] < 0:
                links.append(ordered_links[i])
    return list(set(links))
