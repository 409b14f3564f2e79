"""Synthetic workloads of the named shape: P candidate paths x T waypoints.

target[t] = FK(q*(t)) of a smooth reference joint path q* (sum of three sinusoids per joint inside the limits,
numpy seed `seed`); seeds x0 = clamp(q* + N(0, noise^2)) per path from torch.Generator(1234 + shard)."""
import numpy as np
import torch

from .data_types import Problem, DEFAULT_CONSTRAINTS
from .robot import Robot

# problems/fetch__circle.yaml:13-16 (x, y, z, size_x, size_y, size_z)
FETCH_CIRCLE_OBSTACLES = [
    (0.4, 0.4, 0.825, 0.3, 0.05, 0.8), (0.4, -0.4, 0.825, 0.3, 0.05, 0.8),
    (0.4, 0.0, 1.225, 0.3, 0.85, 0.05), (0.4, 0.0, 0.425, 0.3, 0.85, 0.05),
]


def smooth_joint_path(limits, T: int, seed: int, amp: float = 0.25) -> np.ndarray:
    g = np.random.default_rng(seed)
    lim = np.array(limits)
    ndof = lim.shape[0]
    mid, half = lim.mean(1), (lim[:, 1] - lim[:, 0]) / 2
    t = np.linspace(0, 1, T)[:, None]
    q = mid + half * 0.3 * g.uniform(-1, 1, (1, ndof))
    for _ in range(3):
        q = q + half * amp * g.uniform(0.2, 1.0, (1, ndof)) * np.sin(
            2 * np.pi * (g.uniform(0.3, 1.5, (1, ndof)) * t + g.uniform(0, 1, (1, ndof))))
    return np.clip(q, lim[:, 0] + 0.05 * half, lim[:, 1] - 0.05 * half)


def cuboid_tensors(obstacles):
    cuboids, Tcuboids = [], []
    for (ox, oy, oz, sx, sy, sz) in obstacles:
        cuboids.append(torch.tensor([-sx / 2, -sy / 2, -sz / 2, sx / 2, sy / 2, sz / 2]))
        Tc = torch.zeros((4, 4))
        Tc[:3, :3] = torch.eye(3)
        Tc[0, 3], Tc[1, 3], Tc[2, 3] = ox, oy, oz
        Tcuboids.append(Tc)
    return cuboids, Tcuboids


# A second workload of the same shape whose reference joint path stays clear of the four cuboids (the default path,
# seed 0 / amplitude 0.25, drives the upper arm THROUGH a cuboid: no path of that batch can ever be valid, which is
# fine for timing the kernels but leaves the cost argmin nothing to choose from).
FEASIBLE = dict(seed=3, amp=0.08)


def synthetic_seeds_host(robot: Robot, P: int, T: int, seed: int = 0, noise: float = 0.05, shard: int = 0,
                         pin: bool = False, amp: float = 0.25):
    """-> (q* [T, D] float32 host, x0 [P*T, D] float32 host (optionally pinned))."""
    qstar = torch.tensor(smooth_joint_path(robot.actuated_joints_limits, T, seed, amp), dtype=torch.float32)
    g = torch.Generator().manual_seed(1234 + shard)
    x0 = qstar[None] + noise * torch.randn((P, T, robot.ndof), generator=g)
    lim = torch.tensor(robot.actuated_joints_limits, dtype=torch.float32)
    x0 = torch.minimum(torch.maximum(x0, lim[:, 0]), lim[:, 1]).reshape(P * T, robot.ndof).contiguous()
    if pin:
        x0 = x0.pin_memory()
    return qstar, x0


def synthetic_problem(robot: Robot, T: int, seed: int = 0, obstacles=FETCH_CIRCLE_OBSTACLES, device="cuda:0",
                      amp: float = 0.25) -> Problem:
    qstar = torch.tensor(smooth_joint_path(robot.actuated_joints_limits, T, seed, amp), dtype=torch.float32)
    target = robot.forward_kinematics(qstar.to(device))
    cuboids, Tcuboids = cuboid_tensors(obstacles)
    return Problem(DEFAULT_CONSTRAINTS, target, None, robot, "synthetic", f"{robot.name}__synthetic", list(obstacles),
                   [t.to(device) for t in Tcuboids], [c.to(device) for c in cuboids], [])
