"""Synthetic inputs of the oracle side (same recipe as tests/golden/make_golden.py and cppflow_b200/synthetic.py).
TEST INFRASTRUCTURE - see oracle/__init__.py: imported by tests/ and by bench.py's CPU-baseline / reference legs only."""
import numpy as np
import torch

from . import robots as R, kinematics as K

FETCH_CIRCLE_OBSTACLES = [
    # problems/fetch__circle.yaml:13-16  (x, y, z, size_x, size_y, size_z)
    (0.4, 0.4, 0.825, 0.3, 0.05, 0.8), (0.4, -0.4, 0.825, 0.3, 0.05, 0.8),
    (0.4, 0.0, 1.225, 0.3, 0.85, 0.05), (0.4, 0.0, 0.425, 0.3, 0.85, 0.05),
]
PANDA_1CUBE_OBSTACLES = [(0.0, 0.2, 0.7, 0.25, 0.25, 0.25)]  # problems/panda__1cube.yaml
OBSTACLES = {"fetch": FETCH_CIRCLE_OBSTACLES, "fetch_arm": FETCH_CIRCLE_OBSTACLES[:2], "panda": PANDA_1CUBE_OBSTACLES}


def cuboid_tensors(obstacles, dtype=torch.float32):
    """(x,y,z,sx,sy,sz) -> (cuboids [6], Tcuboids [4,4]) exactly as data_type_utils.py:109-127 (Tcuboid[3,3] = 0)."""
    cuboids, Tcuboids = [], []
    for (ox, oy, oz, sx, sy, sz) in obstacles:
        cuboids.append(torch.tensor([-sx / 2, -sy / 2, -sz / 2, sx / 2, sy / 2, sz / 2], dtype=dtype))
        Tc = torch.zeros((4, 4), dtype=dtype)
        Tc[:3, :3] = torch.eye(3, dtype=dtype)
        Tc[0, 3], Tc[1, 3], Tc[2, 3] = ox, oy, oz
        Tcuboids.append(Tc)
    return cuboids, Tcuboids


def smooth_joint_path(model, T, seed, amp=0.25):
    g = np.random.default_rng(seed)
    lim = np.array(model.actuated_joints_limits)
    mid, half = lim.mean(1), (lim[:, 1] - lim[:, 0]) / 2
    t = np.linspace(0, 1, T)[:, None]
    q = mid + half * 0.3 * g.uniform(-1, 1, (1, model.ndof))
    for _ in range(3):
        q = q + half * amp * g.uniform(0.2, 1.0, (1, model.ndof)) * np.sin(
            2 * np.pi * (g.uniform(0.3, 1.5, (1, model.ndof)) * t + g.uniform(0, 1, (1, model.ndof))))
    return np.clip(q, lim[:, 0] + 0.05 * half, lim[:, 1] - 0.05 * half)


def random_configs(model, n, seed, margin=0.0):
    g = torch.Generator().manual_seed(seed)
    lim = torch.tensor(model.actuated_joints_limits, dtype=torch.float64)
    span = lim[:, 1] - lim[:, 0]
    return (lim[:, 0] + margin * span + torch.rand((n, model.ndof), generator=g, dtype=torch.float64) * span * (1 - 2 * margin)).float()


def synthetic_problem(robot_name, P, T, seed=0, noise=0.05):
    """target = FK of a smooth joint path; seeds = path + N(0, noise^2), clamped."""
    model = R.get_model(robot_name)
    qstar = torch.tensor(smooth_joint_path(model, T, seed), dtype=torch.float64)
    target = K.forward_kinematics(model, qstar).float()
    g = torch.Generator().manual_seed(1234)
    x0 = qstar.float()[None] + noise * torch.randn((P, T, model.ndof), generator=g)
    lim = torch.tensor(model.actuated_joints_limits, dtype=torch.float32)
    x0 = torch.minimum(torch.maximum(x0, lim[:, 0]), lim[:, 1])
    return model, target, x0.reshape(P * T, model.ndof).contiguous()
