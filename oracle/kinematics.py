"""Forward kinematics and geometric Jacobian (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates what the reference obtains from `jrl.robot.Robot`:
  * `forward_kinematics(x)` -> [n,7] = [x,y,z,qw,qx,qy,qz]  (optimization_utils.py:811,
    evaluation_utils.py:115; layout per README.md:8)
  * `jacobian(x)` -> [n,6,ndof], rows 0-2 angular, rows 3-5 linear
    (optimization.py:74-80, optimization_utils.py:281, docstring :806-808).
jrl walks the URDF chain multiplying fixed-origin and joint transforms; the same is done
here.  PARITY UNPINNED against jrl itself (not installable offline).
"""
from typing import List, Tuple

import torch

from .robots import RobotModel
from .math_utils import rpy_to_rotation_matrix, rotation_matrix_to_quaternion


def _axis_angle_matrix(axis: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """Rodrigues: unit axis [3], theta [n] -> [n,3,3]."""
    n = theta.shape[0]
    K = torch.zeros((3, 3), dtype=theta.dtype)
    K[0, 1], K[0, 2] = -axis[2], axis[1]
    K[1, 0], K[1, 2] = axis[2], -axis[0]
    K[2, 0], K[2, 1] = -axis[1], axis[0]
    eye = torch.eye(3, dtype=theta.dtype)
    s = torch.sin(theta)[:, None, None]
    c = torch.cos(theta)[:, None, None]
    return eye.expand(n, 3, 3) + s * K + (1 - c) * (K @ K)


def link_frames(model: RobotModel, x: torch.Tensor):
    """World pose of every frame of the chain.

    Returns (Rs, ps, axes, origins):
      Rs[i], ps[i]: rotation [n,3,3] / position [n,3] of frame i (0 = base, i = child link of chain[i-1])
      axes[d], origins[d]: world axis / origin [n,3] of actuated joint d
    """
    n = x.shape[0]
    dt = x.dtype
    R = torch.eye(3, dtype=dt).expand(n, 3, 3).contiguous()
    p = torch.zeros((n, 3), dtype=dt)
    Rs, ps, axes, origins = [R], [p], [], []
    d = 0
    for e in model.chain:
        Rfix = rpy_to_rotation_matrix(e.rpy, dtype=dt)
        tfix = torch.tensor(e.xyz, dtype=dt)
        p = p + R @ tfix
        R = R @ Rfix
        axis_local = torch.tensor(e.axis, dtype=dt)
        if e.jtype == "revolute":
            axes.append(R @ axis_local)
            origins.append(p)
            R = R @ _axis_angle_matrix(axis_local, x[:, d])
            d += 1
        elif e.jtype == "prismatic":
            a = R @ axis_local
            axes.append(a)
            origins.append(p)
            p = p + a * x[:, d : d + 1]
            d += 1
        Rs.append(R)
        ps.append(p)
    assert d == model.ndof == x.shape[1]
    return Rs, ps, axes, origins


def forward_kinematics(model: RobotModel, x: torch.Tensor) -> torch.Tensor:
    Rs, ps, _, _ = link_frames(model, x)
    return torch.cat([ps[-1], rotation_matrix_to_quaternion(Rs[-1])], dim=1)


def forward_kinematics_matrix(model: RobotModel, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    Rs, ps, _, _ = link_frames(model, x)
    return Rs[-1], ps[-1]


def jacobian(model: RobotModel, x: torch.Tensor) -> torch.Tensor:
    """Geometric Jacobian [n,6,ndof]: revolute column [a ; a x (p_ee - o)], prismatic [0 ; a]."""
    Rs, ps, axes, origins = link_frames(model, x)
    n = x.shape[0]
    J = torch.zeros((n, 6, model.ndof), dtype=x.dtype)
    p_ee = ps[-1]
    for d, ci in enumerate(model.actuated):
        a = axes[d]
        if model.chain[ci].jtype == "revolute":
            J[:, 0:3, d] = a
            J[:, 3:6, d] = torch.cross(a, p_ee - origins[d], dim=1)
        else:
            J[:, 3:6, d] = a
    return J


def capsule_world_endpoints(model: RobotModel, x: torch.Tensor):
    """World endpoints of every collision capsule: P1, P2 [n,C,3], radii [C]; plus the joint data."""
    Rs, ps, axes, origins = link_frames(model, x)
    P1, P2 = [], []
    for c in model.capsules:
        R, p = Rs[c.frame], ps[c.frame]
        P1.append(p + R @ torch.tensor(c.p1, dtype=x.dtype))
        P2.append(p + R @ torch.tensor(c.p2, dtype=x.dtype))
    radii = torch.tensor([c.radius for c in model.capsules], dtype=x.dtype)
    return torch.stack(P1, dim=1), torch.stack(P2, dim=1), radii, axes, origins


def joints_moving_frame(model: RobotModel, frame: int) -> List[int]:
    """dof indices of the actuated joints that move link-frame `frame` (joint chain[i] moves frames > i)."""
    return [d for d, ci in enumerate(model.actuated) if ci < frame]
