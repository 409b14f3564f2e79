"""Quaternion / angle helpers the reference imports from `jrl.math_utils`
(optimization_utils.py:8, evaluation_utils.py:4, data_types.py:11).  TEST INFRASTRUCTURE.

jrl 0.1.2 is not installable offline: these follow the published formulas (Hamilton
product, wxyz order; pytorch3d-style rotation-matrix -> quaternion) - PARITY UNPINNED,
except `geodesic_distance_between_quaternions`, which the reference quotes verbatim in
data_types.py:408-411.
"""
import math
import torch


def quaternion_conjugate(q: torch.Tensor) -> torch.Tensor:
    return torch.cat([q[:, 0:1], -q[:, 1:4]], dim=1)


def quaternion_norm(q: torch.Tensor) -> torch.Tensor:
    return torch.norm(q, dim=1)


def quaternion_inverse(q: torch.Tensor) -> torch.Tensor:
    """Inverse of (unit) quaternions = conjugate; used at optimization_utils.py:816."""
    return quaternion_conjugate(q)


def quaternion_product(q1: torch.Tensor, q2: torch.Tensor) -> torch.Tensor:
    """Hamilton product, wxyz; used at optimization_utils.py:817."""
    w1, x1, y1, z1 = q1[:, 0], q1[:, 1], q1[:, 2], q1[:, 3]
    w2, x2, y2, z2 = q2[:, 0], q2[:, 1], q2[:, 2], q2[:, 3]
    return torch.stack(
        [
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        ],
        dim=1,
    )


def quaternion_to_rpy(q: torch.Tensor) -> torch.Tensor:
    """[n,4] wxyz -> [n,3] roll, pitch, yaw; used at optimization_utils.py:818."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = torch.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    pitch = torch.asin(torch.clamp(2 * (w * y - z * x), -1.0, 1.0))
    yaw = torch.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return torch.stack([roll, pitch, yaw], dim=1)


def angular_subtraction(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a - b wrapped to [-pi, pi); used at optimization_utils.py:468,473."""
    return torch.remainder(a - b + torch.pi, 2 * torch.pi) - torch.pi


def geodesic_distance_between_quaternions(q1: torch.Tensor, q2: torch.Tensor) -> torch.Tensor:
    """The first two lines are quoted by the reference at data_types.py:408-411 (acos clamp epsilon 1e-7).  jrl then
    folds the angle into [0, pi] (`abs(remainder(d + pi, 2 pi) - pi)`), which makes q and -q the same rotation:
    2 acos(dot) -> 2 acos(|dot|).  Without it a target quaternion stored with the opposite sign (the csv target paths
    carry arbitrary signs) would read as a 360 degree error and no plan could ever be valid."""
    acos_clamp_epsilon = 1e-7
    dot = torch.clip(torch.sum(q1 * q2, dim=1), -1, 1)
    distance = 2 * torch.acos(torch.clamp(dot, -1 + acos_clamp_epsilon, 1 - acos_clamp_epsilon))
    return torch.abs(torch.remainder(distance + torch.pi, 2 * torch.pi) - torch.pi)


def rpy_to_rotation_matrix(rpy, dtype=torch.float64) -> torch.Tensor:
    """URDF fixed-axis roll/pitch/yaw -> R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    R = [
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ]
    # URDF angles are multiples of pi/4: snap the 1e-17 residue of cos(pi/2) to exact zeros
    R = [[0.0 if abs(v) < 1e-12 else v for v in row] for row in R]
    return torch.tensor(R, dtype=dtype)


def rotation_matrix_to_quaternion(R: torch.Tensor) -> torch.Tensor:
    """[n,3,3] -> [n,4] wxyz.  Four-candidate method (pytorch3d `matrix_to_quaternion`):
    the component of largest magnitude is made positive."""
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    q_abs = torch.sqrt(
        torch.clamp(
            torch.stack(
                [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=1
            ),
            min=0.0,
        )
    )
    cand = torch.stack(
        [
            torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=1),
            torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], dim=1),
            torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], dim=1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], dim=1),
        ],
        dim=1,
    )  # [n, 4(candidate), 4]
    cand = cand / (2.0 * torch.clamp(q_abs, min=0.1))[:, :, None]
    idx = torch.argmax(q_abs, dim=1)
    return cand[torch.arange(R.shape[0]), idx]


def quaternion_to_rotation_matrix(q: torch.Tensor) -> torch.Tensor:
    """[n,4] wxyz (unit) -> [n,3,3]."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
        ],
        dim=1,
    )
    return R.reshape(-1, 3, 3)
