"""Joint-change / pose-error metrics with the reference's names (cppflow/evaluation_utils.py)."""
from typing import List, Tuple

import numpy as np
import torch


def angular_changes(qpath: torch.Tensor) -> torch.Tensor:
    """evaluation_utils.py:144-154: wrapped first difference, results may be negative."""
    dqs = qpath[1:] - qpath[0:-1]
    if isinstance(qpath, torch.Tensor):
        return torch.remainder(dqs + torch.pi, 2 * torch.pi) - torch.pi
    return np.remainder(dqs + np.pi, 2 * np.pi) - np.pi


def prismatic_changes(x: torch.Tensor) -> torch.Tensor:
    return x[1:] - x[0:-1]


def calculate_mjac_deg(x: torch.Tensor) -> float:
    return torch.rad2deg(angular_changes(x).abs().max()).item()


def calculate_per_timestep_mjac_deg(x: torch.Tensor) -> torch.Tensor:
    return torch.max(torch.rad2deg(angular_changes(x).abs()), dim=1).values


def calculate_per_timestep_mjac_cm(x: torch.Tensor) -> torch.Tensor:
    return 100 * torch.max(prismatic_changes(x).abs(), dim=1).values


def get_mjacs(robot, qpath: torch.Tensor):
    qps_revolute, qps_prismatic = robot.split_configs_to_revolute_and_prismatic(qpath)
    if qps_prismatic.numel() > 0:
        return calculate_mjac_deg(qps_revolute), calculate_per_timestep_mjac_cm(qps_prismatic).abs().max().item()
    return calculate_mjac_deg(qps_revolute), 0.0


def joint_limits_exceeded(robot_joint_limits: List[Tuple[float, float]], qs: np.ndarray):
    """evaluation_utils.py:16-26"""
    assert len(robot_joint_limits) == qs.shape[1]
    n = qs.shape[0]
    pcts = []
    for i, (l, u) in enumerate(robot_joint_limits):
        assert l < u
        n_violating = (qs[:, i] < l).sum() + (u < qs[:, i]).sum()
        pcts.append(100 * n_violating / n)
    return any(vp > 0 for vp in pcts), pcts


def _abs_max(v, empty=0.0) -> float:
    if isinstance(v, torch.Tensor):
        return float(v.abs().max()) if v.numel() > 0 else empty
    return abs(float(v))


def errors_are_below_threshold(max_allowed_position_error_cm, max_allowed_rotation_error_deg, max_allowed_mjac_deg,
                               max_allowed_mjac_cm, error_t_cm, error_R_deg, qdeltas_revolute_deg, qdeltas_prismatic_cm,
                               verbosity: int = 0):
    """evaluation_utils.py:29-75 (strict '<' on the maxima).  The four error arguments may be the reference's tensors
    (per-waypoint errors / joint deltas, reduced here with max / abs-max) or already-reduced numbers - the CUDA
    metrics kernel reduces them on the device (csrc/k_metrics.cu) and the host only sees the maxima."""
    max_pos_cm = float(error_t_cm.max()) if isinstance(error_t_cm, torch.Tensor) else float(error_t_cm)
    max_rot_deg = float(error_R_deg.max()) if isinstance(error_R_deg, torch.Tensor) else float(error_R_deg)
    mjac_deg = _abs_max(qdeltas_revolute_deg)
    mjac_cm = _abs_max(qdeltas_prismatic_cm)  # no prismatic joints: 0 < threshold (the reference's `else True`)
    pose_pos_valid = max_pos_cm < max_allowed_position_error_cm
    pose_rot_valid = max_rot_deg < max_allowed_rotation_error_deg
    mjac_rev_valid = mjac_deg < max_allowed_mjac_deg
    mjac_pris_valid = mjac_cm < max_allowed_mjac_cm
    if verbosity > 0:
        for ok, txt in ((pose_pos_valid, f"pose-position is invalid: {max_pos_cm} < {max_allowed_position_error_cm}"),
                        (pose_rot_valid, f"pose-rotation is invalid: {max_rot_deg} < {max_allowed_rotation_error_deg}"),
                        (mjac_rev_valid, f"mjac_rev is invalid: {mjac_deg:.3f} > {max_allowed_mjac_deg:.3f}"),
                        (mjac_pris_valid, f"mjac_pris is invalid: {mjac_cm:.3f} > {max_allowed_mjac_cm:.3f}")):
            if not ok:
                print("errors_are_below_threshold() |", txt)
    return (pose_pos_valid and pose_rot_valid and mjac_rev_valid and mjac_pris_valid,
            (pose_pos_valid, pose_rot_valid, mjac_rev_valid, mjac_pris_valid))


# ======================
#   ==  Pose error  ==
#


def geodesic_distance_between_quaternions(q1: torch.Tensor, q2: torch.Tensor) -> torch.Tensor:
    """jrl.math_utils.geodesic_distance_between_quaternions (quoted at data_types.py:408-411): 2 acos(dot) with the
    1e-7 clamp, folded into [0, pi] so that q and -q are the same rotation."""
    acos_clamp_epsilon = 1e-7
    dot = torch.clip(torch.sum(q1 * q2, dim=1), -1, 1)
    distance = 2 * torch.acos(torch.clamp(dot, -1 + acos_clamp_epsilon, 1 - acos_clamp_epsilon))
    return torch.abs(torch.remainder(distance + torch.pi, 2 * torch.pi) - torch.pi)


def positional_errors(path_1: torch.Tensor, path_2: torch.Tensor) -> torch.Tensor:
    """evaluation_utils.py:131-133"""
    return torch.norm(path_1[:, :3] - path_2[:, :3], dim=1)


def rotational_errors(path_1: torch.Tensor, path_2: torch.Tensor) -> torch.Tensor:
    """evaluation_utils.py:136-138"""
    return geodesic_distance_between_quaternions(path_1[:, 3:], path_2[:, 3:])


def calculate_pose_error_cm_deg(robot, x: torch.Tensor, target_path: torch.Tensor):
    """evaluation_utils.py:113-117: per-configuration positional (cm) and rotational (deg) errors; FK on the GPU."""
    traced_path = robot.forward_kinematics(x)
    return 100 * positional_errors(target_path, traced_path), torch.rad2deg(rotational_errors(target_path, traced_path))
