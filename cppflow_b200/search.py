"""dp_search with the reference's signature (cppflow/search.py:128-135).  The DP, the k x k x (T-1) joint-jump
tensor, the joint-limit flags and the backtrack all run on the GPU (csrc/k_search.cu); the returned path is
bit-identical to the reference's."""
import numpy as np
import torch

from . import ops

K_JLIM_COST = 100  # search.py:14
K_COLLISION_COST = 1000  # search.py:15
DEFAULT_JLIM_SAFETY_PADDING_REVOLUTE = np.deg2rad(1.5)  # search.py:20
DEFAULT_JLIM_SAFETY_PADDING_PRISMATIC = 3 / 100.0  # search.py:21 (cm_to_m(3))


def joint_limit_almost_violations_3d(robot, qs: torch.Tensor, eps_revolute: float = DEFAULT_JLIM_SAFETY_PADDING_REVOLUTE,
                                     eps_prismatic: float = DEFAULT_JLIM_SAFETY_PADDING_PRISMATIC) -> torch.Tensor:
    """[k, T, ndof] -> float32 [k, T], 1 where a joint is within eps of its limit (search.py:25-52)."""
    assert len(qs.shape) == 3
    k, T, ndof = qs.shape
    flags = ops.joint_limit_flags(robot.robot_id, robot.ndof, qs.reshape(k * T, ndof), eps_revolute, eps_prismatic)
    return flags.reshape(k, T)


def dp_search(robot, q: torch.Tensor, self_collision_violations: torch.Tensor, env_collision_violations: torch.Tensor,
              use_cuda: bool = False, verbosity: int = 1, return_details: bool = False):
    """q [k, T, ndof] -> best path [T, ndof] (search.py:128-173).

    `use_cuda` is accepted (with the reference's default) for signature compatibility and ignored: the search always
    runs on the device q lives on and the path is returned there.  The reference moves q to the CPU unless use_cuda is
    set (search.py:140-141) and its caller moves the result back (`.to(DEVICE)`, planners.py:274) - a no-op here."""
    best, memo, costs, chosen = ops.dp_search(robot.robot_id, robot.ndof, q, self_collision_violations,
                                              env_collision_violations)
    if return_details:
        return best, memo, costs, chosen
    return best
