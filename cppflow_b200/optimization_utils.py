"""Host-side mirror of the hot-path functions of cppflow/optimization_utils.py, forwarding to the CUDA library.

  get_6d_pose_errors     optimization_utils.py:802-820
  clamp_to_joint_limits  optimization_utils.py:823-833 (in place, returns x)
  x_is_valid             optimization_utils.py:836-923 (thresholds on the GPU; the per-config klampt mesh loop of
                         :889-900 is replaced by the capsule distances, with a host callback for a mesh checker)
  LmResidualFns          optimization_utils.py:253-731 (dense r / J for inspection and parity tests; the solver
                         itself never materialises them - see csrc/k_lm_full.cu)
"""
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch

from . import ops
from .config import ENV_COLLISIONS_IGNORED, SELF_COLLISIONS_IGNORED
from .data_types import Constraints, Problem
from .evaluation_utils import angular_changes, errors_are_below_threshold
from .lm_hyper_parameters import OptimizationParameters


def get_6d_pose_errors(robot, x: torch.Tensor, target_poses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> ([n, 6, 1] = [roll, pitch, yaw, x, y, z] errors in rad / m, current poses [n, 7])."""
    assert target_poses.shape[0] == x.shape[0], "one target pose per configuration"
    err, cur = ops.pose_errors(robot.robot_id, robot.ndof, x, target_poses)
    return err.unsqueeze(2), cur


def clamp_to_joint_limits(robot, x: torch.Tensor, verbosity: int = 0) -> torch.Tensor:
    if verbosity > 0:
        for i, (l, u) in enumerate(robot.actuated_joints_limits):
            if x[:, i].min() < l:
                print(f"clamp_to_joint_limits() | joint {i} is below lower limit {l}")
            if x[:, i].max() > u:
                print(f"clamp_to_joint_limits() | joint {i} is above upper limit {u}")
    if x.is_cuda and x.dtype == torch.float32 and x.is_contiguous():
        return ops.clamp_to_joint_limits_(robot.robot_id, robot.ndof, x)
    raise RuntimeError("clamp_to_joint_limits needs a contiguous fp32 CUDA tensor (no CPU fallback)")


# A mesh-level checker (e.g. klampt) can be plugged in: (problem, x_i [T, ndof]) -> (self_collides, env_collides)
MeshValidator = Callable[[Problem, torch.Tensor], Tuple[bool, bool]]


def path_metrics(problem: Problem, x: torch.Tensor, parallel_count: int = 1) -> torch.Tensor:
    """[P, 8] per-path metrics (ops.METRIC_NAMES) from one kernel launch."""
    robot = problem.robot
    return ops.path_metrics(robot.robot_id, robot.ndof, x, problem.target_path, parallel_count, problem.n_timesteps,
                            problem.obstacle_tables)


def x_is_valid(problem: Problem, constraints: Constraints, target_path_stacked: torch.Tensor, x: torch.Tensor,
               parallel_count: int, results_df=None, verbosity: int = 0, mesh_validator: Optional[MeshValidator] = None,
               metrics: Optional[torch.Tensor] = None):
    """Returns (x_i or None, i or None, (pose_pos_valid, pose_rot_valid, mjac_rev_valid, mjac_pris_valid,
    is_a_self_collision, is_a_env_collision)) like the reference; the flags are those of the last path examined."""
    if results_df is not None:
        raise NotImplementedError("results_df logging is not available (data_types.py:419-420 raises in the reference too)")
    n = problem.n_timesteps
    assert x.shape[0] == n * parallel_count
    m = (path_metrics(problem, x, parallel_count) if metrics is None else metrics).cpu()  # the ONE sync per call
    is_a_self_collision = None
    is_a_env_collision = None
    flags = (False, False, False, False)
    for i in range(parallel_count):
        max_pos_cm, max_rot_deg, mjac_deg, mjac_cm, _tl, min_self, min_env, _ = m[i].tolist()
        all_valid, flags = errors_are_below_threshold(
            constraints.max_allowed_position_error_cm, constraints.max_allowed_rotation_error_deg,
            constraints.max_allowed_mjac_deg, constraints.max_allowed_mjac_cm, max_pos_cm, max_rot_deg, mjac_deg, mjac_cm,
        )
        if not all_valid:
            continue
        x_i = x[i * n : (i + 1) * n, :]
        if mesh_validator is not None:
            is_a_self_collision, is_a_env_collision = mesh_validator(problem, x_i)
        else:
            is_a_self_collision = min_self < 0.0
            is_a_env_collision = min_env < 0.0
        if not SELF_COLLISIONS_IGNORED and is_a_self_collision:
            continue
        if not ENV_COLLISIONS_IGNORED and is_a_env_collision:
            continue
        return x_i, i, (*flags, is_a_self_collision, is_a_env_collision)
    return None, None, (*flags, is_a_self_collision, is_a_env_collision)


# ------------------------------------------------------------------------------------------------------------
# dense residual / Jacobian (inspection + parity tests only)


@dataclass
class LmResidual:
    pose: Optional[torch.Tensor] = None
    differencing: Optional[torch.Tensor] = None
    virtual_configs: Optional[torch.Tensor] = None
    self_collisions: Optional[torch.Tensor] = None
    env_collisions: Optional[torch.Tensor] = None

    def get_r(self) -> torch.Tensor:
        parts = [p for p in (self.pose, self.differencing, self.virtual_configs, self.self_collisions, self.env_collisions)
                 if p is not None and p.shape[0] > 0]
        return torch.cat(parts, dim=0)


@dataclass
class LmJacobian:
    pose: Optional[torch.Tensor] = None
    differencing: Optional[torch.Tensor] = None
    virtual_configs: Optional[torch.Tensor] = None
    self_collisions: Optional[torch.Tensor] = None
    env_collisions: Optional[torch.Tensor] = None

    def get_J(self) -> torch.Tensor:
        parts = [p for p in (self.pose, self.differencing, self.virtual_configs, self.self_collisions, self.env_collisions)
                 if p is not None and p.shape[0] > 0]
        return torch.cat(parts, dim=0)


def _wrap(a: torch.Tensor) -> torch.Tensor:
    return torch.remainder(a + torch.pi, 2 * torch.pi) - torch.pi


class LmResidualFns:
    """Dense r and J of ONE path, assembled on the GPU from the CUDA kernels' per-term outputs, in the reference's
    row order (pose, differencing, virtual configs, self collisions, env collisions) and sign convention
    (J = -dr/dx).  O((T*D)^2) memory: for tests and debugging, not for the solver."""

    @staticmethod
    def get_r_and_J(pms: OptimizationParameters, robot, x: torch.Tensor, target_path: torch.Tensor,
                    Tcuboids: Optional[List] = None, cuboids: Optional[List] = None) -> Tuple[LmJacobian, LmResidual]:
        ops.make_params(pms)  # rejects the unsupported row-scaling options loudly
        n, ndof = x.shape
        dev = x.device
        r, J = LmResidual(), LmJacobian()
        cols = torch.arange(n, device=dev)[:, None] * ndof + torch.arange(ndof, device=dev)[None, :]  # [n, ndof]
        if pms.use_pose:
            e, _ = get_6d_pose_errors(robot, x, target_path)
            Jfk = robot.jacobian(x)
            scale = torch.tensor([pms.alpha_rotation] * 3 + [pms.alpha_position] * 3, device=dev)
            Jp = torch.zeros((6 * n, ndof * n), device=dev)
            rows = torch.arange(n, device=dev)[:, None] * 6 + torch.arange(6, device=dev)[None, :]
            Jp[rows[:, :, None], cols[:, None, :]] = Jfk * scale[None, :, None]
            r.pose = (e[:, :, 0] * scale[None, :]).reshape(-1, 1)
            J.pose = Jp
        if pms.use_differencing:
            rd = angular_changes(x).reshape((n - 1) * ndof, 1).clone()
            m = ndof * (n - 1)
            Jd = torch.zeros((m, ndof * n), device=dev)
            idx = torch.arange(m, device=dev)
            Jd[idx, idx] = 1.0
            Jd[idx, idx + ndof] = -1.0
            if robot.has_prismatic_joints:
                pris = torch.zeros(ndof, dtype=torch.bool, device=dev)
                pris[robot.prismatic_joint_idxs] = True
                pris = pris.tile(n - 1)
                rd[pris] *= pms.alpha_differencing_prismatic_scaling
                Jd[pris] *= pms.alpha_differencing_prismatic_scaling
            r.differencing = pms.alpha_differencing * rd
            J.differencing = pms.alpha_differencing * Jd
        if pms.use_virtual_configs:
            xv = pms.virtual_configs
            assert xv is not None and xv.shape == x.shape
            nv = pms.n_virtual_configs
            assert 2 * nv < n
            sel = torch.cat([torch.arange(nv, device=dev), torch.arange(n - nv, n, device=dev)])
            s = pms.alpha_virtual_configs * pms.alpha_differencing
            r.virtual_configs = s * _wrap(x[sel] - xv.to(dev)[sel]).reshape(-1, 1)
            Jv = torch.zeros((2 * nv * ndof, ndof * n), device=dev)
            Jv[torch.arange(2 * nv * ndof, device=dev), cols[sel].reshape(-1)] = -s
            J.virtual_configs = Jv

        def collision_rows(dists, Jc, alpha):
            rr = (-alpha * dists).reshape(-1, 1)
            mask = (rr > 0).reshape(-1)
            if not bool(mask.any()):
                return rr[mask], None
            S = dists.shape[1]
            Jfull = torch.zeros((n * S, ndof * n), device=dev)
            rows = torch.arange(n * S, device=dev).reshape(n, S)
            Jfull[rows[:, :, None], cols[:, None, :]] = alpha * Jc
            return rr[mask], Jfull[mask]

        if pms.use_self_collisions:
            d = robot.self_collision_distances(x)
            r.self_collisions, J.self_collisions = collision_rows(d, robot.self_collision_distances_jacobian(x),
                                                                  pms.alpha_self_collision)
        if pms.use_env_collisions and Tcuboids is not None and len(Tcuboids) > 0:
            rs, Js = [], []
            for Tcuboid, cuboid in zip(Tcuboids, cuboids):
                d = robot.env_collision_distances(x, cuboid, Tcuboid)
                rr, JJ = collision_rows(d, robot.env_collision_distances_jacobian(x, cuboid, Tcuboid), pms.alpha_env_collision)
                if JJ is not None:
                    rs.append(rr)
                    Js.append(JJ)
            if rs:
                r.env_collisions, J.env_collisions = torch.cat(rs, dim=0), torch.cat(Js, dim=0)
        return J, r
