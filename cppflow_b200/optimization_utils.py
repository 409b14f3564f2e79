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


def _get_prismatic_and_revolute_row_mask(robot, n: int, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """optimization_utils.py:224-236 -> (revolute rows, prismatic rows) of an [n x _] residual"""
    assert n % robot.ndof == 0, f"error - n {n} is not divisible by ndof {robot.ndof}"
    revolute_rows = torch.zeros(robot.ndof, dtype=torch.bool, device=device)
    revolute_rows[robot.revolute_joint_idxs] = True
    revolute_rows = revolute_rows.tile(n // robot.ndof)
    return revolute_rows, torch.logical_not(revolute_rows)


def _get_rotation_and_position_row_mask(n: int, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """optimization_utils.py:239-250 -> (rotation rows, position rows) of a [6 n x _] pose residual"""
    rotation_rows = torch.zeros(6, dtype=torch.bool, device=device)
    rotation_rows[:3] = True
    rotation_rows = rotation_rows.tile(n)
    return rotation_rows, torch.logical_not(rotation_rows)


def filter_rows_from_r_J_differencing(robot, r: torch.Tensor, J: torch.Tensor, threshold_rad: float, threshold_m: float,
                                      shift_to_threshold: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """optimization_utils.py:736-768: keep the differencing rows whose |residual| exceeds the threshold of their joint
    type, optionally moved towards zero by the threshold (r is modified in place, like the reference)."""
    assert r.shape[0] == J.shape[0]
    assert r.shape[0] % robot.ndof == 0, f"r.shape[0]: {r.shape[0]}, robot.ndof: {robot.ndof} for {robot}"
    revolute_idxs, prismatic_idxs = _get_prismatic_and_revolute_row_mask(robot, r.shape[0], r.device)
    invalid_idxs_rev = torch.logical_and(r.abs()[:, 0] > threshold_rad, revolute_idxs)
    invalid_idxs_pris = torch.logical_and(r.abs()[:, 0] > threshold_m, prismatic_idxs)
    keep_idxs = torch.logical_or(invalid_idxs_rev, invalid_idxs_pris)
    if not shift_to_threshold:
        return r[keep_idxs, :], J[keep_idxs, :]
    r[torch.logical_and((r < -threshold_rad)[:, 0], revolute_idxs)] += threshold_rad
    r[torch.logical_and((r > threshold_rad)[:, 0], revolute_idxs)] -= threshold_rad
    r[torch.logical_and((r < -threshold_m)[:, 0], prismatic_idxs)] += threshold_m
    r[torch.logical_and((r > threshold_m)[:, 0], prismatic_idxs)] -= threshold_m
    return r[keep_idxs, :], J[keep_idxs, :]


class LmResidualFns:
    """Dense r and J of ONE path, assembled on the GPU from the CUDA kernels' per-term outputs, in the reference's
    row order (pose, differencing, virtual configs, self collisions, env collisions) and sign convention
    (J = -dr/dx).  O((T*D)^2) memory: for tests and debugging, not for the solver.

    The row scaling / filtering options of OptimizationParameters (`pose_do_scale_down_satisfied`,
    `differencing_do_scale_satisfied`, `differencing_do_ignore_satisfied`) are implemented HERE, on the dense form, with
    the reference's helper functions below; the block-tridiagonal CUDA solver does not take them (ops.make_params
    raises).  They are off in both live parameter sets (lm_hyper_parameters.py:86-151) and unreachable in the reference
    itself: its get_r_and_J reads the thresholds from `pms.constraints` (optimization_utils.py:515-520, :562-567), a
    field OptimizationParameters does not declare (lm_hyper_parameters.py:14-63) - switching an option on raises
    AttributeError there, and here unless the caller has attached a `constraints` attribute to the parameters."""

    @staticmethod
    def _scale_down_rows_from_r_J_pose_below_error(r: torch.Tensor, J: torch.Tensor, error_threshold_m: float,
                                                   error_threshold_rad: float, scale: float,
                                                   shift_invalid_to_threshold: bool = False):
        """optimization_utils.py:288-329 -> (r, J, invalid_row_idxs), r and J scaled in place"""
        assert r.shape[0] == J.shape[0]
        assert r.shape[0] % 6 == 0
        assert r.shape[0] == r.numel()
        assert 0.0 <= scale < 1.0, "scale should be in [0, 1)"
        rotation_rows, position_rows = _get_rotation_and_position_row_mask(r.numel() // 6, r.device)
        rot_below_threshold_rows = r[:, 0].abs() < error_threshold_rad
        pos_below_threshold_rows = r[:, 0].abs() < error_threshold_m
        do_scale_rotation = torch.logical_and(rot_below_threshold_rows, rotation_rows)
        do_scale_position = torch.logical_and(pos_below_threshold_rows, position_rows)
        r[do_scale_position, :] *= scale
        r[do_scale_rotation, :] *= scale
        J[do_scale_position, :] *= scale
        J[do_scale_rotation, :] *= scale
        if shift_invalid_to_threshold:
            invalid_rot = torch.logical_and(torch.logical_not(rot_below_threshold_rows), rotation_rows)
            invalid_pos = torch.logical_and(torch.logical_not(pos_below_threshold_rows), position_rows)
            r[torch.logical_and((r < -error_threshold_rad)[:, 0], invalid_rot)] += error_threshold_rad
            r[torch.logical_and((r > error_threshold_rad)[:, 0], invalid_rot)] -= error_threshold_rad
            r[torch.logical_and((r < -error_threshold_m)[:, 0], invalid_pos)] += error_threshold_m
            r[torch.logical_and((r > error_threshold_m)[:, 0], invalid_pos)] -= error_threshold_m
        return r, J, torch.logical_not(torch.logical_or(do_scale_rotation, do_scale_position))

    @staticmethod
    def _scale_down_rows_from_r_J_differencing_below_error(robot, r: torch.Tensor, J: torch.Tensor, mjac_threshold_m: float,
                                                           mjac_threshold_rad: float, scale: float,
                                                           shift_invalid_to_threshold: bool = False):
        """optimization_utils.py:352-397 -> (J, r, invalid_row_idxs) [J first, as the reference], scaled in place"""
        ndof = robot.ndof
        n = (r.shape[0] // ndof) + 1
        assert 0.0 <= scale < 1.0, "values should be scaled down, not up"
        assert r.shape[0] % ndof == 0
        assert J.shape == ((n - 1) * ndof, n * ndof), f"J is {J.shape}, should be ((n-1)*ndof, n*ndof)"
        assert r.shape == ((n - 1) * ndof, 1)
        revolute_rows, prismatic_rows = _get_prismatic_and_revolute_row_mask(robot, r.numel(), r.device)
        prismatic_below_threshold_rows = r[:, 0].abs() < mjac_threshold_m
        revolute_below_threshold_rows = r[:, 0].abs() < mjac_threshold_rad
        valid_prismatic = torch.logical_and(prismatic_below_threshold_rows, prismatic_rows)
        invalid_prismatic = torch.logical_and(torch.logical_not(prismatic_below_threshold_rows), prismatic_rows)
        valid_revolute = torch.logical_and(revolute_below_threshold_rows, revolute_rows)
        invalid_revolute = torch.logical_and(torch.logical_not(revolute_below_threshold_rows), revolute_rows)
        r[valid_prismatic, :] *= scale
        r[valid_revolute, :] *= scale
        J[valid_prismatic, :] *= scale
        J[valid_revolute, :] *= scale
        if shift_invalid_to_threshold:
            r[torch.logical_and((r < -mjac_threshold_rad)[:, 0], invalid_revolute)] += mjac_threshold_rad
            r[torch.logical_and((r > mjac_threshold_rad)[:, 0], invalid_revolute)] -= mjac_threshold_rad
            r[torch.logical_and((r < -mjac_threshold_m)[:, 0], invalid_prismatic)] += mjac_threshold_m
            r[torch.logical_and((r > mjac_threshold_m)[:, 0], invalid_prismatic)] -= mjac_threshold_m
        return J, r, torch.logical_not(torch.logical_or(valid_prismatic, valid_revolute))

    @staticmethod
    def get_r_and_J(pms: OptimizationParameters, robot, x: torch.Tensor, target_path: torch.Tensor,
                    Tcuboids: Optional[List] = None, cuboids: Optional[List] = None) -> Tuple[LmJacobian, LmResidual]:
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
            Jp[rows[:, :, None], cols[:, None, :]] = Jfk
            rp = e[:, :, 0].reshape(-1, 1).clone()
            if getattr(pms, "pose_do_scale_down_satisfied", False):  # optimization_utils.py:513-531
                thr_m = pms.pose_ignore_satisfied_threshold_scale * pms.constraints.max_allowed_position_error_m
                thr_rad = pms.pose_ignore_satisfied_threshold_scale * pms.constraints.max_allowed_rotation_error_deg
                rp, Jp, _ = LmResidualFns._scale_down_rows_from_r_J_pose_below_error(
                    rp, Jp, error_threshold_m=thr_m, error_threshold_rad=thr_rad, scale=pms.pose_ignore_satisfied_scale_down)
            r.pose = rp * scale.tile(n)[:, None]
            J.pose = Jp * scale.tile(n)[:, None]
        if pms.use_differencing:
            rd = angular_changes(x).reshape((n - 1) * ndof, 1).clone()
            m = ndof * (n - 1)
            Jd = torch.zeros((m, ndof * n), device=dev)
            idx = torch.arange(m, device=dev)
            Jd[idx, idx] = 1.0
            Jd[idx, idx + ndof] = -1.0
            do_ignore = getattr(pms, "differencing_do_ignore_satisfied", False)
            do_scale = getattr(pms, "differencing_do_scale_satisfied", False)
            assert not (do_scale and do_ignore), "use one or the other, not both"
            if do_ignore or do_scale:  # optimization_utils.py:560-593
                import math

                thr_rad = math.radians(pms.constraints.max_allowed_mjac_deg - pms.differencing_ignore_satisfied_margin_deg)
                thr_m = (pms.constraints.max_allowed_mjac_cm - pms.differencing_ignore_satisfied_margin_cm) / 100.0
            if do_ignore:
                rd, Jd = filter_rows_from_r_J_differencing(robot, rd, Jd, threshold_rad=thr_rad, threshold_m=thr_m,
                                                           shift_to_threshold=True)
            if do_scale:
                Jd, rd, _ = LmResidualFns._scale_down_rows_from_r_J_differencing_below_error(
                    robot, rd, Jd, mjac_threshold_m=thr_m, mjac_threshold_rad=thr_rad,
                    scale=pms.differencing_scale_down_satisfied_scale,
                    shift_invalid_to_threshold=pms.differencing_scale_down_satisfied_shift_invalid_to_threshold)
            if robot.has_prismatic_joints and not do_ignore:  # optimization_utils.py:606-609
                _, pris = _get_prismatic_and_revolute_row_mask(robot, rd.shape[0], dev)
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
