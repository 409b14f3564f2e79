"""Levenberg-Marquardt residual / Jacobian assembly and steps (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates, term by term and dense like the reference:
  optimization.py:61-92    levenberg_marquardt_only_pose
  optimization.py:95-113   _lm_full_step  (dense J^T J + lambda I, Cholesky, two triangular solves)
  optimization.py:116-144  levenberg_marquardt_full
  optimization_utils.py:264-285   pose residual / Jacobian
  optimization_utils.py:335-349   differencing residual / Jacobian
  optimization_utils.py:430-483   virtual-config residual / Jacobian
  optimization_utils.py:486-731   LmResidualFns.get_r_and_J (row order :89-102, :185-198)
  optimization_utils.py:802-833   get_6d_pose_errors, clamp_to_joint_limits
  evaluation_utils.py:29-75,97-98,113-154   thresholds, pose errors, angular_changes
  lm_hyper_parameters.py:86-151   ALT_LOSS_V2_1_DIFF / ALT_LOSS_V2_1_POSE
Everything is dtype-generic: run in float32 to mimic the reference, float64 for a ground truth.
"""
from dataclasses import dataclass, replace
from typing import List, Optional, Tuple

import numpy as np
import torch

from .robots import RobotModel
from . import kinematics as K
from . import geometry as G
from .math_utils import (
    quaternion_inverse,
    quaternion_product,
    quaternion_to_rpy,
    angular_subtraction,
    geodesic_distance_between_quaternions,
)


@dataclass
class LmParams:
    """The subset of `OptimizationParameters` (lm_hyper_parameters.py:14-80) the live parameter sets use."""

    lm_lambda: float = 1e-6
    alpha_position: float = 3.5
    alpha_rotation: float = 0.35
    alpha_differencing: float = 0.00375
    alpha_differencing_prismatic_scaling: float = 1.0
    alpha_virtual_configs: float = 1.0
    alpha_self_collision: float = 0.01
    alpha_env_collision: float = 0.01
    use_pose: bool = False
    use_differencing: bool = True
    use_virtual_configs: bool = True
    n_virtual_configs: int = 4
    use_self_collisions: bool = True
    use_env_collisions: bool = True
    virtual_configs: Optional[torch.Tensor] = None
    # row scaling / filtering options (off in both live parameter sets; lm_hyper_parameters.py:30-45).  The reference's
    # get_r_and_J reads the thresholds from `pms.constraints` (optimization_utils.py:515-520, :562-567) - a field its
    # OptimizationParameters does not have, so it cannot reach these branches today; here `constraints` is the tuple
    # (max_allowed_position_error_cm, max_allowed_rotation_error_deg, max_allowed_mjac_deg, max_allowed_mjac_cm).
    pose_do_scale_down_satisfied: bool = False
    pose_ignore_satisfied_threshold_scale: float = 1.0
    pose_ignore_satisfied_scale_down: float = 0.5
    differencing_do_ignore_satisfied: bool = False
    differencing_ignore_satisfied_margin_deg: float = 0.0
    differencing_ignore_satisfied_margin_cm: float = 0.0
    differencing_do_scale_satisfied: bool = False
    differencing_scale_down_satisfied_scale: float = 0.5
    differencing_scale_down_satisfied_shift_invalid_to_threshold: bool = False
    constraints: Optional[Tuple[float, float, float, float]] = None


# lm_hyper_parameters.py:86-118
ALT_LOSS_V2_1_DIFF = LmParams()
# lm_hyper_parameters.py:119-151
ALT_LOSS_V2_1_POSE = LmParams(
    use_pose=True, use_differencing=False, use_virtual_configs=False, use_self_collisions=False,
    use_env_collisions=False,
)
# every term on: the "fused" iteration of BASELINE.json's north_star (pose + joint-difference + collisions)
ALL_TERMS = LmParams(use_pose=True)


def angular_changes(qpath: torch.Tensor) -> torch.Tensor:
    """evaluation_utils.py:144-154"""
    dqs = qpath[1:] - qpath[0:-1]
    return torch.remainder(dqs + torch.pi, 2 * torch.pi) - torch.pi


def prismatic_changes(x: torch.Tensor) -> torch.Tensor:
    """evaluation_utils.py:97-98"""
    return x[1:] - x[0:-1]


def get_6d_pose_errors(model: RobotModel, x: torch.Tensor, target_poses: torch.Tensor):
    """optimization_utils.py:802-820 -> ([n,6,1] = [r,p,y,x,y,z] errors, current poses [n,7])."""
    n = x.shape[0]
    current_poses = K.forward_kinematics(model, x)
    pose_errors = torch.zeros((n, 6, 1), dtype=x.dtype)
    for i in range(3):
        pose_errors[:, i + 3, 0] = target_poses[:, i] - current_poses[:, i]
    current_pose_quat_inv = quaternion_inverse(current_poses[:, 3:7])
    rotation_error_quat = quaternion_product(target_poses[:, 3:], current_pose_quat_inv)
    pose_errors[:, 0:3, 0] = quaternion_to_rpy(rotation_error_quat)
    return pose_errors, current_poses


def clamp_to_joint_limits(model: RobotModel, x: torch.Tensor) -> torch.Tensor:
    """optimization_utils.py:823-833 (in place)."""
    for i, (l, u) in enumerate(model.actuated_joints_limits):
        x[:, i] = torch.clamp(x[:, i], l, u)
    return x


def levenberg_marquardt_only_pose(model: RobotModel, x: torch.Tensor, target_path: torch.Tensor, pms: LmParams,
                                  return_residual: bool = False):
    """optimization.py:61-92: per-waypoint (J^T J + lambda I) dx = J^T e with LU (`torch.linalg.solve`)."""
    n, ndof = x.shape
    error, _ = get_6d_pose_errors(model, x, target_path)
    J_batch = K.jacobian(model, x)
    error[:, 3:, 0] *= pms.alpha_position
    error[:, :3, 0] *= pms.alpha_rotation
    J_batch[:, 3:] *= pms.alpha_position
    J_batch[:, :3] *= pms.alpha_rotation
    J_batch_T = torch.transpose(J_batch, 1, 2)
    eye = torch.eye(ndof, dtype=x.dtype)[None, :, :].repeat(n, 1, 1)
    lhs_A = torch.bmm(J_batch_T, J_batch) + pms.lm_lambda * eye
    rhs_B = torch.bmm(J_batch_T, error)
    delta_x = torch.linalg.solve(lhs_A, rhs_B)
    if return_residual:
        return x + delta_x[:, :, 0], J_batch, error
    return x + delta_x[:, :, 0]


def _row_masks(model: RobotModel, n_rows: int):
    """optimization_utils.py:224-236"""
    revolute = torch.zeros(model.ndof, dtype=torch.bool)
    revolute[model.revolute_joint_idxs] = True
    revolute = revolute.tile(n_rows // model.ndof)
    return revolute, torch.logical_not(revolute)


def _rotation_and_position_row_mask(n: int):
    """optimization_utils.py:239-250: rows [rot, rot, rot, pos, pos, pos] per waypoint"""
    rotation = torch.zeros(6, dtype=torch.bool)
    rotation[:3] = True
    rotation = rotation.tile(n)
    return rotation, torch.logical_not(rotation)


def scale_down_rows_pose_below_error(r: torch.Tensor, J: torch.Tensor, error_threshold_m: float, error_threshold_rad: float,
                                     scale: float):
    """LmResidualFns._scale_down_rows_from_r_J_pose_below_error (optimization_utils.py:288-329, without its untested
    shift option): rows whose |error| is below the threshold of their kind are multiplied by `scale`, in place.
    -> (r, J, invalid_row_idxs)"""
    assert r.shape[0] == J.shape[0] and r.shape[0] % 6 == 0 and 0.0 <= scale < 1.0
    rotation_rows, position_rows = _rotation_and_position_row_mask(r.numel() // 6)
    do_rot = torch.logical_and(r[:, 0].abs() < error_threshold_rad, rotation_rows)
    do_pos = torch.logical_and(r[:, 0].abs() < error_threshold_m, position_rows)
    r[do_pos, :] *= scale
    r[do_rot, :] *= scale
    J[do_pos, :] *= scale
    J[do_rot, :] *= scale
    return r, J, torch.logical_not(torch.logical_or(do_rot, do_pos))


def scale_down_rows_differencing_below_error(model: RobotModel, r: torch.Tensor, J: torch.Tensor, mjac_threshold_m: float,
                                             mjac_threshold_rad: float, scale: float,
                                             shift_invalid_to_threshold: bool = False):
    """LmResidualFns._scale_down_rows_from_r_J_differencing_below_error (optimization_utils.py:352-397), in place.
    -> (J, r, invalid_row_idxs)  [the reference returns J first]"""
    assert 0.0 <= scale < 1.0 and r.shape[0] % model.ndof == 0
    revolute_rows, prismatic_rows = _row_masks(model, r.numel())
    below_m = r[:, 0].abs() < mjac_threshold_m
    below_rad = r[:, 0].abs() < mjac_threshold_rad
    valid_pri = torch.logical_and(below_m, prismatic_rows)
    invalid_pri = torch.logical_and(torch.logical_not(below_m), prismatic_rows)
    valid_rev = torch.logical_and(below_rad, revolute_rows)
    invalid_rev = torch.logical_and(torch.logical_not(below_rad), revolute_rows)
    r[valid_pri, :] *= scale
    r[valid_rev, :] *= scale
    J[valid_pri, :] *= scale
    J[valid_rev, :] *= scale
    if shift_invalid_to_threshold:
        r[torch.logical_and((r < -mjac_threshold_rad)[:, 0], invalid_rev)] += mjac_threshold_rad
        r[torch.logical_and((r > mjac_threshold_rad)[:, 0], invalid_rev)] -= mjac_threshold_rad
        r[torch.logical_and((r < -mjac_threshold_m)[:, 0], invalid_pri)] += mjac_threshold_m
        r[torch.logical_and((r > mjac_threshold_m)[:, 0], invalid_pri)] -= mjac_threshold_m
    return J, r, torch.logical_not(torch.logical_or(valid_pri, valid_rev))


def filter_rows_from_r_J_differencing(model: RobotModel, r: torch.Tensor, J: torch.Tensor, threshold_rad: float,
                                      threshold_m: float, shift_to_threshold: bool = True):
    """optimization_utils.py:736-768: keep only the rows whose |residual| exceeds the threshold of their joint type,
    optionally moved towards zero by the threshold."""
    assert r.shape[0] == J.shape[0] and r.shape[0] % model.ndof == 0
    revolute_idxs, prismatic_idxs = _row_masks(model, r.shape[0])
    keep = torch.logical_or(torch.logical_and(r.abs()[:, 0] > threshold_rad, revolute_idxs),
                            torch.logical_and(r.abs()[:, 0] > threshold_m, prismatic_idxs))
    if shift_to_threshold:
        r[torch.logical_and((r < -threshold_rad)[:, 0], revolute_idxs)] += threshold_rad
        r[torch.logical_and((r > threshold_rad)[:, 0], revolute_idxs)] -= threshold_rad
        r[torch.logical_and((r < -threshold_m)[:, 0], prismatic_idxs)] += threshold_m
        r[torch.logical_and((r > threshold_m)[:, 0], prismatic_idxs)] -= threshold_m
    return r[keep, :], J[keep, :]


def get_r_and_J(pms: LmParams, model: RobotModel, x: torch.Tensor, target_path: torch.Tensor,
                Tcuboids: Optional[List] = None, cuboids: Optional[List] = None):
    """optimization_utils.py:486-731.  Returns dicts of the per-term dense residuals / Jacobians (None when the
    term is off or has no active rows) in the reference's row order: pose, differencing, virtual, self, env."""
    n, ndof = x.shape
    dt = x.dtype
    r = {"pose": None, "differencing": None, "virtual_configs": None, "self_collisions": None, "env_collisions": None}
    J = dict(r)

    if pms.use_pose:  # :504-542
        J_fk = K.jacobian(model, x)
        Jp = torch.zeros((6 * n, ndof * n), dtype=dt)
        for i in range(n):
            Jp[6 * i : 6 * i + 6, i * ndof : (i + 1) * ndof] = J_fk[i]
        pose_errors, _ = get_6d_pose_errors(model, x, target_path)
        rp = pose_errors.flatten()[:, None].clone()
        if pms.pose_do_scale_down_satisfied:  # :513-531 (threshold units as in the reference: m and "deg" taken as rad)
            thr_m = pms.pose_ignore_satisfied_threshold_scale * pms.constraints[0] / 100
            thr_rad = pms.pose_ignore_satisfied_threshold_scale * pms.constraints[1]
            rp, Jp, _ = scale_down_rows_pose_below_error(rp, Jp, thr_m, thr_rad, pms.pose_ignore_satisfied_scale_down)
        rot_rows, pos_rows = _rotation_and_position_row_mask(n)
        rp[rot_rows, :] *= pms.alpha_rotation
        rp[pos_rows, :] *= pms.alpha_position
        Jp[rot_rows, :] *= pms.alpha_rotation
        Jp[pos_rows, :] *= pms.alpha_position
        r["pose"], J["pose"] = rp, Jp

    if pms.use_differencing:  # :550-612
        rd = angular_changes(x).reshape(((n - 1) * ndof, 1)).clone()
        zeros = torch.zeros((ndof * (n - 1), ndof * n), dtype=dt)
        Jd = torch.diagonal_scatter(zeros, torch.ones(ndof * (n - 1), dtype=dt), 0)
        Jd = torch.diagonal_scatter(Jd, -torch.ones(ndof * (n - 1), dtype=dt), offset=ndof)
        assert not (pms.differencing_do_scale_satisfied and pms.differencing_do_ignore_satisfied), "use one or the other"
        if pms.differencing_do_ignore_satisfied or pms.differencing_do_scale_satisfied:  # :560-567
            thr_rad = float(np.deg2rad(pms.constraints[2] - pms.differencing_ignore_satisfied_margin_deg))
            thr_m = (pms.constraints[3] - pms.differencing_ignore_satisfied_margin_cm) / 100
        if pms.differencing_do_ignore_satisfied:  # :570-578
            rd, Jd = filter_rows_from_r_J_differencing(model, rd, Jd, thr_rad, thr_m, shift_to_threshold=True)
        if pms.differencing_do_scale_satisfied:  # :581-593
            Jd, rd, _ = scale_down_rows_differencing_below_error(
                model, rd, Jd, thr_m, thr_rad, pms.differencing_scale_down_satisfied_scale,
                pms.differencing_scale_down_satisfied_shift_invalid_to_threshold)
        if model.has_prismatic_joints and not pms.differencing_do_ignore_satisfied:  # :606-609
            _, pris_rows = _row_masks(model, rd.shape[0])
            rd[pris_rows] *= pms.alpha_differencing_prismatic_scaling
            Jd[pris_rows] *= pms.alpha_differencing_prismatic_scaling
        r["differencing"] = pms.alpha_differencing * rd
        J["differencing"] = pms.alpha_differencing * Jd

    if pms.use_virtual_configs:  # :620-634, :430-483
        xv = pms.virtual_configs
        assert xv is not None and xv.shape == x.shape
        nv = pms.n_virtual_configs
        assert 2 * nv < n
        rv = torch.zeros((2 * nv * ndof, 1), dtype=dt)
        Jv = torch.zeros((ndof * 2 * nv, ndof * n), dtype=dt)
        for i in range(nv):
            rv[ndof * i : ndof * (i + 1), 0] = angular_subtraction(x[i, :], xv[i, :])
            Jv[i * ndof : (i + 1) * ndof, i * ndof : (i + 1) * ndof] = -torch.eye(ndof, dtype=dt)
        for i in range(nv):
            row = nv * ndof + ndof * i
            ci = n - nv + i
            rv[row : row + ndof, 0] = angular_subtraction(x[ci, :], xv[ci, :])
            col = n * ndof - ndof * nv + i * ndof
            Jv[row : row + ndof, col : col + ndof] = -torch.eye(ndof, dtype=dt)
        s = pms.alpha_virtual_configs * pms.alpha_differencing
        r["virtual_configs"], J["virtual_configs"] = s * rv, s * Jv

    if pms.use_self_collisions:  # :643-677
        dists, Jsc = G.self_collision_distances(model, x, with_jacobian=True)
        rs = (-pms.alpha_self_collision * dists).reshape(-1, 1)
        mask = (rs > 0).reshape(-1)
        r["self_collisions"] = rs[mask, :]
        if bool(mask.any()):
            Jfull = torch.block_diag(*torch.unbind(pms.alpha_self_collision * Jsc))
            J["self_collisions"] = Jfull[mask, :]

    if pms.use_env_collisions and Tcuboids is not None and len(Tcuboids) > 0:  # :682-725
        r_list, J_list = [], []
        for Tcuboid, cuboid in zip(Tcuboids, cuboids):
            dists, Jec = G.env_collision_distances(model, x, cuboid, Tcuboid, with_jacobian=True)
            re = (-pms.alpha_env_collision * dists).reshape(-1, 1)
            mask = (re > 0).reshape(-1)
            if int(mask.sum()) > 0:
                r_list.append(re[mask, :])
                Jfull = torch.block_diag(*torch.unbind(pms.alpha_env_collision * Jec))
                J_list.append(Jfull[mask, :])
        if len(r_list) > 0:
            r["env_collisions"] = torch.cat(r_list, dim=0)
            J["env_collisions"] = torch.cat(J_list, dim=0)
    return J, r


_ORDER = ("pose", "differencing", "virtual_configs", "self_collisions", "env_collisions")


def stack_rows(d) -> torch.Tensor:
    """LmResidual.get_r / LmJacobian.get_J (optimization_utils.py:89-102, :185-198)."""
    return torch.cat([d[k] for k in _ORDER if d[k] is not None and d[k].shape[0] > 0], dim=0)


def lm_full_step(J: torch.Tensor, r: torch.Tensor, x: torch.Tensor, lambd: float) -> torch.Tensor:
    """optimization.py:95-113"""
    n, ndof = x.shape
    eye = torch.eye(n * ndof, dtype=x.dtype)
    J_T = torch.transpose(J, 0, 1)
    A = torch.matmul(J_T, J) + lambd * eye
    b = torch.matmul(J_T, r)
    L = torch.linalg.cholesky(A, upper=False)
    y = torch.linalg.solve_triangular(L, b, upper=False)
    delta_x = torch.linalg.solve_triangular(L.T, y, upper=True).reshape((n, ndof))
    return x + delta_x


def levenberg_marquardt_full(model: RobotModel, x: torch.Tensor, target_path: torch.Tensor, pms: LmParams,
                             Tcuboids=None, cuboids=None) -> torch.Tensor:
    """optimization.py:116-144 for ONE path x [T,D]."""
    Jd, rd = get_r_and_J(pms, model, x, target_path, Tcuboids, cuboids)
    return lm_full_step(stack_rows(Jd), stack_rows(rd), x, pms.lm_lambda)


def run_fixed_schedule(model: RobotModel, x0: torch.Tensor, target_path: torch.Tensor, schedule: str,
                       Tcuboids=None, cuboids=None, pose_pms: LmParams = ALT_LOSS_V2_1_POSE,
                       diff_pms: LmParams = ALT_LOSS_V2_1_DIFF, all_pms: LmParams = ALL_TERMS) -> torch.Tensor:
    """The alternating loop of optimization.py:230-259 with a FIXED step sequence instead of the data-dependent
    branch: `schedule` is a string of 'p' (pose-only step), 'd' (differencing step with
    virtual_configs = x.clone(), optimization.py:253) and 'a' (all terms on); each step is followed by
    clamp_to_joint_limits (optimization.py:259).  x0 is [P*T, D] or [T, D]; paths are independent."""
    T = target_path.shape[0]
    x = x0.clone()
    P = x.shape[0] // T
    for step in schedule:
        if step == "p":
            tp = target_path if P == 1 else target_path.repeat(P, 1)
            x = levenberg_marquardt_only_pose(model, x, tp, pose_pms)
        else:
            base = diff_pms if step == "d" else all_pms
            outs = []
            for p in range(P):
                xp = x[p * T : (p + 1) * T]
                pms = replace(base, virtual_configs=xp.clone())
                outs.append(levenberg_marquardt_full(model, xp, target_path, pms, Tcuboids, cuboids))
            x = torch.cat(outs, dim=0)
        x = clamp_to_joint_limits(model, x)
    return x


# ---------------------------------------------------------------------------------------------------------
# validity metrics (evaluation_utils.py:29-75, :113-141; optimization_utils.py:836-923 without the klampt calls)


def calculate_pose_error_cm_deg(model: RobotModel, x: torch.Tensor, target_path: torch.Tensor):
    traced = K.forward_kinematics(model, x)
    pos = torch.norm(target_path[:, :3] - traced[:, :3], dim=1)
    rot = geodesic_distance_between_quaternions(target_path[:, 3:], traced[:, 3:])
    return 100 * pos, torch.rad2deg(rot)


def path_metrics(model: RobotModel, x: torch.Tensor, target_path: torch.Tensor, Tcuboids=None, cuboids=None):
    """Per-path summary used by `x_is_valid` and by the multi-GPU argmin: max position error (cm), max rotation
    error (deg), max revolute |dq| (deg), max prismatic |dq| (cm), trajectory length TL = sum |wrapped dq_revolute|
    (optimization.py:173-175), min capsule self distance, min capsule env distance."""
    err_cm, err_deg = calculate_pose_error_cm_deg(model, x, target_path)
    rev = model.revolute_joint_idxs
    pri = model.prismatic_joint_idxs
    dq_rev = angular_changes(x[:, rev])
    out = {
        "max_pos_cm": err_cm.max(),
        "max_rot_deg": err_deg.max(),
        "mjac_deg": torch.rad2deg(dq_rev.abs().max()),
        "mjac_cm": 100 * prismatic_changes(x[:, pri]).abs().max() if len(pri) > 0 else torch.zeros((), dtype=x.dtype),
        "tl": dq_rev.abs().sum(),
        "min_self": G.self_collision_distances(model, x).min(),
    }
    if Tcuboids:
        out["min_env"] = torch.stack(
            [G.env_collision_distances(model, x, c, T).min() for T, c in zip(Tcuboids, cuboids)]
        ).min()
    else:
        out["min_env"] = torch.tensor(float("inf"), dtype=x.dtype)
    return out


# ---------------------------------------------------------------------------------------------------------
# the alternating loop (optimization.py:147-373) and the validity check it calls (optimization_utils.py:836-923)


def errors_are_below_threshold(constraints, error_t_cm, error_R_deg, qdeltas_revolute_deg, qdeltas_prismatic_cm):
    """evaluation_utils.py:29-75: strict '<' on the maxima.  `constraints` = (max_allowed_position_error_cm,
    max_allowed_rotation_error_deg, max_allowed_mjac_deg, max_allowed_mjac_cm)."""
    pos_cm, rot_deg, mjac_deg, mjac_cm = constraints
    pose_pos_valid = bool(error_t_cm.max() < pos_cm)
    pose_rot_valid = bool(error_R_deg.max() < rot_deg)
    mjac_rev_valid = bool(qdeltas_revolute_deg.abs().max() < mjac_deg)
    mjac_pris_valid = bool(qdeltas_prismatic_cm.abs().max() < mjac_cm) if qdeltas_prismatic_cm.numel() > 0 else True
    return (pose_pos_valid and pose_rot_valid and mjac_rev_valid and mjac_pris_valid,
            (pose_pos_valid, pose_rot_valid, mjac_rev_valid, mjac_pris_valid))


def x_is_valid(model: RobotModel, constraints, target_path: torch.Tensor, x: torch.Tensor, Tcuboids=None, cuboids=None):
    """optimization_utils.py:836-923 for parallel_count == 1.  The klampt mesh checks of :889-900 (one
    `config_self_collides` / `config_collides_with_env` query per configuration, collision_detection.py:89-120) are
    answered by the capsule distances: a configuration collides when any capsule distance is negative - the check the
    CUDA loop performs.  -> (x or None, (pose_pos_valid, pose_rot_valid, mjac_rev_valid, mjac_pris_valid,
    is_a_self_collision, is_a_env_collision))"""
    error_t_cm, error_R_deg = calculate_pose_error_cm_deg(model, x, target_path)
    revolute_diffs_deg = torch.rad2deg(angular_changes(x[:, model.revolute_joint_idxs]))
    prismatic_diffs_cm = 100 * prismatic_changes(x[:, model.prismatic_joint_idxs])
    all_valid, flags = errors_are_below_threshold(constraints, error_t_cm, error_R_deg, revolute_diffs_deg,
                                                  prismatic_diffs_cm)
    is_a_self_collision = None
    is_a_env_collision = None
    if not all_valid:
        return None, (*flags, is_a_self_collision, is_a_env_collision)
    is_a_self_collision = bool((G.self_collision_distances(model, x).min(dim=1).values < 0).any())
    if is_a_self_collision:
        return None, (*flags, is_a_self_collision, is_a_env_collision)
    is_a_env_collision = False
    for Tc, c in zip(Tcuboids or [], cuboids or []):
        if bool((G.env_collision_distances(model, x, c, Tc).min(dim=1).values < 0).any()):
            is_a_env_collision = True
    if is_a_env_collision:
        return None, (*flags, is_a_self_collision, is_a_env_collision)
    return x, (*flags, is_a_self_collision, is_a_env_collision)


def run_lm_alternating_loss(model: RobotModel, x_seed: torch.Tensor, target_path: torch.Tensor, constraints,
                            max_n_steps: int, return_if_valid_after_n_steps: int, convergence_threshold: float,
                            Tcuboids=None, cuboids=None, params_diff: LmParams = ALT_LOSS_V2_1_DIFF,
                            params_pose: LmParams = ALT_LOSS_V2_1_POSE):
    """optimization.py:147-373 for ONE path, without the wall-clock exit (tmax_sec = infinity).
    -> (x_opt, n_steps_taken, is_valid, schedule) with schedule[i] = 'p' (levenberg_marquardt_only_pose, :258) or
    'd' (levenberg_marquardt_full with virtual_configs = x.clone(), :253-255) for iteration i."""

    def calc_TL(qpath):  # :173-175
        return float(angular_changes(qpath[:, model.revolute_joint_idxs]).abs().sum())

    x = x_seed.clone()
    tls_post_differencing = []
    last_valid = None
    last_valid_idx = -1
    pose_pos_valid, pose_rot_valid = True, False  # :219-220: the first step is pose-only
    converged = False
    schedule = ""
    i = 0
    for i in range(max_n_steps):  # :230
        took_differencing = pose_pos_valid and pose_rot_valid
        if took_differencing:  # :251-255
            pms = replace(params_diff, virtual_configs=x.clone())
            x_new = levenberg_marquardt_full(model, x, target_path, pms, Tcuboids, cuboids)
            schedule += "d"
        else:  # :258
            x_new = levenberg_marquardt_only_pose(model, x, target_path, params_pose)
            schedule += "p"
        x = clamp_to_joint_limits(model, x_new)  # :259
        tl_new = calc_TL(x)  # :268
        if took_differencing:  # :275-297
            if not converged and len(tls_post_differencing) > 0:
                if abs(tl_new - tls_post_differencing[-1]) < convergence_threshold:
                    converged = True
                    if last_valid_idx == i - 1:
                        break
            tls_post_differencing.append(tl_new)
        x_sol, (pose_pos_valid, pose_rot_valid, _, _, _, _) = x_is_valid(model, constraints, target_path, x, Tcuboids,
                                                                         cuboids)  # :318
        if x_sol is not None:  # :326-335
            last_valid_idx = i
            last_valid = x.clone()
            if converged:
                break
        if last_valid is not None:  # :346-358
            if i > return_if_valid_after_n_steps:
                break
            if i > max_n_steps:
                break
    x_return = last_valid if last_valid is not None else x
    return x_return, i, last_valid is not None, schedule
