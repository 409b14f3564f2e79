"""Data model of the hot path: the fields of the reference's dataclasses that search / collision / optimisation
read (cppflow/data_types.py: TimingData :27-50, Constraints :53-62, PlannerSettings :65-83, Problem :377-392,
PlannerResult) and `Plan` with its derived metrics (:86-348).  The reference fills a Plan's per-timestep collision
flags with klampt mesh queries (data_type_utils.py:244-276); here they come from the capsule kernels.  `PathReport`
is the planners' fast summary of the same quantities (one metrics kernel, one 8-float read)."""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from .config import (DEFAULT_RERUN_MJAC_THRESHOLD_CM, DEFAULT_RERUN_MJAC_THRESHOLD_DEG,
                     SUCCESS_THRESHOLD_initial_q_norm_dist)


@dataclass
class TimingData:
    total: float
    ikflow: float
    coll_checking: float
    batch_opt: float
    dp_search: float
    optimizer: float


@dataclass
class Constraints:
    max_allowed_position_error_cm: float
    max_allowed_rotation_error_deg: float
    max_allowed_mjac_deg: float
    max_allowed_mjac_cm: float

    @property
    def max_allowed_position_error_m(self):
        return self.max_allowed_position_error_cm / 100


# scripts/evaluate.py:51-56
DEFAULT_CONSTRAINTS = Constraints(
    max_allowed_position_error_cm=0.01, max_allowed_rotation_error_deg=0.1, max_allowed_mjac_deg=7.0,
    max_allowed_mjac_cm=2.0,
)


@dataclass
class PlannerSettings:
    k: int
    tmax_sec: float
    anytime_mode_enabled: bool
    latent_distribution: str = "uniform"
    latent_vector_scale: float = 2.0
    run_dp_search: bool = True
    do_rerun_if_optimization_fails: bool = False
    do_rerun_if_large_dp_search_mjac: bool = False
    rerun_mjac_threshold_deg: float = DEFAULT_RERUN_MJAC_THRESHOLD_DEG
    rerun_mjac_threshold_cm: float = DEFAULT_RERUN_MJAC_THRESHOLD_CM
    do_return_search_path_mjac: bool = False
    return_only_1st_plan: bool = False
    verbosity: int = 1

    def __post_init__(self):
        assert self.latent_distribution in {"uniform", "gaussian"}
        assert self.latent_vector_scale > 0.0


@dataclass
class Problem:
    constraints: Constraints
    target_path: torch.Tensor
    initial_configuration: Optional[torch.Tensor]
    robot: object
    name: str
    full_name: str
    obstacles: Optional[List]
    obstacles_Tcuboids: Optional[List]
    obstacles_cuboids: Optional[List]
    obstacles_klampt: List = field(default_factory=list)

    def __post_init__(self):
        norms = torch.norm(self.target_path[:, 3:7], dim=1)
        if float(norms.max()) > 1.01 or float(norms.min()) < 0.99:
            raise ValueError("quaternion(s) are not unit quaternion(s)")  # data_types.py:445-447
        if self.initial_configuration is not None:
            assert len(self.initial_configuration.shape) == 2, \
                f"'initial_configuration' should be [1, ndof], is {self.initial_configuration.shape}"
        self._obstacles_host = None

    @property
    def n_timesteps(self) -> int:
        return self.target_path.shape[0]

    @property
    def fancy_name(self) -> str:
        return f"{self.robot.formal_robot_name} - {self.name}"

    @property
    def obstacle_tables(self):
        """Host copy of the cuboid tables, built once per Problem."""
        from .ops import Obstacles

        if self._obstacles_host is None:
            self._obstacles_host = Obstacles(self.obstacles_cuboids or [], self.obstacles_Tcuboids or [])
        return self._obstacles_host

    @property
    def path_length_cumultive_positional_change_cm(self) -> float:
        return float(torch.norm(self.target_path[1:, 0:3] - self.target_path[:-1, 0:3], dim=1).sum()) * 100.0

    @property
    def path_length_cumulative_rotational_change_deg(self) -> float:
        """data_types.py:403-418: sum of the geodesic distances between consecutive target poses, minus the "imagined"
        error of the 1e-7 acos clamp (2 acos(1 - 1e-7) = 0.05 deg for every pair of identical orientations)."""
        from .evaluation_utils import geodesic_distance_between_quaternions

        q0, q1 = self.target_path[0:-1, 3:7], self.target_path[1:, 3:7]
        acos_clamp_epsilon = 1e-7
        dot = torch.sum(q0 * q1, dim=1)
        dot_is_1 = torch.logical_or(dot > 1 - acos_clamp_epsilon, dot < -1 + acos_clamp_epsilon)
        imagined_error_per_elem = 2 * torch.acos(torch.tensor([1 - acos_clamp_epsilon], device=self.target_path.device))
        total_imagined_error = imagined_error_per_elem * torch.sum(dot_is_1)
        rotational_changes = geodesic_distance_between_quaternions(q0, q1)
        return torch.rad2deg(rotational_changes.abs().sum() - total_imagined_error).item()


@dataclass
class Plan:
    """data_types.py:86-348: a joint-space path with its per-timestep errors and the metrics derived from them."""

    q_path: torch.Tensor
    q_path_revolute: torch.Tensor
    q_path_prismatic: torch.Tensor
    pose_path: torch.Tensor
    target_path: torch.Tensor
    robot_joint_limits: List[Tuple[float, float]]
    self_colliding_per_ts: torch.Tensor
    env_colliding_per_ts: torch.Tensor
    positional_errors: torch.Tensor
    rotational_errors: torch.Tensor
    provided_initial_configuration: Optional[torch.Tensor]
    constraints: Constraints

    def __post_init__(self):
        assert isinstance(self.q_path, torch.Tensor)
        assert self.q_path.shape == (self.target_path.shape[0], len(self.robot_joint_limits)), (
            f"Error: qpath.shape = {self.q_path.shape}, should be {(self.target_path.shape[0], len(self.robot_joint_limits))}")
        assert self.positional_errors.numel() == self.q_path.shape[0]
        assert self.rotational_errors.numel() == self.q_path.shape[0]

    # path length
    @property
    def path_length_rad(self) -> float:
        from .evaluation_utils import angular_changes

        return angular_changes(self.q_path_revolute).abs().sum().item()

    @property
    def path_length_m(self) -> float:
        from .evaluation_utils import prismatic_changes

        if self.q_path_prismatic.numel() > 0:
            return prismatic_changes(self.q_path_prismatic).abs().sum().item()
        return 0.0

    @property
    def is_a_prismatic_joint(self) -> bool:
        return self.q_path_prismatic.numel() > 0

    # rotational error
    @property
    def rotational_errors_deg(self):
        return torch.rad2deg(self.rotational_errors)

    @property
    def max_rotational_error_deg(self) -> float:
        return float(self.rotational_errors_deg.max())

    @property
    def mean_rotational_error_deg(self) -> float:
        return float(self.rotational_errors_deg.mean())

    # positional error
    @property
    def positional_errors_cm(self):
        return 100 * self.positional_errors

    @property
    def positional_errors_mm(self):
        return 1000 * self.positional_errors

    @property
    def max_positional_error_cm(self) -> float:
        return float(self.positional_errors_cm.max())

    @property
    def max_positional_error_mm(self) -> float:
        return self.max_positional_error_cm * 10.0

    @property
    def mean_positional_error_cm(self) -> float:
        return float(self.positional_errors_cm.mean())

    @property
    def mean_positional_error_mm(self) -> float:
        return self.mean_positional_error_cm * 10.0

    # mjac
    @property
    def mjac_per_timestep_deg(self):
        from .evaluation_utils import calculate_per_timestep_mjac_deg

        return calculate_per_timestep_mjac_deg(self.q_path_revolute)

    @property
    def mjac_deg(self) -> float:
        return self.mjac_per_timestep_deg.max().item()

    @property
    def mjac_per_timestep_cm(self):
        from .evaluation_utils import calculate_per_timestep_mjac_cm

        if self.q_path_prismatic.numel() == 0:
            return torch.zeros(self.target_path.shape[0] - 1, device=self.q_path.device, dtype=self.q_path.dtype)
        return calculate_per_timestep_mjac_cm(self.q_path_prismatic)

    @property
    def mjac_cm(self) -> float:
        if self.q_path_prismatic.numel() == 0:
            return 0.0
        return float(self.mjac_per_timestep_cm.max())

    # validity
    @property
    def joint_limits_violated(self) -> bool:
        from .evaluation_utils import joint_limits_exceeded

        return joint_limits_exceeded(self.robot_joint_limits, self.q_path)[0]

    @property
    def initial_q_norm_dist(self) -> float:
        if self.provided_initial_configuration is None:
            return 0.0
        return torch.norm(self.provided_initial_configuration.to(self.q_path.device) - self.q_path[0]).item()

    def is_valid_(self, verbose: bool = False):
        from .evaluation_utils import errors_are_below_threshold

        errs_below_thresh = errors_are_below_threshold(
            self.constraints.max_allowed_position_error_cm, self.constraints.max_allowed_rotation_error_deg,
            self.constraints.max_allowed_mjac_deg, self.constraints.max_allowed_mjac_cm, self.positional_errors_cm,
            self.rotational_errors_deg, self.mjac_per_timestep_deg, self.mjac_per_timestep_cm)[0]
        checks = {
            "joint limits in bounds": not self.joint_limits_violated,
            "errors_are_below_threshold(...)": bool(errs_below_thresh),
            "self.self_colliding_per_ts.sum() == 0": bool(self.self_colliding_per_ts.sum() == 0),
            "self.env_colliding_per_ts.sum() == 0": bool(self.env_colliding_per_ts.sum() == 0),
            "self.initial_q_norm_dist < SUCCESS_THRESHOLD_initial_q_norm_dist":
                self.initial_q_norm_dist < SUCCESS_THRESHOLD_initial_q_norm_dist,
        }
        iv = all(checks.values())
        if not verbose:
            return iv
        return iv, f"is_valid_ = {iv}\n" + "".join(f"{k}: {v}\n" for k, v in checks.items())

    @property
    def is_valid(self) -> bool:
        return self.is_valid_(verbose=False)

    def __str__(self) -> str:
        r = 5
        c = self.constraints
        return (
            "Plan {\n"
            f"  is_valid:                        {self.is_valid}\n"
            f"  mjac < {c.max_allowed_mjac_deg} deg:                  {self.mjac_deg < c.max_allowed_mjac_deg}\n"
            f"  mjac < {c.max_allowed_mjac_cm} cm:                   {self.mjac_cm < c.max_allowed_mjac_cm}\n"
            f"  max positional error < {10 * c.max_allowed_position_error_cm} mm:   "
            f"{self.max_positional_error_cm < c.max_allowed_position_error_cm}\n"
            f"  max rotational error < {c.max_allowed_rotation_error_deg} deg:  "
            f"{self.max_rotational_error_deg < c.max_allowed_rotation_error_deg}\n"
            f"  joint limits in bounds:          {not self.joint_limits_violated}\n"
            f"  close-to-initial-configuration:  {self.initial_q_norm_dist < SUCCESS_THRESHOLD_initial_q_norm_dist}\n"
            f"  # self collisions:               {int(self.self_colliding_per_ts.sum())}\n"
            f"  # env. collisions:               {int(self.env_colliding_per_ts.sum())}\n"
            "  .\n"
            f"  mjac:                  {round(self.mjac_deg, r)} deg\n"
            f"  mjac:                  {round(self.mjac_cm, r)} cm\n"
            f"  ave positional error:  {round(self.mean_positional_error_mm, r)} mm\n"
            f"  max positional error:  {round(self.max_positional_error_mm, r)} mm\n"
            f"  ave rotational error:  {round(self.mean_rotational_error_deg, r)} deg\n"
            f"  max rotational error:  {round(self.max_rotational_error_deg, r)} deg\n"
            f"  q_initial norm dist:   {round(self.initial_q_norm_dist, r)}\n"
            "  .\n"
            f"  trajectory length:     {round(self.path_length_rad, r)} rad\n"
            f"  trajectory length:     {round(self.path_length_m, r)} m\n"
            "}"
        )


@dataclass
class PlanNp:
    """data_types.py:351-366: every tensor attribute of a Plan as a numpy array."""

    plan: Plan

    def __getattribute__(self, attr):
        if attr == "plan":
            return super().__getattribute__("__dict__")["plan"]
        assert attr in dir(self.plan), f"Error: '{attr}' not found in Plan class"
        item = self.plan.__getattribute__(attr)
        if isinstance(item, torch.Tensor):
            return item.cpu().numpy()
        return item


@dataclass
class PathReport:
    """Capsule-based validity report of one joint-space path (the tensor part of Plan / x_is_valid)."""

    q_path: torch.Tensor
    max_pos_error_cm: float
    max_rot_error_deg: float
    mjac_deg: float
    mjac_cm: float
    trajectory_length_rad: float
    min_self_distance_m: float
    min_env_distance_m: float
    is_valid: bool
    initial_q_norm_dist: float = 0.0  # data_types.py:236-241: 0 when the problem has no initial configuration

    # the names the reference's Plan gives the same quantities (data_types.py:133-215)
    @property
    def max_positional_error_cm(self) -> float:
        return self.max_pos_error_cm

    @property
    def max_positional_error_mm(self) -> float:
        return self.max_pos_error_cm * 10.0

    @property
    def max_rotational_error_deg(self) -> float:
        return self.max_rot_error_deg

    @property
    def path_length_rad(self) -> float:
        return self.trajectory_length_rad


@dataclass
class PlannerResult:
    plan: PathReport
    timing: TimingData
    other_plans: List
    other_plans_names: List[str]
    debug_info: Dict
