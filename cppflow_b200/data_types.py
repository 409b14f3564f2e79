"""Data model of the hot path: the fields of the reference's dataclasses that search / collision / optimisation
read (cppflow/data_types.py: TimingData :27-50, Constraints :53-62, PlannerSettings :65-83, Problem :377-392,
PlannerResult).  `Plan` and its klampt-backed validity report are out of scope; `PathReport`
carries the capsule-based metrics the CUDA path computes instead."""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .config import DEFAULT_RERUN_MJAC_THRESHOLD_CM, DEFAULT_RERUN_MJAC_THRESHOLD_DEG


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


@dataclass
class PlannerResult:
    plan: PathReport
    timing: TimingData
    other_plans: List
    other_plans_names: List[str]
    debug_info: Dict
