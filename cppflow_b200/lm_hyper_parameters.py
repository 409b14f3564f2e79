"""Hyper-parameters of the LM optimiser.

`OptimizationParameters` carries the field names of the reference's dataclass (cppflow/lm_hyper_parameters.py:14-80) so
that code written against it keeps working, but every field has a default ("term off / weight unset"): a parameter set
is written as the handful of values that differ.  The two live sets are the reference's ALT_LOSS_V2_1_DIFF (:86-118) and
ALT_LOSS_V2_1_POSE (:119-151); `all_terms_parameters()` is the fused iteration of BASELINE.json's north_star.
The CUDA kernels receive a parameter set as the POD `cppflow_lm_params` (ops.make_params)."""
import warnings
from dataclasses import dataclass, replace
from typing import Optional

import torch

ALTERNATING_LOSS_MAX_N_STEPS = 20
ALTERNATING_LOSS_RETURN_IF_SOL_FOUND_AFTER = 15
ALTERNATING_LOSS_CONVERGENCE_THRESHOLD = 0.3


@dataclass
class OptimizationParameters:
    seed_w_only_pose: Optional[bool] = None
    lm_lambda: float = 1e-6
    # residual weights; alpha_virtual_configs multiplies alpha_differencing
    alpha_position: Optional[float] = None
    alpha_rotation: Optional[float] = None
    alpha_differencing: Optional[float] = None
    alpha_differencing_prismatic_scaling: Optional[float] = None
    alpha_virtual_configs: Optional[float] = None
    alpha_self_collision: Optional[float] = None
    alpha_env_collision: Optional[float] = None
    # pose rows
    use_pose: bool = False
    pose_do_scale_down_satisfied: bool = False
    pose_ignore_satisfied_threshold_scale: Optional[float] = None
    pose_ignore_satisfied_scale_down: Optional[float] = None
    # joint-differencing rows
    use_differencing: bool = False
    differencing_do_ignore_satisfied: bool = False
    differencing_ignore_satisfied_margin_deg: Optional[float] = None
    differencing_ignore_satisfied_margin_cm: Optional[float] = None
    differencing_do_scale_satisfied: bool = False
    differencing_scale_down_satisfied_scale: Optional[float] = None
    differencing_scale_down_satisfied_shift_invalid_to_threshold: Optional[bool] = None
    # virtual configurations pinning the first / last n waypoints
    use_virtual_configs: bool = False
    virtual_configs: Optional[torch.Tensor] = None
    n_virtual_configs: Optional[int] = None
    # capsule collision rows
    use_self_collisions: bool = False
    use_env_collisions: bool = False

    def __post_init__(self):
        """The consistency rules of the reference (:56-80), one (condition, message) pair each."""
        ignore, scale = self.differencing_do_ignore_satisfied, self.differencing_do_scale_satisfied
        if scale and not self.use_virtual_configs:
            warnings.warn("differencing_do_scale_satisfied is True but virtual_configs are disabled")
        margins_ok = all(m is not None and m > 0 for m in (self.differencing_ignore_satisfied_margin_deg,
                                                           self.differencing_ignore_satisfied_margin_cm))
        rules = [
            (not (self.use_differencing and ignore and scale), "ignore- and scale-satisfied differencing exclude each other"),
            (margins_ok or not (ignore or scale), "satisfied-row options need positive margins (deg and cm)"),
            (not self.use_virtual_configs or self.virtual_configs is not None, "use_virtual_configs without virtual_configs"),
            (not self.use_virtual_configs or (isinstance(self.n_virtual_configs, int) and self.n_virtual_configs > 0),
             "n_virtual_configs must be a positive int"),
            (not self.use_self_collisions or (self.alpha_self_collision or 0) > 0, "alpha_self_collision must be > 0"),
            (not self.use_env_collisions or (self.alpha_env_collision or 0) > 0, "alpha_env_collision must be > 0"),
            (not self.pose_do_scale_down_satisfied or (isinstance(self.pose_ignore_satisfied_threshold_scale, float)
                                                      and self.pose_ignore_satisfied_threshold_scale > 0),
             "pose_ignore_satisfied_threshold_scale must be a positive float"),
        ]
        for ok, message in rules:
            assert ok, message


# joint-differencing steps: smooth the path, keep the ends pinned, push out of capsule collisions.
# (the reference notes that these weights expect dp_search's 1.5 deg / 3 cm joint-limit padding)
ALT_LOSS_V2_1_DIFF = OptimizationParameters(
    alpha_differencing=0.00375, alpha_differencing_prismatic_scaling=1.0,
    use_differencing=True,
    alpha_virtual_configs=1.0, use_virtual_configs=True, virtual_configs=torch.tensor([]), n_virtual_configs=4,
    alpha_self_collision=0.01, use_self_collisions=True,
    alpha_env_collision=0.01, use_env_collisions=True,
)

# pose-only steps: pull every waypoint back onto its target pose
ALT_LOSS_V2_1_POSE = OptimizationParameters(
    alpha_position=3.5, alpha_rotation=0.35, use_pose=True,
    differencing_scale_down_satisfied_shift_invalid_to_threshold=True,
)


def all_terms_parameters() -> OptimizationParameters:
    """Every residual term on (pose + differencing + virtual configs + self/env collisions): the 'fused' LM
    iteration of BASELINE.json's north_star, i.e. levenberg_marquardt_full with use_pose=True."""
    return replace(ALT_LOSS_V2_1_DIFF, use_pose=True, alpha_position=ALT_LOSS_V2_1_POSE.alpha_position,
                   alpha_rotation=ALT_LOSS_V2_1_POSE.alpha_rotation)
