"""`Robot`: the slice of jrl.robot.Robot the hot path reads, backed by the CUDA library.

Methods keep jrl's names, argument meaning and output shapes:
  forward_kinematics(x) [n,7]; jacobian(x) [n,6,ndof]; self_collision_distances(x) [n,S];
  self_collision_distances_jacobian(x) [n,S,ndof]; env_collision_distances(x, cuboid, Tcuboid) [n,C];
  env_collision_distances_jacobian(...) [n,C,ndof]; split_configs_to_revolute_and_prismatic(x).
The klampt-backed members of jrl.Robot (config_self_collides, config_collides_with_env,
forward_kinematics_klampt, _klampt_world_model) are out of scope and raise."""
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib, ops

_JOINT_NAMES = {
    "fetch": ["torso_lift_joint", "shoulder_pan_joint", "shoulder_lift_joint", "upperarm_roll_joint",
              "elbow_flex_joint", "forearm_roll_joint", "wrist_flex_joint", "wrist_roll_joint"],
    "fetch_arm": ["shoulder_pan_joint", "shoulder_lift_joint", "upperarm_roll_joint", "elbow_flex_joint",
                  "forearm_roll_joint", "wrist_flex_joint", "wrist_roll_joint"],
    "panda": [f"panda_joint{i}" for i in range(1, 8)],
}
_CAPSULE_LINKS = {
    "fetch": ["base_link", "torso_lift_link", "shoulder_pan_link", "shoulder_lift_link", "upperarm_roll_link",
              "elbow_flex_link", "forearm_roll_link", "wrist_flex_link", "wrist_roll_link", "gripper_link"],
    "panda": [f"panda_link{i}" for i in range(8)] + ["panda_hand"],
}
_CAPSULE_LINKS["fetch_arm"] = _CAPSULE_LINKS["fetch"]
_FORMAL = {"fetch": "Fetch", "fetch_arm": "Fetch.Arm", "panda": "Panda"}
_EE = {"fetch": "gripper_link", "fetch_arm": "gripper_link", "panda": "panda_hand"}


class Robot:
    name: str = None

    def __init__(self):
        assert self.name in ops.ROBOT_IDS, f"unknown robot '{self.name}'"
        self._rid = ops.ROBOT_IDS[self.name]
        info = ops._info(self._rid)
        self._ndof = int(info.ndof)
        self._n_pairs = int(info.n_pairs)
        self._n_capsules = int(info.n_capsules)
        # shortest decimal that round-trips in fp32: (0, 0.38615), (-1.6056, 1.6056), ... as in tests/search_test.py:35-42
        def dec(v):
            return float(np.format_float_positional(np.float32(v), unique=True))

        self._limits = [(dec(info.lower[d]), dec(info.upper[d])) for d in range(self._ndof)]
        self._prismatic = [d for d in range(self._ndof) if info.is_prismatic[d]]
        self._revolute = [d for d in range(self._ndof) if not info.is_prismatic[d]]
        self._collision_capsules_by_link: Dict[str, torch.Tensor] = {
            link: torch.tensor([float(v) for v in info.capsules[c]]) for c, link in enumerate(_CAPSULE_LINKS[self.name])
        }
        self._collision_pairs = [(int(info.pairs[p][0]), int(info.pairs[p][1])) for p in range(self._n_pairs)]

    # ---- properties read by search.py / optimization_utils.py / evaluation_utils.py
    @property
    def robot_id(self) -> int:
        return self._rid

    @property
    def ndof(self) -> int:
        return self._ndof

    @property
    def formal_robot_name(self) -> str:
        return _FORMAL[self.name]

    @property
    def actuated_joints_limits(self) -> List[Tuple[float, float]]:
        return self._limits

    @property
    def actuated_joint_names(self) -> List[str]:
        return _JOINT_NAMES[self.name]

    @property
    def prismatic_joint_idxs(self) -> List[int]:
        return self._prismatic

    @property
    def revolute_joint_idxs(self) -> List[int]:
        return self._revolute

    @property
    def has_prismatic_joints(self) -> bool:
        return len(self._prismatic) > 0

    @property
    def end_effector_link_name(self) -> str:
        return _EE[self.name]

    @property
    def n_self_collision_pairs(self) -> int:
        return self._n_pairs

    @property
    def n_collision_capsules(self) -> int:
        return self._n_capsules

    def __str__(self):
        return f"<Robot[{self.name}] ndof={self.ndof}>"

    # ---- jrl.Robot methods on the hot path
    def forward_kinematics(self, x: torch.Tensor, out_device=None, dtype=None) -> torch.Tensor:
        out = ops.forward_kinematics(self._rid, self._ndof, x)
        if dtype is not None and dtype != out.dtype:
            out = out.to(dtype)
        if out_device is not None and torch.device(out_device) != out.device:
            out = out.to(out_device)
        return out

    def jacobian(self, x: torch.Tensor) -> torch.Tensor:
        return ops.jacobian(self._rid, self._ndof, x)

    def self_collision_distances(self, x: torch.Tensor) -> torch.Tensor:
        return ops.self_collision_distances(self._rid, self._ndof, self._n_pairs, x)

    def self_collision_distances_jacobian(self, x: torch.Tensor) -> torch.Tensor:
        return ops.self_collision_distances(self._rid, self._ndof, self._n_pairs, x, with_jacobian=True)[1]

    def env_collision_distances(self, x: torch.Tensor, cuboid: torch.Tensor, Tcuboid: torch.Tensor) -> torch.Tensor:
        return ops.env_collision_distances(self._rid, self._ndof, self._n_capsules, x, ops.Obstacles([cuboid], [Tcuboid]))

    def env_collision_distances_jacobian(self, x: torch.Tensor, cuboid: torch.Tensor, Tcuboid: torch.Tensor) -> torch.Tensor:
        return ops.env_collision_distances(self._rid, self._ndof, self._n_capsules, x, ops.Obstacles([cuboid], [Tcuboid]),
                                           with_jacobian=True)[1]

    def split_configs_to_revolute_and_prismatic(self, x: torch.Tensor):
        return x[:, self._revolute], x[:, self._prismatic]

    def clamp_to_joint_limits(self, x: torch.Tensor) -> torch.Tensor:
        return ops.clamp_to_joint_limits_(self._rid, self._ndof, x)

    def sample_joint_angles(self, n: int, generator=None, device=None) -> torch.Tensor:
        lim = torch.tensor(self._limits, dtype=torch.float32)
        u = torch.rand((n, self._ndof), generator=generator)
        out = lim[:, 0] + u * (lim[:, 1] - lim[:, 0])
        return out.to(device) if device is not None else out

    # ---- klampt-backed members: out of scope
    def _klampt(self, *_a, **_k):
        raise NotImplementedError("klampt-backed Robot members are outside the B200 hot path (SURVEY.md 2, row 19)")

    config_self_collides = config_collides_with_env = forward_kinematics_klampt = _klampt


class Fetch(Robot):
    name = "fetch"


class FetchArm(Robot):
    name = "fetch_arm"


class Panda(Robot):
    name = "panda"


def get_robot(name: str) -> Robot:
    """jrl.robots.get_robot (data_type_utils.py:176)."""
    return {"fetch": Fetch, "fetch_arm": FetchArm, "panda": Panda}[name]()
