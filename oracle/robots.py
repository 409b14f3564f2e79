"""Robot tables for the oracle (TEST INFRASTRUCTURE - see oracle/__init__.py).

The reference takes its robots from `jrl.robots` (`planners.py:10`, `planners.py:34-41`:
Panda, Fetch, FetchArm), which is not available offline (PARITY UNPINNED for this file).
The kinematic chains below are restated from the public URDFs jrl ships
(fetch_description `fetch.urdf`; franka_description `panda_arm_hand.urdf`), anchored on:
  * Fetch joint limits spelled out in tests/search_test.py:35-42,
  * Fetch joint 0 prismatic / Panda all revolute: tests/optimization_utils_test.py:67-119,
  * torso delta -> EE z delta: tests/optimization_utils_test.py:377-402,
  * torso_lift_link has identity orientation at q=0: data_type_utils.py:68-73,
  * FetchArm = Fetch with the torso fixed at 0: data_type_utils.py:155-158,
  * Panda FK point (tests/planners_test.py:282-309), frames `panda_link0`->`panda_hand`
    (ros2/ros2_publisher.py:60-61).
The collision capsules ([x1,y1,z1,x2,y2,z2,r] per link, in the link frame, the layout of
`robot._collision_capsules_by_link`, collision_detection.py:137) and the self-collision
pair list are NOT in the reference tree; the tables here are this repo's own approximation
of the URDF link geometry.  Pairs = all capsule pairs whose links are >= 3 apart in the chain.

These tables are deliberately written out a second time in
`cppflow_b200/csrc/robots.cuh`; tests/test_robot_tables.py checks both copies agree.
"""
from dataclasses import dataclass, field
from typing import List, Tuple
import math

PI = math.pi


@dataclass(frozen=True)
class ChainElement:
    name: str
    jtype: str  # "revolute" | "prismatic" | "fixed"
    xyz: Tuple[float, float, float]
    rpy: Tuple[float, float, float]
    axis: Tuple[float, float, float]
    lower: float
    upper: float
    child_link: str


@dataclass(frozen=True)
class Capsule:
    link: str
    frame: int  # 0 = base link frame, i = frame of chain[i-1].child_link
    p1: Tuple[float, float, float]
    p2: Tuple[float, float, float]
    radius: float
    sep: int = -1  # chain position used for the pair rule (defaults to `frame`)

    @property
    def sep_index(self) -> int:
        return self.frame if self.sep < 0 else self.sep


@dataclass
class RobotModel:
    name: str
    formal_robot_name: str
    base_link: str
    end_effector_link_name: str
    chain: List[ChainElement]
    capsules: List[Capsule]
    pair_min_separation: int = 3
    pairs: List[Tuple[int, int]] = field(default_factory=list)

    def __post_init__(self):
        if not self.pairs:
            n = len(self.capsules)
            self.pairs = [
                (i, j)
                for i in range(n)
                for j in range(i + 1, n)
                if self.capsules[j].sep_index - self.capsules[i].sep_index >= self.pair_min_separation
            ]

    @property
    def actuated(self) -> List[int]:
        return [i for i, e in enumerate(self.chain) if e.jtype != "fixed"]

    @property
    def ndof(self) -> int:
        return len(self.actuated)

    @property
    def actuated_joint_names(self) -> List[str]:
        return [self.chain[i].name for i in self.actuated]

    @property
    def actuated_joints_limits(self) -> List[Tuple[float, float]]:
        return [(self.chain[i].lower, self.chain[i].upper) for i in self.actuated]

    @property
    def prismatic_joint_idxs(self) -> List[int]:
        return [d for d, i in enumerate(self.actuated) if self.chain[i].jtype == "prismatic"]

    @property
    def revolute_joint_idxs(self) -> List[int]:
        return [d for d, i in enumerate(self.actuated) if self.chain[i].jtype == "revolute"]

    @property
    def has_prismatic_joints(self) -> bool:
        return len(self.prismatic_joint_idxs) > 0


_Z3 = (0.0, 0.0, 0.0)


def _fetch_chain(torso_fixed: bool) -> List[ChainElement]:
    X, Y, Z = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)
    return [
        ChainElement("torso_lift_joint", "fixed" if torso_fixed else "prismatic", (-0.086875, 0.0, 0.37743), _Z3, Z,
                     0.0, 0.38615, "torso_lift_link"),
        ChainElement("shoulder_pan_joint", "revolute", (0.119525, 0.0, 0.34858), _Z3, Z, -1.6056, 1.6056,
                     "shoulder_pan_link"),
        ChainElement("shoulder_lift_joint", "revolute", (0.117, 0.0, 0.06), _Z3, Y, -1.221, 1.518,
                     "shoulder_lift_link"),
        ChainElement("upperarm_roll_joint", "revolute", (0.219, 0.0, 0.0), _Z3, X, -PI, PI, "upperarm_roll_link"),
        ChainElement("elbow_flex_joint", "revolute", (0.133, 0.0, 0.0), _Z3, Y, -2.251, 2.251, "elbow_flex_link"),
        ChainElement("forearm_roll_joint", "revolute", (0.197, 0.0, 0.0), _Z3, X, -PI, PI, "forearm_roll_link"),
        ChainElement("wrist_flex_joint", "revolute", (0.1245, 0.0, 0.0), _Z3, Y, -2.16, 2.16, "wrist_flex_link"),
        ChainElement("wrist_roll_joint", "revolute", (0.1385, 0.0, 0.0), _Z3, X, -PI, PI, "wrist_roll_link"),
        ChainElement("gripper_axis", "fixed", (0.16645, 0.0, 0.0), _Z3, X, 0.0, 0.0, "gripper_link"),
    ]


_FETCH_CAPSULES = [
    Capsule("base_link", 0, (-0.02, 0.0, 0.15), (-0.02, 0.0, 0.22), 0.27),
    Capsule("torso_lift_link", 1, (0.0, 0.0, 0.05), (0.0, 0.0, 0.55), 0.10),
    Capsule("shoulder_pan_link", 2, (0.0, 0.0, 0.0), (0.117, 0.0, 0.06), 0.07),
    Capsule("shoulder_lift_link", 3, (0.0, 0.0, 0.0), (0.219, 0.0, 0.0), 0.065),
    Capsule("upperarm_roll_link", 4, (0.0, 0.0, 0.0), (0.133, 0.0, 0.0), 0.06),
    Capsule("elbow_flex_link", 5, (0.0, 0.0, 0.0), (0.197, 0.0, 0.0), 0.06),
    Capsule("forearm_roll_link", 6, (0.0, 0.0, 0.0), (0.1245, 0.0, 0.0), 0.055),
    Capsule("wrist_flex_link", 7, (0.0, 0.0, 0.0), (0.1385, 0.0, 0.0), 0.055),
    Capsule("wrist_roll_link", 8, (0.0, 0.0, 0.0), (0.08, 0.0, 0.0), 0.045),
    Capsule("gripper_link", 9, (-0.05, 0.0, 0.0), (0.03, 0.0, 0.0), 0.07),
]


def _panda_chain() -> List[ChainElement]:
    Z = (0.0, 0.0, 1.0)
    H = PI / 2
    return [
        ChainElement("panda_joint1", "revolute", (0.0, 0.0, 0.333), _Z3, Z, -2.8973, 2.8973, "panda_link1"),
        ChainElement("panda_joint2", "revolute", _Z3, (-H, 0.0, 0.0), Z, -1.7628, 1.7628, "panda_link2"),
        ChainElement("panda_joint3", "revolute", (0.0, -0.316, 0.0), (H, 0.0, 0.0), Z, -2.8973, 2.8973, "panda_link3"),
        ChainElement("panda_joint4", "revolute", (0.0825, 0.0, 0.0), (H, 0.0, 0.0), Z, -3.0718, -0.0698, "panda_link4"),
        ChainElement("panda_joint5", "revolute", (-0.0825, 0.384, 0.0), (-H, 0.0, 0.0), Z, -2.8973, 2.8973,
                     "panda_link5"),
        ChainElement("panda_joint6", "revolute", _Z3, (H, 0.0, 0.0), Z, -0.0175, 3.7525, "panda_link6"),
        ChainElement("panda_joint7", "revolute", (0.088, 0.0, 0.0), (H, 0.0, 0.0), Z, -2.8973, 2.8973, "panda_link7"),
        ChainElement("panda_joint8", "fixed", (0.0, 0.0, 0.107), _Z3, Z, 0.0, 0.0, "panda_link8"),
        ChainElement("panda_hand_joint", "fixed", _Z3, (0.0, 0.0, -PI / 4), Z, 0.0, 0.0, "panda_hand"),
    ]


_PANDA_CAPSULES = [
    Capsule("panda_link0", 0, (-0.09, 0.0, 0.06), (-0.06, 0.0, 0.06), 0.09),
    Capsule("panda_link1", 1, (0.0, 0.0, -0.30), (0.0, 0.0, -0.05), 0.07),
    Capsule("panda_link2", 2, (0.0, 0.0, -0.06), (0.0, 0.0, 0.06), 0.07),
    Capsule("panda_link3", 3, (0.0, 0.0, -0.22), (0.0, 0.0, -0.07), 0.07),
    Capsule("panda_link4", 4, (0.0, 0.0, -0.06), (0.0, 0.0, 0.06), 0.07),
    Capsule("panda_link5", 5, (0.0, 0.0, -0.30), (0.0, 0.06, -0.06), 0.065),
    Capsule("panda_link6", 6, (0.0, 0.0, -0.07), (0.0, 0.0, 0.01), 0.06),
    Capsule("panda_link7", 7, (0.0, 0.0, -0.06), (0.0, 0.0, 0.08), 0.05),
    # panda_hand is rigidly attached to panda_link7 (two fixed joints): counted as position 8 for the pair rule
    Capsule("panda_hand", 9, (0.0, -0.07, 0.04), (0.0, 0.07, 0.04), 0.05, sep=8),
]


def make_fetch() -> RobotModel:
    return RobotModel("fetch", "Fetch", "base_link", "gripper_link", _fetch_chain(False), list(_FETCH_CAPSULES))


def make_fetch_arm() -> RobotModel:
    return RobotModel("fetch_arm", "Fetch.Arm", "base_link", "gripper_link", _fetch_chain(True), list(_FETCH_CAPSULES))


def make_panda() -> RobotModel:
    return RobotModel("panda", "Panda", "panda_link0", "panda_hand", _panda_chain(), list(_PANDA_CAPSULES))


_FACTORIES = {"fetch": make_fetch, "fetch_arm": make_fetch_arm, "panda": make_panda}


def get_model(name: str) -> RobotModel:
    return _FACTORIES[name]()
