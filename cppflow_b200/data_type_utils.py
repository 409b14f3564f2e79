"""Problem loading with the reference's entry point (cppflow/data_type_utils.py: problem_from_filename :148-219,
get_obstacles :87-145, offset_target_path :55-84, problem lists :24-52).

The problem yamls and path csvs of the reference are packed into cppflow_b200/data/{problems.json,
target_paths.npz} by data/make_problem_data.py (the GPU box has no copy of the reference tree).  klampt is not
needed: the only link pose the offsets use is `torso_lift_link` at q = 0, which is the fixed origin of the torso
joint (identity orientation, data_type_utils.py:73)."""
import json
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import config
from .data_types import Constraints, Problem, DEFAULT_CONSTRAINTS
from .robot import Robot, get_robot

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

ALL_PROBLEM_FILENAMES = [
    "fetch_arm__hello", "fetch_arm__circle", "fetch_arm__rot_yz2", "fetch_arm__s", "fetch_arm__square",
    "fetch__circle", "fetch__hello", "fetch__rot_yz2", "fetch__s", "fetch__square",
    "panda__1cube", "panda__2cubes", "panda__flappy_bird",
]
ALL_OBS_PROBLEM_FILENAMES = [
    "fetch_arm__circle", "fetch_arm__s", "fetch_arm__square", "fetch__circle", "fetch__s", "fetch__square",
    "panda__1cube", "panda__2cubes", "panda__flappy_bird",
]

# world position of the frames `path_offset_frame` may name, at q = 0
_FRAME_ORIGIN_AT_ZERO = {"world": (0.0, 0.0, 0.0), "torso_lift_link": (-0.086875, 0.0, 0.37743)}

_problems_cache = None
_paths_cache = None


def _problems() -> Dict:
    global _problems_cache
    if _problems_cache is None:
        with open(os.path.join(_DATA, "problems.json")) as f:
            _problems_cache = json.load(f)
    return _problems_cache


def _paths():
    global _paths_cache
    if _paths_cache is None:
        _paths_cache = np.load(os.path.join(_DATA, "target_paths.npz"))
    return _paths_cache


def _quat_to_R(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def _R_to_quat(R):
    t = np.array([1 + R[0, 0] + R[1, 1] + R[2, 2], 1 + R[0, 0] - R[1, 1] - R[2, 2],
                  1 - R[0, 0] + R[1, 1] - R[2, 2], 1 - R[0, 0] - R[1, 1] + R[2, 2]])
    i = int(np.argmax(t))
    if i == 0:
        q = np.array([t[0], R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    elif i == 1:
        q = np.array([R[2, 1] - R[1, 2], t[1], R[1, 0] + R[0, 1], R[0, 2] + R[2, 0]])
    elif i == 2:
        q = np.array([R[0, 2] - R[2, 0], R[1, 0] + R[0, 1], t[2], R[1, 2] + R[2, 1]])
    else:
        q = np.array([R[1, 0] - R[0, 1], R[2, 0] + R[0, 2], R[2, 1] + R[1, 2], t[3]])
    q = q / (2 * np.sqrt(t[i]))
    return q if q[0] >= 0 else -q


def offset_target_path(target_path: np.ndarray, path_offset_frame: str, xyz_offset: List[float],
                       R_offset: List[List[float]]) -> torch.Tensor:
    """data_type_utils.py:55-84: translate by xyz_offset + frame origin; right-multiply each pose's rotation."""
    path = target_path.copy()
    frame_origin = _FRAME_ORIGIN_AT_ZERO[path_offset_frame]
    for i in range(3):
        path[:, i] += xyz_offset[i] + frame_origin[i]
    R_off = np.array(R_offset, dtype=np.float64)
    if not np.allclose(R_off, np.eye(3)):
        for i in range(path.shape[0]):
            path[i, 3:7] = _R_to_quat(_quat_to_R(path[i, 3:7]) @ R_off)
    return torch.tensor(path, dtype=torch.float32)


def get_obstacles(problem_dict: Dict, device=None):
    """data_type_utils.py:87-145: cuboid = [-sx/2,-sy/2,-sz/2, sx/2,sy/2,sz/2]; Tcuboid 4x4 with Tcuboid[3,3] left 0."""
    device = config.DEVICE if device is None else device
    obstacles, Tcuboids, cuboids = [], [], []
    for obs in problem_dict.get("obstacles", []):
        obs = dict(obs)
        obs["x"] += problem_dict["obstacle_xyz_offset"][0]
        obs["y"] += problem_dict["obstacle_xyz_offset"][1]
        obs["z"] += problem_dict["obstacle_xyz_offset"][2]
        assert abs(obs["roll"]) < 1e-8 and abs(obs["pitch"]) < 1e-8 and abs(obs["yaw"]) < 1e-8
        cuboids.append(torch.tensor([-obs["size_x"] / 2, -obs["size_y"] / 2, -obs["size_z"] / 2,
                                     obs["size_x"] / 2, obs["size_y"] / 2, obs["size_z"] / 2], device=device))
        Tc = torch.zeros((4, 4))
        Tc[:3, :3] = torch.eye(3)
        Tc[0, 3], Tc[1, 3], Tc[2, 3] = obs["x"], obs["y"], obs["z"]
        Tcuboids.append(Tc.to(device))
        obstacles.append(obs)
    return obstacles, Tcuboids, cuboids


def problem_dict_from_yaml(filepath: str) -> Dict:
    """Normalise one problem yaml (the reference's format, problems/*.yaml: `obstacles` is a list of lists of
    single-key maps) into the dict layout of data/problems.json."""
    import yaml

    with open(filepath) as f:
        d = yaml.load(f, Loader=yaml.FullLoader)
    for key in ("robot", "path_name", "path_offset_frame", "path_xyz_offset", "path_R_offset"):
        assert key in d, f"problem file '{filepath}' has no '{key}'"
    obstacles = []
    for obs in d.get("obstacles", []) or []:
        parsed = {}
        for item in ([obs] if isinstance(obs, dict) else obs):
            parsed.update(item)
        obstacles.append({k: float(v) for k, v in parsed.items()})
    out = {
        "robot": d["robot"],
        "path_name": d["path_name"],
        "path_offset_frame": d["path_offset_frame"],
        "path_xyz_offset": [float(v) for v in d["path_xyz_offset"]],
        "path_R_offset": [[float(v) for v in row] for row in d["path_R_offset"]],
        "obstacle_xyz_offset": [float(v) for v in d.get("obstacle_xyz_offset", [0, 0, 0])],
        "source": filepath,
    }
    if "obstacles" in d:
        out["obstacles"] = obstacles
    return out


def target_path_from_csv(filepath: str) -> np.ndarray:
    """A path csv of the reference (paths/*.csv: header row, then time,x,y,z,qw,qx,qy,qz) -> float64 [T, 7]."""
    import csv

    with open(filepath) as f:
        rows = [[float(x) for x in row] for i, row in enumerate(csv.reader(f, delimiter=",")) if i > 0 and row]
    path = np.array(rows, dtype=np.float64)
    assert path.ndim == 2 and path.shape[1] == 8, f"'{filepath}': expected 8 columns (time, xyz, wxyz), got {path.shape}"
    return path[:, 1:]


def _named_target_path(path_name: str, search_from: Optional[str]) -> np.ndarray:
    """The packed path of that name, else `<path_name>.csv` beside the problem file, in its `paths/`, or in `../paths/`
    (the reference keeps problems/ and paths/ as siblings)."""
    paths = _paths()
    if path_name in paths:
        return paths[path_name]
    tried = []
    if search_from is not None:
        here = os.path.dirname(os.path.abspath(search_from))
        for folder in (here, os.path.join(here, "paths"), os.path.join(here, os.pardir, "paths")):
            candidate = os.path.join(folder, path_name + ".csv")
            tried.append(candidate)
            if os.path.isfile(candidate):
                return target_path_from_csv(candidate)
    raise FileNotFoundError(f"no target path '{path_name}': not a packed path {sorted(paths.keys())} and not at {tried}")


def problem_from_filename(constraints: Optional[Constraints], problem_filename: str, filepath_override: Optional[str] = None,
                          robot: Optional[Robot] = None, device=None) -> Problem:
    """data_type_utils.py:148-219.  Build a Problem (target path on `device`) from one of the packed problem
    definitions, or, with `filepath_override`, from a problem yaml in the reference's format (its target path is a
    packed path or a csv next to the yaml).  As in the reference (:184), a caller-provided robot excludes obstacles."""
    device = config.DEVICE if device is None else device
    if filepath_override is None:
        assert "yaml" not in problem_filename, "problem_filename should not include the .yaml file extension"
        assert problem_filename in _problems(), f"unknown problem '{problem_filename}' (known: {sorted(_problems())})"
        d = _problems()[problem_filename]
    else:
        d = problem_dict_from_yaml(filepath_override)
    if robot is None:
        robot = get_robot(d["robot"])
    elif filepath_override is not None:
        assert "obstacles" not in d, f"Error - obstacles found for {problem_filename} but robot is provided"
    obstacles, Tcuboids, cuboids = get_obstacles(d, device=device)
    target_path = offset_target_path(_named_target_path(d["path_name"], filepath_override), d["path_offset_frame"],
                                     d["path_xyz_offset"], d["path_R_offset"]).to(device)
    return Problem(constraints if constraints is not None else DEFAULT_CONSTRAINTS, target_path, None, robot,
                   d["path_name"], problem_filename, obstacles, Tcuboids, cuboids, [])


def get_all_problems(device=None) -> List[Problem]:
    return [problem_from_filename(None, name, device=device) for name in ALL_PROBLEM_FILENAMES]


def plan_from_qpath(qpath: torch.Tensor, problem: Problem):
    """data_type_utils.py:244-276: FK, per-timestep errors and collision flags of a joint-space path -> Plan.
    The reference asks klampt's mesh checker for the per-timestep collisions ("a tighter bound" than the capsules, :252);
    klampt is out of scope, so the flags are the capsule kernels' (d < 0 for any pair / capsule-cuboid test)."""
    from .collision_detection import qpaths_batched_collisions
    from .data_types import Plan
    from .evaluation_utils import positional_errors, rotational_errors

    assert isinstance(qpath, torch.Tensor), f"qpath must be a torch.Tensor, got {type(qpath)}"
    robot = problem.robot
    traced_path = robot.forward_kinematics(qpath)
    qpath_revolute, qpath_prismatic = robot.split_configs_to_revolute_and_prismatic(qpath)
    self_colliding, env_colliding = qpaths_batched_collisions(problem, qpath[None].contiguous())
    self_colliding, env_colliding = self_colliding[0], env_colliding[0]
    if config.SELF_COLLISIONS_IGNORED:
        self_colliding = torch.zeros_like(self_colliding)
    if config.ENV_COLLISIONS_IGNORED:
        env_colliding = torch.zeros_like(env_colliding)
    return Plan(
        q_path=qpath, q_path_revolute=qpath_revolute, q_path_prismatic=qpath_prismatic, pose_path=traced_path,
        target_path=problem.target_path, robot_joint_limits=robot.actuated_joints_limits,
        self_colliding_per_ts=self_colliding, env_colliding_per_ts=env_colliding,
        positional_errors=positional_errors(traced_path, problem.target_path),
        rotational_errors=rotational_errors(traced_path, problem.target_path),
        provided_initial_configuration=problem.initial_configuration, constraints=problem.constraints,
    )
