"""The two service callbacks of the reference's SubscriberNode (cppflow/ros2/ros2_subscriber.py) as a plain class.

  environment_setup(request, response)   CppFlowEnvironmentConfig  (ros2_subscriber.py:59-107)
  planning_query(request, response)      CppFlowQuery              (ros2_subscriber.py:109-222)

Same checks, in the same order, with the same messages and response fields; requests / responses are whatever the
transport hands in (rclpy messages, or any object with those attributes).  Differences, all on the side of doing more:
the obstacles of the environment message are USED (the reference stores them and plans without them, its own TODO at
:166): cuboid obstacles reach the Problem and with them the capsule collision checks and the LM collision terms."""
import math
import traceback
from dataclasses import replace
from time import time
from typing import Callable, List, Optional

import torch

from ..collision_detection import qpaths_batched_env_collisions, qpaths_batched_self_collisions
from ..data_types import Constraints, PlannerSettings, Problem
from ..planners import CppFlowPlanner, Planner, PlannerSearcher
from ..robot import Robot, get_robot
from .ros2_utils import plan_to_ros_trajectory, waypoints_to_se3_sequence

PLANNERS = {"CppFlowPlanner": CppFlowPlanner, "PlannerSearcher": PlannerSearcher}
# ros2_subscriber.py:32-43
PLANNER_SETTINGS = {
    "CppFlowPlanner": PlannerSettings(k=175, tmax_sec=5.0, anytime_mode_enabled=False, do_rerun_if_large_dp_search_mjac=True,
                                      do_rerun_if_optimization_fails=True, do_return_search_path_mjac=True),
    "PlannerSearcher": PlannerSettings(k=175, tmax_sec=5.0, anytime_mode_enabled=False, verbosity=0),
}
PLANNER = "CppFlowPlanner"
_BASE_LINK = {"fetch": "base_link", "fetch_arm": "torso_lift_link", "panda": "panda_link0"}


def _cuboid_obstacles(obstacles, device):
    """Environment-message obstacles -> (problem-file style dicts, Tcuboids, cuboids) in the layout of
    data_type_utils.py:109-127.  An obstacle is anything with .position (x, y, z) and .size (x, y, z) - axis-aligned
    cuboids, the only obstacle type the reference's problems use; an optional .orientation (w, x, y, z) rotates it."""
    specs, Tcuboids, cuboids = [], [], []
    for ob in obstacles or []:
        px, py, pz = (float(getattr(ob.position, a)) for a in "xyz")
        sx, sy, sz = (float(getattr(ob.size, a)) for a in "xyz")
        assert min(sx, sy, sz) > 0, "obstacle sizes must be positive"
        Tc = torch.zeros((4, 4), dtype=torch.float32)
        Tc[:3, :3] = torch.eye(3)
        roll = pitch = yaw = 0.0
        q = getattr(ob, "orientation", None)
        if q is not None:
            w, x, y, z = (float(getattr(q, a)) for a in "wxyz")
            n = (w * w + x * x + y * y + z * z) ** 0.5
            assert n > 1e-9, "obstacle orientation must be a non-zero quaternion"
            w, x, y, z = w / n, x / n, y / n, z / n
            Tc[:3, :3] = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            roll = math.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
            pitch = math.asin(max(-1.0, min(1.0, 2 * (w * y - z * x))))
            yaw = math.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
        Tc[0, 3], Tc[1, 3], Tc[2, 3] = px, py, pz
        specs.append({"x": px, "y": py, "z": pz, "size_x": sx, "size_y": sy, "size_z": sz, "roll": roll, "pitch": pitch,
                      "yaw": yaw})
        Tcuboids.append(Tc.to(device))
        cuboids.append(torch.tensor([-sx / 2, -sy / 2, -sz / 2, sx / 2, sy / 2, sz / 2], dtype=torch.float32, device=device))
    return specs, Tcuboids, cuboids


class CppFlowQueryService:
    def __init__(self, device="cuda:0", log: Optional[Callable[[str], None]] = None, planner_name: str = PLANNER,
                 candidate_generator=None):
        self.device = torch.device(device)
        self.log = log if log is not None else (lambda _msg: None)
        self.planner_name = planner_name
        self.candidate_generator = candidate_generator
        self.robot: Optional[Robot] = None
        self.planner: Optional[Planner] = None
        self.obstacles: List = []

    # ---- /cppflow_environment_configuration -------------------------------------------------------------------------
    def environment_setup(self, request, response):
        t0 = time()
        self.log(f"Received a CppFlowEnvironmentConfig message: {request}")

        def specify_malformed_query(msg: str):
            response.success = False
            response.error = msg
            self.log(f"Returning response to malformed query: {response}")
            return response

        if (self.robot is None) or (self.robot.name != request.jrl_robot_name):
            try:
                t0_robot = time()
                self.robot = get_robot(request.jrl_robot_name)
                self.log(f"Loaded robot '{self.robot.name}' in {time() - t0_robot:.3f} seconds")
            except (KeyError, ValueError):
                return specify_malformed_query(f"Robot '{request.jrl_robot_name}' doesn't exist in the Jrl library")
        if self.robot.end_effector_link_name != request.end_effector_frame:
            return specify_malformed_query(
                f"The provided dnd-effector frame '{request.end_effector_frame}' does not match the robot's"
                f" end-effector link '{self.robot.end_effector_link_name}")
        robot_base_link_name = _BASE_LINK[self.robot.name]
        if robot_base_link_name != request.base_frame:
            return specify_malformed_query(
                f"The provided base frame '{request.base_frame}' does not match the robot's base link"
                f" '{robot_base_link_name}")
        try:
            _cuboid_obstacles(request.obstacles, "cpu")
        except (AssertionError, AttributeError, TypeError, ValueError) as e:
            return specify_malformed_query(f"Malformed obstacle: {e}")
        self.obstacles = list(request.obstacles or [])
        self.planner = PLANNERS[self.planner_name](PLANNER_SETTINGS[self.planner_name], self.robot, self.candidate_generator)
        response.success = True
        self.log(f"Returning response: {response} ({time() - t0:.3f} seconds)")
        return response

    # ---- /cppflow_planning_query --------------------------------------------------------------------------------------
    def planning_query(self, request, response):
        t0 = time()

        def specify_malformed_query(msg: str):
            response.is_malformed_query = True
            response.malformed_query_error = msg
            self.log(f"Returning response: {response} for malformed query")
            return response

        if len(request.problems) != 1:
            return specify_malformed_query(
                f"Only 1 planning problem per query currently supported ({len(request.problems)} problems provided)")
        if request.max_planning_time_sec < 1e-6:
            return specify_malformed_query(
                f"Planning time is too short (`max_planning_time_sec`: {request.max_planning_time_sec})")
        if self.planner is None:
            return specify_malformed_query(
                "Planner has not been configured. Send a 'CppFlowEnvironmentConfig' message on the"
                " '/cppflow_environment_configuration' topic to configure the scene first.")
        request_problem = request.problems[0]
        if len(request_problem.waypoints) < 3:
            return specify_malformed_query(
                f"At least 3 waypoints are required per problem (only {len(request_problem.waypoints)} provided)")

        ndof = self.planner.robot.ndof
        # a copy per query: the reference mutates its module-level settings object in place
        settings = replace(PLANNER_SETTINGS[self.planner_name], tmax_sec=0.9 * request.max_planning_time_sec,
                           verbosity=request.verbosity, anytime_mode_enabled=request.anytime_mode_enabled)
        constraints = Constraints(
            max_allowed_position_error_cm=request.max_allowed_position_error_cm,
            max_allowed_rotation_error_deg=request.max_allowed_rotation_error_deg,
            max_allowed_mjac_deg=request.max_allowed_mjac_deg,
            max_allowed_mjac_cm=request.max_allowed_mjac_cm,
        )
        self.planner.set_settings(settings)
        q0 = None
        if request.initial_configuration_is_set:
            q0 = torch.tensor(list(request.initial_configuration.position), dtype=torch.float32, device=self.device)
            if q0.numel() != ndof:
                return specify_malformed_query(
                    f"Initial configuration has {q0.numel()} joint positions, the robot has {ndof} actuated joints")
            q0 = q0.view(1, ndof)
        try:
            specs, Tcuboids, cuboids = _cuboid_obstacles(self.obstacles, self.device)
            problem = Problem(
                constraints,
                target_path=waypoints_to_se3_sequence(request_problem.waypoints).to(self.device),
                initial_configuration=q0,
                robot=self.robot,
                name="ros2-queried-problem",
                full_name="ros2-queried-problem",
                obstacles=specs,
                obstacles_Tcuboids=Tcuboids,
                obstacles_cuboids=cuboids,
                obstacles_klampt=[],
            )
            n = problem.n_timesteps
            self.log(f"target-path cumulative positional-change, cm:         {problem.path_length_cumultive_positional_change_cm}")
            self.log(f"target-path cumulative rotational-change, deg:        {problem.path_length_cumulative_rotational_change_deg}")
            self.log(f"target-path mean positional change per waypoint, cm:  {problem.path_length_cumultive_positional_change_cm / n}")
            self.log(f"target-path mean rotational change per waypoint, deg: {problem.path_length_cumulative_rotational_change_deg / n}")
        except (AssertionError, ValueError) as e:
            return specify_malformed_query(f"Creating 'Problem' dataclass failed: {e}")

        if q0 is not None:
            if qpaths_batched_env_collisions(problem, q0.view(1, 1, ndof)).item():
                return specify_malformed_query("Initial configuration is in collision with environment")
            if qpaths_batched_self_collisions(problem, q0.view(1, 1, ndof)).item():
                return specify_malformed_query("Initial configuration is self-colliding")

        try:
            planning_result = self.planner.generate_plan(problem)
        except (RuntimeError, AttributeError, AssertionError) as e:
            tb = traceback.extract_tb(e.__traceback__)[-1]
            error_msg = f"{e} (File: {tb.filename}, Line: {tb.lineno})"
            response.trajectories = []
            response.success = [False]
            response.errors = [error_msg]
            self.log(f"Planning failed with exception: '{error_msg}'")
            return response

        plan = planning_result.plan
        response.trajectories = [plan_to_ros_trajectory(plan, self.robot)]
        response.success = [bool(plan.is_valid)]
        response.errors = [""]
        self.log(f"Planning complete - returning {sum(response.success)} / {len(response.trajectories)} successful"
                 f" trajectories ({time() - t0:.3f} seconds)")
        self.log(f"{planning_result.plan}")
        self.log(f"{planning_result.timing}")
        return response
