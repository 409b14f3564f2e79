"""Example client of the two services (reference cppflow/ros2/ros2_publisher.py:38-175): configure the scene for the
Panda, then ask for a plan along the beginning of the panda__1cube path, starting from a given configuration.

    ros2 run cppflow_b200 ros2_publisher        (or: python -m cppflow_b200.ros2.ros2_publisher)

The requests are built by plain functions that only need message FACTORIES (the rclpy message classes, or any callable
returning an attribute bag), so the same requests can be handed to service.CppFlowQueryService without a transport -
that is what tests/test_ros2_service.py does.  The rclpy node is import-guarded (no ROS2 in this image)."""
import warnings
from types import SimpleNamespace
from typing import Callable, List, Optional, Sequence

import torch

# the beginning of the panda__1cube problem (ros2_publisher.py:98-117): x, y, z, qw, qx, qy, qz before the offset
DUMMY_XYZ_OFFSET = (0.0, 0.5421984559194368, 0.7885155964931997)
DUMMY_TARGET_X = (0.45, 0.44547737, 0.44095477, 0.43643215, 0.43190953, 0.4273869, 0.42286432, 0.4183417, 0.41381907,
                  0.40929648, 0.40477386)


def dummy_target_path() -> List[List[float]]:
    """[11][7] rows x, y, z, qw, qx, qy, qz in the robot's base frame."""
    ox, oy, oz = DUMMY_XYZ_OFFSET
    return [[x + ox, oy, oz, 1.0, 0.0, 0.0, 0.0] for x in DUMMY_TARGET_X]


def get_initial_configuration(robot, target_pose: Sequence[float], device="cuda:0", positional_tolerance: float = 5e-5,
                              n_tries: int = 25, seed: int = 0) -> List[float]:
    """A configuration that reaches `target_pose` (x, y, z, qw, qx, qy, qz) within `positional_tolerance` metres and does
    not self-collide (ros2_publisher.py:17-35, which asks klampt's IK up to 25 times; here: LM pose steps on the GPU from
    random configurations inside the joint limits, capsule self-collision check).  Like the reference, no check against
    the obstacles of the scene."""
    from .. import ops
    from ..lm_hyper_parameters import ALT_LOSS_V2_1_POSE

    warnings.warn("No collision checking is performed against obstacles in the scene to find an initial configuration.")
    dev = torch.device(device)
    target = torch.tensor([list(map(float, target_pose))], dtype=torch.float32, device=dev)
    lim = torch.tensor(robot.actuated_joints_limits, dtype=torch.float32, device=dev)
    gen = torch.Generator().manual_seed(seed)
    prm = ops.make_params(ALT_LOSS_V2_1_POSE)
    u = torch.rand((n_tries, robot.ndof), generator=gen).to(dev)
    q = (lim[:, 0] + u * (lim[:, 1] - lim[:, 0])).contiguous()
    tgt = target.repeat(n_tries, 1).contiguous()
    for _ in range(60):
        q = ops.lm_pose_step(robot.robot_id, robot.ndof, prm, q, tgt, True)
    pose = robot.forward_kinematics(q)
    pos_err = (pose[:, :3] - target[:, :3]).norm(dim=1)
    rot_err = 1.0 - (pose[:, 3:] * target[:, 3:]).sum(dim=1).abs()
    self_colliding = ops.collision_flags(robot.robot_id, robot.ndof, q, None)[0].bool()
    ok = (pos_err < positional_tolerance) & (rot_err < 1e-6) & ~self_colliding
    if not bool(ok.any()):
        raise RuntimeError("Could not find collision free initial configuration")
    return [float(v) for v in q[int(ok.nonzero()[0])].cpu()]


def build_environment_request(request_factory: Callable = SimpleNamespace):
    """CppFlowEnvironmentConfig.Request of the example (ros2_publisher.py:56-71)."""
    request = request_factory()
    request.base_frame = "panda_link0"
    request.end_effector_frame = "panda_hand"
    request.jrl_robot_name = "panda"
    if not hasattr(request, "obstacles"):
        request.obstacles = []
    return request


def build_dummy_query(initial_configuration: Optional[Sequence[float]], request_factory: Callable = SimpleNamespace,
                      problem_factory: Callable = SimpleNamespace, pose_factory: Optional[Callable] = None,
                      joint_state_factory: Callable = SimpleNamespace):
    """CppFlowQuery.Request of the example (ros2_publisher.py:80-136): one problem of 11 waypoints, constraints
    0.1 cm / 1 deg / 2.5 deg / 0.5 cm, 3 s, the initial configuration when one is given."""
    if pose_factory is None:
        def pose_factory(x, y, z, qw, qx, qy, qz):
            return SimpleNamespace(position=SimpleNamespace(x=x, y=y, z=z), orientation=SimpleNamespace(x=qx, y=qy, z=qz, w=qw))
    request = request_factory()
    request.base_frame = "panda_link0"
    request.end_effector_frame = "panda_hand"
    request.jrl_robot_name = "panda"
    request.verbosity = 1
    request.max_planning_time_sec = 3.0
    request.anytime_mode_enabled = False
    request.max_allowed_position_error_cm = 0.1
    request.max_allowed_rotation_error_deg = 1.0
    request.max_allowed_mjac_deg = 2.5
    request.max_allowed_mjac_cm = 0.5
    problem = problem_factory()
    problem.waypoints = [pose_factory(*row) for row in dummy_target_path()]
    request.problems = [problem]
    request.initial_configuration_is_set = initial_configuration is not None
    request.initial_configuration = joint_state_factory(
        position=[float(v) for v in initial_configuration] if initial_configuration is not None else [])
    return request


def describe_response(response) -> List[str]:
    """The lines the example logs for a CppFlowQuery.Response (ros2_publisher.py:138-153)."""
    lines = ["Received CppFlowQuery.Response"]
    for i, (trajectory, success, error) in enumerate(zip(response.trajectories, response.success, response.errors)):
        lines.append(f"Problem {i}: Success = {success}, Error = {error}")
        lines.append(f"Trajectory {i}: {trajectory.joint_names}, {len(trajectory.points)} points")
        lines += [f"  {j}: {point.positions}" for j, point in enumerate(trajectory.points)]
    return lines


def _require_ros2():
    try:
        import rclpy
        from rclpy.node import Node
        from cppflow_msgs.srv import CppFlowQuery, CppFlowEnvironmentConfig
        from cppflow_msgs.msg import CppFlowProblem
        from geometry_msgs.msg import Pose, Point, Quaternion
        from sensor_msgs.msg import JointState
    except ImportError as e:
        raise ImportError("cppflow_b200.ros2.ros2_publisher needs a ROS2 environment (rclpy, geometry_msgs, sensor_msgs) with "
                          "the cppflow_msgs package; the request builders above work without it") from e
    return rclpy, Node, CppFlowQuery, CppFlowEnvironmentConfig, CppFlowProblem, Pose, Point, Quaternion, JointState


def make_client(device="cuda:0"):
    rclpy, Node, CppFlowQuery, CppFlowEnvironmentConfig, CppFlowProblem, Pose, Point, Quaternion, JointState = _require_ros2()
    from ..robot import get_robot

    class CppFlowQueryClient(Node):
        def __init__(self):
            super().__init__("cppflow_publisher")
            self.planning_client = self.create_client(CppFlowQuery, "/cppflow_planning_query")
            while not self.planning_client.wait_for_service(timeout_sec=1.0):
                self.get_logger().info("Waiting for service /cppflow_planning_query to be available...")
            self.scene_configuration_client = self.create_client(CppFlowEnvironmentConfig, "/cppflow_environment_configuration")
            while not self.scene_configuration_client.wait_for_service(timeout_sec=1.0):
                self.get_logger().info("Waiting for service /cppflow_environment_configuration to be available...")
            self.send_scene_configuration_request()
            self.send_dummy_problem_planning_request()

        def _call(self, client, request):
            future = client.call_async(request)
            rclpy.spin_until_future_complete(self, future)
            try:
                return future.result()
            except Exception as e:  # the reference logs and carries on
                self.get_logger().error(f"Service call failed: {str(e)}")
                return None

        def send_scene_configuration_request(self):
            response = self._call(self.scene_configuration_client, build_environment_request(CppFlowEnvironmentConfig.Request))
            if response is not None:
                self.get_logger().info(f"Received response: {response}")

        def send_dummy_problem_planning_request(self):
            q0 = get_initial_configuration(get_robot("panda"), dummy_target_path()[0], device=device)
            request = build_dummy_query(
                q0, CppFlowQuery.Request, CppFlowProblem,
                lambda x, y, z, qw, qx, qy, qz: Pose(position=Point(x=x, y=y, z=z), orientation=Quaternion(x=qx, y=qy, z=qz, w=qw)),
                JointState)
            self.get_logger().info(f"request.initial_configuration: {request.initial_configuration.position}")
            response = self._call(self.planning_client, request)
            if response is not None:
                for line in describe_response(response):
                    self.get_logger().info(line)

    return rclpy, CppFlowQueryClient


def main(args=None):
    rclpy, CppFlowQueryClient = make_client()
    rclpy.init(args=args)
    client = CppFlowQueryClient()
    client.destroy_node()
    rclpy.shutdown()


if __name__ == "__main__":
    main()
