"""ROS2 front end: the service callbacks of cppflow/ros2/ros2_subscriber.py without a transport.
The CPU tests cover the request checks and the message conversions with duck-typed messages; the GPU test plans through
the service end to end."""
from types import SimpleNamespace as NS

import pytest
import torch

from cppflow_b200.ros2.ros2_utils import plan_to_ros_trajectory, waypoints_to_se3_sequence
from cppflow_b200.ros2.service import CppFlowQueryService


def _pose(row):
    x, y, z, qw, qx, qy, qz = [float(v) for v in row]
    return NS(position=NS(x=x, y=y, z=z), orientation=NS(w=qw, x=qx, y=qy, z=qz))


def _env_request(robot="fetch", ee="gripper_link", base="base_link", obstacles=()):
    return NS(jrl_robot_name=robot, end_effector_frame=ee, base_frame=base, obstacles=list(obstacles))


def _query(waypoints, n_problems=1, tmax=5.0, q0=None):
    return NS(problems=[NS(waypoints=waypoints) for _ in range(n_problems)], max_planning_time_sec=tmax, verbosity=0,
              anytime_mode_enabled=False, max_allowed_position_error_cm=0.01, max_allowed_rotation_error_deg=0.1,
              max_allowed_mjac_deg=7.0, max_allowed_mjac_cm=2.0, initial_configuration_is_set=q0 is not None,
              initial_configuration=NS(position=q0 if q0 is not None else []))


def test_message_conversions():
    rows = torch.tensor([[0.1, 0.2, 0.3, 1.0, 0.0, 0.0, 0.0], [0.4, 0.5, 0.6, 0.0, 1.0, 0.0, 0.0], [0.7, 0.8, 0.9, 0.5, 0.5, 0.5, 0.5]])
    se3 = waypoints_to_se3_sequence([_pose(r) for r in rows])
    assert se3.shape == (3, 7) and torch.equal(se3, rows)  # x y z qw qx qy qz (ros2_utils.py:22-33)
    robot = NS(ndof=2, actuated_joint_names=["a", "b"])
    traj = plan_to_ros_trajectory(NS(q_path=torch.tensor([[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]])), robot)
    assert traj.joint_names == ["a", "b"] and len(traj.points) == 3
    assert traj.points[2].positions == [4.0, 5.0] and traj.points[2].velocities == [0.0, 0.0]
    assert (traj.points[2].time_from_start.sec, traj.points[2].time_from_start.nanosec) == (2, 12)  # ros2_utils.py:48-49


def test_environment_setup_checks():
    svc = CppFlowQueryService(device="cpu")
    r = svc.environment_setup(_env_request(robot="ur5"), NS())
    assert r.success is False and "doesn't exist" in r.error
    r = svc.environment_setup(_env_request(ee="wrist_roll_link"), NS())
    assert r.success is False and "does not match the robot's end-effector link 'gripper_link" in r.error
    r = svc.environment_setup(_env_request(base="map"), NS())
    assert r.success is False and "does not match the robot's base link 'base_link" in r.error
    assert svc.planner is None
    bad = NS(position=NS(x=0.0, y=0.0, z=0.0), size=NS(x=0.1, y=-0.1, z=0.1))
    r = svc.environment_setup(_env_request(obstacles=[bad]), NS())
    assert r.success is False and "Malformed obstacle" in r.error
    r = svc.environment_setup(_env_request(), NS())
    assert r.success is True and svc.planner is not None and svc.planner.robot.name == "fetch"
    for name, ee, base in (("panda", "panda_hand", "panda_link0"), ("fetch_arm", "gripper_link", "torso_lift_link")):
        assert svc.environment_setup(_env_request(name, ee, base), NS()).success is True


def test_planning_query_checks():
    wp = [_pose([0.5, 0.0, 0.8, 1, 0, 0, 0])] * 4
    svc = CppFlowQueryService(device="cpu")
    r = svc.planning_query(_query(wp), NS())
    assert r.is_malformed_query and "Planner has not been configured" in r.malformed_query_error
    assert svc.environment_setup(_env_request(), NS()).success
    r = svc.planning_query(_query(wp, n_problems=2), NS())
    assert r.is_malformed_query and "Only 1 planning problem per query" in r.malformed_query_error
    r = svc.planning_query(_query(wp, tmax=0.0), NS())
    assert r.is_malformed_query and "Planning time is too short" in r.malformed_query_error
    r = svc.planning_query(_query(wp[:2]), NS())
    assert r.is_malformed_query and "At least 3 waypoints" in r.malformed_query_error
    r = svc.planning_query(_query(wp, q0=[0.0] * 5), NS())
    assert r.is_malformed_query and "8 actuated joints" in r.malformed_query_error


def test_rclpy_glue_fails_loudly_without_ros2():
    # rclpy is not part of this image; with ROS2 installed the node would be created instead
    try:
        import rclpy  # noqa: F401
    except ImportError:
        from cppflow_b200.ros2 import ros2_subscriber

        with pytest.raises(ImportError, match="needs a ROS2 environment"):
            ros2_subscriber.make_node()


@pytest.mark.gpu
def test_planning_query_end_to_end():
    from cppflow_b200.data_type_utils import problem_from_filename

    ref = problem_from_filename(None, "fetch__circle", device="cuda:0")
    obstacles = [NS(position=NS(x=o["x"], y=o["y"], z=o["z"]), size=NS(x=o["size_x"], y=o["size_y"], z=o["size_z"]))
                 for o in ref.obstacles]
    svc = CppFlowQueryService(device="cuda:0")
    assert svc.environment_setup(_env_request(obstacles=obstacles), NS()).success
    wp = [_pose(r) for r in ref.target_path.cpu().tolist()]
    r = svc.planning_query(_query(wp, tmax=30.0), NS())
    assert not getattr(r, "is_malformed_query", False)
    assert r.errors == [""] and r.success == [True]
    traj = r.trajectories[0]
    assert len(traj.points) == len(wp) and len(traj.points[0].positions) == 8
    assert traj.joint_names[0] == "torso_lift_joint"
    # the plan honours the obstacles of the environment message: no waypoint of it collides
    from cppflow_b200.collision_detection import qpaths_batched_env_collisions

    q = torch.tensor([p.positions for p in traj.points], device="cuda:0")[None]
    assert not bool(qpaths_batched_env_collisions(ref, q).any())
    # an initial configuration that touches a cuboid is refused (ros2_subscriber.py:196-199)
    cand = ref.robot.sample_joint_angles(4000, generator=torch.Generator().manual_seed(0), device="cuda:0")
    hit = qpaths_batched_env_collisions(ref, cand[None])[0]
    assert bool(hit.any())
    q0 = cand[int(torch.nonzero(hit)[0])].cpu().tolist()
    refused = svc.planning_query(_query(wp, tmax=30.0, q0=q0), NS())
    assert refused.is_malformed_query and refused.malformed_query_error == "Initial configuration is in collision with environment"


def test_publisher_request_builders():
    """cppflow/ros2/ros2_publisher.py:56-136: the example client's two requests, built without a transport."""
    from cppflow_b200.ros2.ros2_publisher import build_dummy_query, build_environment_request, describe_response, dummy_target_path

    env = build_environment_request()
    assert (env.jrl_robot_name, env.end_effector_frame, env.base_frame, env.obstacles) == ("panda", "panda_hand", "panda_link0", [])
    rows = dummy_target_path()
    assert len(rows) == 11 and rows[0][:3] == pytest.approx([0.45, 0.5421984559194368, 0.7885155964931997])
    q = build_dummy_query([0.1] * 7)
    assert len(q.problems) == 1 and len(q.problems[0].waypoints) == 11
    assert q.initial_configuration_is_set and q.initial_configuration.position == [0.1] * 7
    assert (q.max_allowed_position_error_cm, q.max_allowed_rotation_error_deg, q.max_allowed_mjac_deg, q.max_allowed_mjac_cm,
            q.max_planning_time_sec, q.anytime_mode_enabled) == (0.1, 1.0, 2.5, 0.5, 3.0, False)
    se3 = waypoints_to_se3_sequence(q.problems[0].waypoints)
    assert torch.allclose(se3, torch.tensor(rows))
    assert not build_dummy_query(None).initial_configuration_is_set
    resp = NS(trajectories=[NS(joint_names=["a"], points=[NS(positions=[1.0])])], success=[True], errors=[""])
    assert describe_response(resp) == ["Received CppFlowQuery.Response", "Problem 0: Success = True, Error = ",
                                       "Trajectory 0: ['a'], 1 points", "  0: [1.0]"]
    svc = CppFlowQueryService(device="cpu")
    assert svc.environment_setup(env, NS()).success is True  # the example's scene request is accepted as it is


@pytest.mark.gpu
def test_publisher_example_through_the_service():
    """The reference's example session (ros2 run cppflow ros2_subscriber / ros2_publisher) without ROS2: scene
    configuration for the Panda, an initial configuration that reaches the first waypoint to 5e-5 m without
    self-collision, the 11-waypoint query; the service answers with a valid 11-point trajectory that starts at the
    requested configuration."""
    from cppflow_b200.robot import get_robot
    from cppflow_b200.ros2.ros2_publisher import (build_dummy_query, build_environment_request, describe_response,
                                                  dummy_target_path, get_initial_configuration)

    dev = "cuda:0"
    svc = CppFlowQueryService(device=dev)
    assert svc.environment_setup(build_environment_request(), NS()).success is True
    robot = get_robot("panda")
    with pytest.warns(UserWarning):
        q0 = get_initial_configuration(robot, dummy_target_path()[0], device=dev)
    pose = robot.forward_kinematics(torch.tensor([q0], device=dev))[0].cpu()
    assert float((pose[:3] - torch.tensor(dummy_target_path()[0][:3])).norm()) < 5e-5
    resp = svc.planning_query(build_dummy_query(q0), NS(is_malformed_query=False, malformed_query_error=""))
    assert not resp.is_malformed_query, resp.malformed_query_error
    assert resp.success == [True], resp.errors
    traj = resp.trajectories[0]
    assert len(traj.points) == 11 and traj.joint_names == robot.actuated_joint_names
    start = torch.tensor(traj.points[0].positions)
    assert float((start - torch.tensor(q0)).norm()) < 0.02
    assert describe_response(resp)[1] == "Problem 0: Success = True, Error = "
