"""Message <-> tensor conversions of the ROS2 front end (reference cppflow/ros2/ros2_utils.py:12-52).

The reference builds `trajectory_msgs.msg.JointTrajectory` messages; that package only exists inside a ROS2 install, so
`plan_to_ros_trajectory` builds the real message when it can be imported and an attribute-compatible plain object
otherwise (same fields, same values)."""
from types import SimpleNamespace
from typing import List

import torch


def waypoints_to_se3_sequence(waypoints: List) -> torch.Tensor:
    """geometry_msgs/Pose-like objects (.position.x/y/z, .orientation.w/x/y/z) -> [N, 7] x y z qw qx qy qz
    (ros2_utils.py:12-34)."""
    rows = [[w.position.x, w.position.y, w.position.z, w.orientation.w, w.orientation.x, w.orientation.y, w.orientation.z]
            for w in waypoints]
    return torch.tensor(rows, dtype=torch.float32).reshape(len(rows), 7)


def _trajectory_types():
    try:  # inside a ROS2 environment: the real messages
        from trajectory_msgs.msg import JointTrajectory, JointTrajectoryPoint
        import rclpy.time

        return JointTrajectory, JointTrajectoryPoint, lambda: rclpy.time.Time().to_msg()
    except ImportError:
        def trajectory():
            return SimpleNamespace(header=SimpleNamespace(stamp=None), joint_names=[], points=[])

        def point():
            return SimpleNamespace(positions=[], velocities=[], time_from_start=SimpleNamespace(sec=0, nanosec=0))

        return trajectory, point, lambda: None


def plan_to_ros_trajectory(plan, robot):
    """Plan -> JointTrajectory: one point per waypoint, zero velocities, time_from_start = (i s, 12 ns) exactly as the
    reference fills it (ros2_utils.py:37-52)."""
    JointTrajectory, JointTrajectoryPoint, now = _trajectory_types()
    trajectory = JointTrajectory()
    trajectory.header.stamp = now()
    trajectory.joint_names = list(robot.actuated_joint_names)
    zero_velocity = [0.0] * robot.ndof
    q_path = plan.q_path.detach().cpu()  # one device -> host copy for the whole path
    for i in range(q_path.shape[0]):
        point = JointTrajectoryPoint()
        point.positions = q_path[i].tolist()
        point.velocities = list(zero_velocity)
        point.time_from_start.sec = i
        point.time_from_start.nanosec = 12
        trajectory.points.append(point)
    return trajectory
