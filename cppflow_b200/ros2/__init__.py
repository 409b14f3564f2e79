"""ROS2 front end of the planner (reference cppflow/ros2/).

`service.CppFlowQueryService` holds the two service callbacks of the reference's SubscriberNode
(ros2_subscriber.py:47-236) without any transport: requests and responses are duck-typed, so the logic runs (and is
tested) without rclpy.  `ros2_subscriber` binds it to rclpy when ROS2 is installed."""
