"""rclpy node serving /cppflow_planning_query and /cppflow_environment_configuration (reference
cppflow/ros2/ros2_subscriber.py:47-57, :238-247): transport only, the callbacks are service.CppFlowQueryService.

    ros2 run cppflow_b200 ros2_subscriber        (or: python -m cppflow_b200.ros2.ros2_subscriber)

Needs a ROS2 environment with the reference's `cppflow_msgs` package (CppFlowQuery, CppFlowEnvironmentConfig)."""
from .service import CppFlowQueryService

SAVE_MESSAGES = True


def _require_ros2():
    try:
        import rclpy
        from rclpy.node import Node
        from rclpy.serialization import serialize_message
        from cppflow_msgs.srv import CppFlowQuery, CppFlowEnvironmentConfig
    except ImportError as e:  # no ROS2 here: say what is missing instead of failing somewhere inside rclpy
        raise ImportError("cppflow_b200.ros2.ros2_subscriber needs a ROS2 environment (rclpy) with the cppflow_msgs package; "
                          "the service logic itself is cppflow_b200.ros2.service.CppFlowQueryService") from e
    return rclpy, Node, serialize_message, CppFlowQuery, CppFlowEnvironmentConfig


def make_node(device="cuda:0"):
    rclpy, Node, serialize_message, CppFlowQuery, CppFlowEnvironmentConfig = _require_ros2()

    class SubscriberNode(Node):
        def __init__(self):
            super().__init__("cppflow_query_server")
            self.service = CppFlowQueryService(device=device, log=lambda msg: self.get_logger().info(msg))
            self.srv = self.create_service(CppFlowQuery, "/cppflow_planning_query", self.planning_query_callback)
            self.environment_setup_srv = self.create_service(
                CppFlowEnvironmentConfig, "/cppflow_environment_configuration", self.environment_setup_callback)
            self.get_logger().info("CppFlowQuery service server started...")

        def _save(self, request, path):
            if SAVE_MESSAGES:
                with open(path, "wb") as file:
                    file.write(serialize_message(request))
                self.get_logger().info(f"Saved request to '{path}'")

        def environment_setup_callback(self, request, response):
            self._save(request, "/tmp/CppFlowEnvironmentConfig_request.bin")
            return self.service.environment_setup(request, response)

        def planning_query_callback(self, request, response):
            self._save(request, "/tmp/CppFlowQuery_request.bin")
            return self.service.planning_query(request, response)

    return rclpy, SubscriberNode


def main(args=None):
    rclpy, SubscriberNode = make_node()
    rclpy.init(args=args)
    node = SubscriberNode()
    try:
        rclpy.spin(node)
    except KeyboardInterrupt:
        pass
    finally:
        node.destroy_node()
        rclpy.shutdown()


if __name__ == "__main__":
    main()
