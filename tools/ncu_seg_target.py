"""ncu target: one assemble + segmented solve per (paths, segments) pair (PATHS, SEGS env), after one warm-up each."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
dev = torch.device("cuda:0")
robot = get_robot("fetch"); T, D = 300, 8
problem = synthetic_problem(robot, T, device=dev)
PATHS = [int(s) for s in os.environ.get("PATHS", "1,1024").split(",")]
SEGS = [int(s) for s in os.environ.get("SEGS", "0,8,16").split(",")]
_, xh = synthetic_seeds_host(robot, max(PATHS), T)
x0 = xh.to(dev)
prm = ops.make_params(all_terms_parameters())
for P in PATHS:
    for S in SEGS:
        for _ in range(2):
            ops.lm_full_step(robot.robot_id, D, prm, x0[:P * T], None, problem.target_path, P, T, problem.obstacle_tables, True, segments=S)
        torch.cuda.synchronize()
