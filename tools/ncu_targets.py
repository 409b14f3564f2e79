"""Launches every kernel of the path once at the shapes the bench quotes (for `ncu -k regex:... ` captures):
8192 x 300 Fetch for the batch kernels, fetch__circle sizes (k = 175 / 300, T = 295) for dp_search, one path for the
resident solve and the cluster metrics kernel, 1024 paths for the register-resident solve."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters, ALT_LOSS_V2_1_DIFF, ALT_LOSS_V2_1_POSE

dev = torch.device("cuda:0")
robot = get_robot("fetch"); P, T, D = 8192, 300, 8
problem = synthetic_problem(robot, T, device=dev)
x0 = synthetic_seeds_host(robot, P, T)[1].to(dev); xo = torch.empty_like(x0)
ob, rid, tg = problem.obstacle_tables, robot.robot_id, problem.target_path
for _ in range(1):
    ops.lm_full_step(rid, D, ops.make_params(all_terms_parameters()), x0, None, tg, P, T, ob, True, out=xo)       # assemble + TMA solve
    ops.lm_full_step(rid, D, ops.make_params(ALT_LOSS_V2_1_DIFF), x0, None, tg, P, T, ob, True, out=xo)           # differencing step
    ops.lm_pose_step(rid, D, ops.make_params(ALT_LOSS_V2_1_POSE), x0, tg, True, out=xo)
    ops.collision_flags(rid, D, x0, ob)
    ops.path_metrics(rid, D, x0, tg, P, T, ob)
    ops.path_metrics(rid, D, x0, tg, P, T, ob, sign_only=True)
    ops.path_metrics(rid, D, x0[:T].contiguous(), tg, 1, T, ob)                                                    # cluster variant
    ops.lm_full_step(rid, D, ops.make_params(ALT_LOSS_V2_1_DIFF), x0[:T].contiguous(), None, tg, 1, T, ob, True)   # resident solve
    ops.lm_full_step(rid, D, ops.make_params(all_terms_parameters()), x0[:1024 * T].contiguous(), None, tg, 1024, T, ob, True)  # v2 solve
    for k in (175, 300):
        T2 = 295
        q = x0[:k * T2].reshape(k, T2, D).contiguous()
        sf = torch.zeros((k, T2), dtype=torch.uint8, device=dev)
        ops.dp_search(rid, D, q, sf, sf)
torch.cuda.synchronize()
print("done")
