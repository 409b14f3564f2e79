import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops
from cppflow_b200.collision_detection import qpaths_batched_collisions
from cppflow_b200.data_type_utils import problem_from_filename
from cppflow_b200.optimization_utils import path_metrics
from cppflow_b200.search import dp_search
from cppflow_b200.planners import LatentIkCandidateGenerator, _with_unreached_waypoints
from cppflow_b200.lm_hyper_parameters import ALT_LOSS_V2_1_DIFF, ALT_LOSS_V2_1_POSE
dev = torch.device("cuda:0")
problem = problem_from_filename(None, "panda__flappy_bird", device=dev); rob = problem.robot; T = problem.n_timesteps
g = LatentIkCandidateGenerator(seed=1)
qs = g(problem, 175).contiguous()
sv, ev = qpaths_batched_collisions(problem, qs)
print("colliding waypoint fraction per t (env):", [round(float(ev[:, t].float().mean()), 2) for t in range(0, T, 10)])
free = ~(sv | ev) & g.last_converged
print("candidates free at every waypoint:", int(free.all(dim=1).sum()), " min over t of #free candidates:", int(free.sum(dim=0).min()))
best = dp_search(rob, qs, sv, _with_unreached_waypoints(ev, g), verbosity=0).contiguous()
x = best
prm_p, prm_d = ops.make_params(ALT_LOSS_V2_1_POSE), ops.make_params(ALT_LOSS_V2_1_DIFF)
def show(tag, x):
    m = path_metrics(problem, x, 1).cpu()[0].tolist()
    d = rob.env_collision_distances(x, problem.obstacles_cuboids[0], problem.obstacles_Tcuboids[0]).min(dim=1).values
    d2 = rob.env_collision_distances(x, problem.obstacles_cuboids[1], problem.obstacles_Tcuboids[1]).min(dim=1).values
    dm = torch.minimum(d, d2)
    print(f"{tag}: pos {m[0]:.4f} rot {m[1]:.3f} mjac {m[2]:.1f} TL {m[4]:.2f} minself {m[5]:.3f} minenv {m[6]:.4f}  n_coll_waypoints {int((dm<0).sum())} argmin t {int(dm.argmin())}")
show("dp", x)
for c in "pdpppdpppdppp":
    if c == "p":
        x = ops.lm_pose_step(rob.robot_id, rob.ndof, prm_p, x, problem.target_path, True)
    else:
        x = ops.lm_full_step(rob.robot_id, rob.ndof, prm_d, x, None, problem.target_path, 1, T, problem.obstacle_tables, True)
    show(c, x)
