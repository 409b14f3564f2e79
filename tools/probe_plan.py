import torch, time, sys, statistics
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops
from cppflow_b200.collision_detection import qpaths_batched_collisions
from cppflow_b200.data_type_utils import problem_from_filename
from cppflow_b200.optimization import run_lm_optimization
from cppflow_b200.optimization_utils import path_metrics
from cppflow_b200.planners import LmIkCandidateGenerator
from cppflow_b200.search import dp_search
from cppflow_b200.lm_hyper_parameters import ALT_LOSS_V2_1_DIFF, ALT_LOSS_V2_1_POSE
dev=torch.device('cuda:0')
problem=problem_from_filename(None,'fetch__circle',device=dev)
rob=problem.robot; T=problem.n_timesteps; D=rob.ndof
qs=LmIkCandidateGenerator(seed=1)(problem,175).contiguous()
def wall(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        t0=time.perf_counter(); r=fn(); torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    return statistics.median(ts), r
t_flags,(sv,ev)=wall(lambda: qpaths_batched_collisions(problem,qs))
t_dp,best=wall(lambda: dp_search(rob,qs,sv,ev,verbosity=0))
best=best.to(dev).contiguous()
t_lm,res=wall(lambda: run_lm_optimization(problem,best,max_n_steps=20,tmax_sec=30.0,return_if_valid_after_n_steps=0,convergence_threshold=1e6,verbosity=0))
print(f"flags {t_flags:.3f} ms, dp_search {t_dp:.3f} ms, LM loop {t_lm:.3f} ms ({res.n_steps_taken+1} steps, schedule {res.schedule}, valid {res.is_valid})")
pp=ops.make_params(ALT_LOSS_V2_1_POSE); pd=ops.make_params(ALT_LOSS_V2_1_DIFF)
t_pose,_=wall(lambda: ops.lm_pose_step(rob.robot_id,D,pp,best,problem.target_path,True))
t_diff,_=wall(lambda: ops.lm_full_step(rob.robot_id,D,pd,best,None,problem.target_path,1,T,problem.obstacle_tables,True))
t_met,_=wall(lambda: path_metrics(problem,best,1))
t_metcpu,_=wall(lambda: path_metrics(problem,best,1).cpu())
print(f"single path: pose step {t_pose:.3f} ms, diff step {t_diff:.3f} ms, metrics {t_met:.3f} ms, metrics+.cpu() {t_metcpu:.3f} ms (wall incl. launch+sync)")
def gpu(fn,n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print(f"gpu time back-to-back: pose {gpu(lambda: ops.lm_pose_step(rob.robot_id,D,pp,best,problem.target_path,True)):.4f} diff {gpu(lambda: ops.lm_full_step(rob.robot_id,D,pd,best,None,problem.target_path,1,T,problem.obstacle_tables,True)):.4f} metrics {gpu(lambda: path_metrics(problem,best,1)):.4f} flags {gpu(lambda: qpaths_batched_collisions(problem,qs)):.4f} dp {gpu(lambda: dp_search(rob,qs,sv,ev,verbosity=0)):.4f}")
