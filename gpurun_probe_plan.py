import sys, time, torch
sys.path.insert(0,'/root/repo')
from cppflow_b200.data_type_utils import ALL_PROBLEM_FILENAMES, problem_from_filename
from cppflow_b200.data_types import PlannerSettings
from cppflow_b200.planners import CppFlowPlanner, LmIkCandidateGenerator
for k in (64, 175):
  for name in ALL_PROBLEM_FILENAMES:
    problem = problem_from_filename(None, name, device='cuda:0')
    rob = problem.robot
    planner = CppFlowPlanner(PlannerSettings(k=k, tmax_sec=30.0, anytime_mode_enabled=False, verbosity=0), rob, LmIkCandidateGenerator(seed=1))
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.perf_counter()
        res = planner.generate_plan(problem)
        torch.cuda.synchronize(); dt=time.perf_counter()-t0
    p=res.plan; td=res.timing
    print(f"k={k} {name:22s} T={problem.n_timesteps} valid={p.is_valid} pos_cm={p.max_pos_error_cm:.4f} rot_deg={p.max_rot_error_deg:.4f} mjac_deg={p.mjac_deg:.2f} mjac_cm={p.mjac_cm:.2f} self={p.min_self_distance_m:.3f} env={p.min_env_distance_m:.3f} steps={res.debug_info.get('n_optimization_steps')} total_ms={dt*1e3:.1f} (gen {td.ikflow*1e3:.1f} coll {td.coll_checking*1e3:.1f} dp {td.dp_search*1e3:.1f} opt {td.optimizer*1e3:.1f})")
