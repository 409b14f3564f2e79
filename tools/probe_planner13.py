import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200.data_type_utils import problem_from_filename, ALL_PROBLEM_FILENAMES
from cppflow_b200.data_types import PlannerSettings
from cppflow_b200.planners import CppFlowPlanner, LatentIkCandidateGenerator, LmIkCandidateGenerator
dev = torch.device("cuda:0")
for label, kw, gen in (("reference ROS settings (rerun on large mjac and on failure)",
                        dict(do_rerun_if_large_dp_search_mjac=True, do_rerun_if_optimization_fails=True),
                        lambda s: LatentIkCandidateGenerator(seed=s)),
                       ("rerun on large mjac", dict(do_rerun_if_large_dp_search_mjac=True), lambda s: LatentIkCandidateGenerator(seed=s)),
                       ("no rerun", dict(), lambda s: LatentIkCandidateGenerator(seed=s))):
    for seed in (1, 2, 3, 4):
        nv, bad, t_tot = 0, [], 0.0
        for name in ALL_PROBLEM_FILENAMES:
            problem = problem_from_filename(None, name, device=dev)
            pl = CppFlowPlanner(PlannerSettings(k=175, tmax_sec=30.0, anytime_mode_enabled=False, verbosity=0, **kw), problem.robot, gen(seed))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = pl.generate_plan(problem)
            torch.cuda.synchronize(); t_tot += time.perf_counter() - t0
            nv += int(res.plan.is_valid)
            if not res.plan.is_valid:
                bad.append((name, round(res.plan.max_pos_error_cm, 4), round(res.plan.mjac_deg, 1), round(res.plan.min_env_distance_m, 3)))
        print(f"{label} seed {seed}: valid {nv}/13 in {t_tot*1e3:.1f} ms", bad)
