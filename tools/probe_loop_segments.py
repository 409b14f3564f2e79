"""Plan latency with the single-path differencing step on the segmented (CPPFLOW_LOOP_SEGMENTS, default 16) or the twisted
(0) solve: one problem alone and the 13 benchmark problems sequentially / batched.  Run once per setting."""
import os, statistics, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200.collision_detection import qpaths_batched_collisions
from cppflow_b200.data_type_utils import ALL_PROBLEM_FILENAMES, problem_from_filename
from cppflow_b200.data_types import PlannerSettings
from cppflow_b200.optimization import run_lm_optimization
from cppflow_b200.planners import CppFlowPlanner, LatentIkCandidateGenerator, plan_many
from cppflow_b200.search import dp_search
dev = torch.device("cuda:0")
problem = problem_from_filename(None, "fetch__circle", device=dev)
qs = LatentIkCandidateGenerator(seed=3)(problem, 175).contiguous()
sv, ev = qpaths_batched_collisions(problem, qs)
best = dp_search(problem.robot, qs, sv, ev, verbosity=0).to(dev).contiguous()
def wall(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), r
t_lm, res = wall(lambda: run_lm_optimization(problem, best, max_n_steps=20, tmax_sec=30.0, return_if_valid_after_n_steps=0,
                                             convergence_threshold=1e6, verbosity=0))
problems = [problem_from_filename(None, n, device=dev) for n in ALL_PROBLEM_FILENAMES]
qsets = {}
class Cached:
    def __init__(self): self.last_converged = None
    def __call__(self, p, k):
        if p.full_name not in qsets:
            g = LatentIkCandidateGenerator(seed=3); qsets[p.full_name] = (g(p, k).contiguous(), g.last_converged)
        self.last_converged = qsets[p.full_name][1]
        return qsets[p.full_name][0]
def factory(p): return CppFlowPlanner(PlannerSettings(k=175, tmax_sec=30.0, anytime_mode_enabled=False, verbosity=0), p.robot, Cached())
t_many, r = wall(lambda: plan_many(factory, problems), n=10)
t_seq, _ = wall(lambda: [factory(p).generate_plan(p) for p in problems], n=10)
print({"loop_segments": os.environ.get("CPPFLOW_LOOP_SEGMENTS", "16"), "lm_loop_ms": round(t_lm, 4), "schedule": res.schedule,
       "plan_many_ms": round(t_many, 3), "sequential_ms": round(t_seq, 3), "valid": sum(int(x.plan.is_valid) for x in r)})
