import torch, time, sys, cProfile, pstats
sys.path.insert(0,'/root/repo')
from cppflow_b200.data_type_utils import ALL_PROBLEM_FILENAMES, problem_from_filename
from cppflow_b200.data_types import PlannerSettings
from cppflow_b200.planners import CppFlowPlanner, LmIkCandidateGenerator, plan_many
dev=torch.device('cuda:0')
problems=[problem_from_filename(None,n,device=dev) for n in ALL_PROBLEM_FILENAMES]
def factory(p): return CppFlowPlanner(PlannerSettings(k=175,tmax_sec=30.0,anytime_mode_enabled=False,verbosity=0),p.robot,LmIkCandidateGenerator(seed=1))
for _ in range(3): plan_many(factory,problems)
torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable()
for _ in range(5): plan_many(factory,problems)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
