import torch, time, sys
sys.path.insert(0,'/root/repo')
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from cppflow_b200.pipeline import HostPipeline
dev=torch.device('cuda:0')
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T,pin=True)
oh=torch.empty_like(xh).pin_memory()
for nch,nrs in ((16,4),(8,4),(16,8),(32,4),(12,3),(16,2)):
  for ov in (True,False):
    pipe=HostPipeline(problem,P,all_terms_parameters(),n_chunks=nch,n_run_streams=nrs,overlap=ov)
    for _ in range(3): pipe.refine(xh,oh)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): pipe.refine(xh,oh)
    e1.record(); torch.cuda.synchronize()
    print(f"chunks={nch} run_streams={nrs} overlap={ov}: {e0.elapsed_time(e1)/20:.3f} ms/step", flush=True)
    del pipe
