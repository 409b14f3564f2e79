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
for nch,nrs in ((12,3),(6,3),(9,3),(15,3),(18,3),(24,3),(12,4),(12,2),(12,6),(10,5),(16,4),(12,3)):
  for ov in (True,):
    pipe=HostPipeline(problem,P,all_terms_parameters(),n_chunks=nch,n_run_streams=nrs,overlap=ov)
    for _ in range(3): pipe.refine(xh,oh)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): pipe.refine(xh,oh)
    e1.record(); torch.cuda.synchronize()
    print(f"chunks={nch} run_streams={nrs} overlap={ov}: {e0.elapsed_time(e1)/20:.3f} ms/step", flush=True)
    del pipe

# raw copy bandwidths
import torch
x=xh; d=torch.empty_like(xh,device=dev); o=oh
def t(fn,n=10):
    fn(); torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print('H2D alone ms', t(lambda: d.copy_(x,non_blocking=True)), 'D2H alone ms', t(lambda: o.copy_(d,non_blocking=True)))
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(x,non_blocking=True)
    with torch.cuda.stream(s2): o.copy_(d,non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
print('H2D + D2H concurrently ms', t(both))
