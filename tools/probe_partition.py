import torch, time, sys
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from cppflow_b200.pipeline import ResidentPipeline
dev=torch.device('cuda:0')
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T)
x0=xh.to(dev); xo=torch.empty_like(x0)
prm=ops.make_params(all_terms_parameters())
ref=ops.lm_full_step(robot.robot_id,D,prm,x0,None,problem.target_path,P,T,problem.obstacle_tables,True)
def run(pipe,K=100):
    def go():
        pipe.begin()
        for _ in range(K): pipe.enqueue_step(x0,xo)
        pipe.end()
    go(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); go(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/K, bool(torch.equal(xo,ref))
print('shared SMs, 4 chunks: %.3f ms eq=%s'%run(ResidentPipeline(problem,P,all_terms_parameters(),n_chunks=4)))
for sms in (int(a) for a in (sys.argv[1:] or ['24','32','40'])):
    for nch,ov in ((4,True),(4,False),(8,True)):
        try:
            pipe=ResidentPipeline(problem,P,all_terms_parameters(),n_chunks=nch,solve_sms=sms,overlap=ov)
            t,eq=run(pipe)
            print(f'solve partition {pipe.partition.sms_first} SMs / assembly {pipe.partition.sms_second} SMs, {nch} chunks ring4={ov}: {t:.3f} ms eq={eq}',flush=True)
        except Exception as e:
            print('solve_sms',sms,'chunks',nch,'failed:',e,flush=True); break
