import torch, time, sys, os
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from cppflow_b200.pipeline import ResidentPipeline
dev=torch.device('cuda:0'); lib=_lib.load()
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T)
x0=xh.to(dev); xo2=torch.empty_like(x0)
ob=problem.obstacle_tables; rid=robot.robot_id
prm=ops.make_params(all_terms_parameters())
os.environ.pop('CPPFLOW_DEBUG_SOLVE_VARIANT',None)
ref=ops.lm_full_step(rid,D,prm,x0,None,problem.target_path,P,T,ob,True)
def timeit(fn,n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
names={'-':'default ring3 alone / ring4 overlap, 4w, q smem','a':'ring3 4w qreg','b':'ring4 4w qreg','c':'ring3 6w qreg','d':'ring5 4w qreg','e':'ring2 8w qreg'}
for var in sys.argv[1:] or list(names):
    if var=='-': os.environ.pop('CPPFLOW_DEBUG_SOLVE_VARIANT',None)
    else: os.environ['CPPFLOW_DEBUG_SOLVE_VARIANT']=var
    t1=timeit(lambda: ops.lm_full_step(rid,D,prm,x0,None,problem.target_path,P,T,ob,True,out=xo2))
    eq1=bool(torch.equal(xo2,ref))
    res=[]
    for nch in (3,4,6,8):
        pipe=ResidentPipeline(problem,P,all_terms_parameters(),n_chunks=nch)
        K=100
        def run():
            pipe.begin()
            for _ in range(K): pipe.enqueue_step(x0,xo2)
            pipe.end()
        run(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        res.append('%d:%.3f%s'%(nch,e0.elapsed_time(e1)/K,'' if torch.equal(xo2,ref) else '(NEQ)'))
        del pipe
    print(f"variant {var} [{names[var]}]: single-stream step {t1:.3f} ms eq={eq1}; pipelined chunks:ms {' '.join(res)}", flush=True)
