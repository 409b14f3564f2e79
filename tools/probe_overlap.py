import torch, time, sys
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters, ALT_LOSS_V2_1_DIFF
from cppflow_b200.pipeline import ResidentPipeline
dev=torch.device('cuda:0'); lib=_lib.load()
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T)
x0=xh.to(dev); xo=torch.empty_like(x0); xo2=torch.empty_like(x0)
ob=problem.obstacle_tables; rid=robot.robot_id
cu,tc,no=ops._obs(ob); st=_lib.stream_ptr(dev)
ws=ops._workspace(dev, lib.cppflow_lm_full_workspace_bytes(rid,P,T), "lm_full")
prm=ops.make_params(all_terms_parameters())
def timeit(fn,n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
fa=lambda: _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(ws), ws.numel(), st))
for flags,name in ((1,"ring 3 (alone)"),(3,"ring 4 (overlap)")):
    out = xo if flags==1 else xo2
    fa(); 
    fs=lambda: _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, flags, _lib.ptr(ws), ws.numel(), _lib.ptr(out), st))
    # solve overwrites ws: time assemble+solve pairs and assemble alone
    ta=timeit(fa); 
    def both(): fa(); fs()
    tb=timeit(both)
    print(f"{name}: assemble {ta:.3f} ms, assemble+solve {tb:.3f} ms -> solve {tb-ta:.3f} ms")
fa(); _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, 1, _lib.ptr(ws), ws.numel(), _lib.ptr(xo), st))
fa(); _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, 3, _lib.ptr(ws), ws.numel(), _lib.ptr(xo2), st))
torch.cuda.synchronize()
print('ring 4 (overlap) == ring 3 (alone) bitwise:', bool(torch.equal(xo,xo2)))
for nch in (2,3,4,5,6,8):
    for ov in (True,):
        pipe=ResidentPipeline(problem,P,all_terms_parameters(),n_chunks=nch,overlap=ov)
        K=100
        def run():
            pipe.begin()
            for _ in range(K): pipe.enqueue_step(x0,xo2)
            pipe.end()
        run(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        t0=time.perf_counter(); e0.record(); run(); e1.record(); t1=time.perf_counter(); torch.cuda.synchronize()
        print(f"chunks={len(pipe.chunks)} overlap={ov}: {e0.elapsed_time(e1)/K:.3f} ms/step (host enqueue {(t1-t0)/K*1e3:.3f}) equal={bool(torch.equal(xo,xo2))}")
        del pipe
# dependent iterations
pipe=ResidentPipeline(problem,P,all_terms_parameters(),n_chunks=4)
r=pipe.iterate(x0,5); 
y=x0
for i in range(5): y=ops.lm_full_step(rid,D,prm,y,None,problem.target_path,P,T,ob,True)
torch.cuda.synchronize(); print('iterate(5) == 5 sequential steps:', bool(torch.equal(r,y)))
