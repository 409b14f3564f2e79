import torch, sys, os, ctypes as C
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters, ALT_LOSS_V2_1_DIFF
dev=torch.device('cuda:0'); lib=_lib.load()
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T)
x0=xh.to(dev); xo=torch.empty_like(x0)
ob=problem.obstacle_tables; rid=robot.robot_id
cu,tc,no=ops._obs(ob); st=_lib.stream_ptr(dev)
ws=ops._workspace(dev, lib.cppflow_lm_full_workspace_bytes(rid,P,T), "lm_full")
def timeit(fn,n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for var in sys.argv[1:]:
    key,vals=var.split('=')
    for v in vals.split(','):
        os.environ[key]=v
        for name,pm in (('all',all_terms_parameters()),('diff',ALT_LOSS_V2_1_DIFF)):
            prm=ops.make_params(pm)
            fa=lambda: _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(ws), ws.numel(), st))
            fs=lambda: _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, 1, _lib.ptr(ws), ws.numel(), _lib.ptr(xo), st))
            print(f"{key}={v} {name}: assemble {timeit(fa):.3f} ms  solve {timeit(fs):.3f} ms")
