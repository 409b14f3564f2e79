import torch, time, sys
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters, ALT_LOSS_V2_1_DIFF, ALT_LOSS_V2_1_POSE
dev=torch.device('cuda:0')
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T)
x0=xh.to(dev); xo=torch.empty_like(x0)
ob=problem.obstacle_tables
def timeit(fn,n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1=time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n, (t1-t0)/n*1e3
for name,pm in (('all',all_terms_parameters()),('diff',ALT_LOSS_V2_1_DIFF)):
    prm=ops.make_params(pm)
    print(name,'full step ms (gpu, host-launch): %.3f %.3f'%timeit(lambda: ops.lm_full_step(robot.robot_id,D,prm,x0,None,problem.target_path,P,T,ob,True,out=xo)))
import ctypes
lib=_lib.load(); cu,tc,no=ops._obs(ob); st=_lib.stream_ptr(dev)
for name,pm in (('all',all_terms_parameters()),('diff',ALT_LOSS_V2_1_DIFF)):
    prm=ops.make_params(pm)
    ws=ops._workspace(dev, lib.cppflow_lm_full_workspace_bytes(robot.robot_id,P,T), "lm_full")
    fa=lambda: _lib.check(lib.cppflow_lm_full_assemble(robot.robot_id, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(ws), ws.numel(), st))
    fs=lambda: _lib.check(lib.cppflow_lm_full_solve(robot.robot_id, prm, _lib.ptr(x0), P, T, 1, _lib.ptr(ws), ws.numel(), _lib.ptr(xo), st))
    print(name,'assemble ms: %.3f %.3f'%timeit(fa), ' solve ms: %.3f %.3f'%timeit(fs))
prm=ops.make_params(ALT_LOSS_V2_1_POSE)
print('pose step ms: %.3f %.3f'%timeit(lambda: ops.lm_pose_step(robot.robot_id,D,prm,x0,problem.target_path,True,out=xo)))
print('flags ms: %.3f %.3f'%timeit(lambda: ops.collision_flags(robot.robot_id,D,x0,ob)))
print('metrics ms: %.3f %.3f'%timeit(lambda: ops.path_metrics(robot.robot_id,D,x0,problem.target_path,P,T,ob)))
print('fk ms: %.3f %.3f'%timeit(lambda: ops.forward_kinematics(robot.robot_id,D,x0)))
# dp_search fetch circle size
k,T2=175,295
q=x0[:k*T2].reshape(k,T2,D).contiguous()
sf=torch.zeros((k,T2),dtype=torch.uint8,device=dev)
print('dp_search k=175 T=295 ms: %.3f %.3f'%timeit(lambda: ops.dp_search(robot.robot_id,D,q,sf,sf),n=20))
k=300
q=x0[:k*T2].reshape(k,T2,D).contiguous(); sf=torch.zeros((k,T2),dtype=torch.uint8,device=dev)
print('dp_search k=300 T=295 ms: %.3f %.3f'%timeit(lambda: ops.dp_search(robot.robot_id,D,q,sf,sf),n=20))
prm=ops.make_params(all_terms_parameters())
fn=lambda: ops.lm_full_step(robot.robot_id,D,prm,x0,None,problem.target_path,P,T,ob,True,out=xo)
for n in (20,50,100,200,400):
    print('all-terms n=%d: gpu %.3f ms/step host %.3f'%((n,)+timeit(fn,n=n)))
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active --format=csv",shell=True,capture_output=True,text=True).stdout)
# per-iteration events
evs=[torch.cuda.Event(enable_timing=True) for _ in range(201)]
torch.cuda.synchronize(); evs[0].record()
for i in range(200):
    fn(); evs[i+1].record()
torch.cuda.synchronize()
ts=[evs[i].elapsed_time(evs[i+1]) for i in range(200)]
print('per-iter ms: first10', [round(t,2) for t in ts[:10]], 'around 50', [round(t,2) for t in ts[45:55]], 'last10', [round(t,2) for t in ts[-10:]])
