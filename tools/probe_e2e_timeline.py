import torch, sys
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from cppflow_b200.pipeline import HostPipeline
dev=torch.device('cuda:0')
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T,pin=True)
oh=torch.empty_like(xh).pin_memory()
pipe=HostPipeline(problem,P,all_terms_parameters(),n_chunks=16,n_run_streams=4,use_graph=False)
# re-implement _enqueue with timing events
def enqueue_timed():
    T_,D_,rid=pipe.T,pipe.robot.ndof,pipe.robot.robot_id
    cur=torch.cuda.current_stream(dev)
    ev0=torch.cuda.Event(enable_timing=True); ev0.record(cur)
    for s in [pipe.s_in,pipe.s_out]+pipe.s_run: s.wait_stream(cur)
    lib=ops._lib.load(); cu,tc,no=ops._obs(problem.obstacle_tables)
    evs=[]
    for c,(p0,n) in enumerate(pipe.chunks):
        sl=slice(p0*T_,(p0+n)*T_)
        e_in=torch.cuda.Event(enable_timing=True); e_run0=torch.cuda.Event(enable_timing=True); e_run=torch.cuda.Event(enable_timing=True); e_out0=torch.cuda.Event(enable_timing=True); e_out=torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(pipe.s_in):
            pipe.x_dev[sl].copy_(xh[sl],non_blocking=True); e_in.record(pipe.s_in)
        sr=pipe.s_run[c%len(pipe.s_run)]
        with torch.cuda.stream(sr):
            sr.wait_event(e_in); e_run0.record(sr)
            ws=pipe.ws[c%len(pipe.s_run)]
            ops.check(lib.cppflow_lm_full_step(rid,pipe.prm,ops.ptr(pipe.x_dev[sl]),None,ops.ptr(problem.target_path),n,T_,cu,tc,no,pipe.flags,ops.ptr(ws),ws.numel(),ops.ptr(pipe.out_dev[sl]),ops.stream_ptr(dev)))
            e_run.record(sr)
        with torch.cuda.stream(pipe.s_out):
            pipe.s_out.wait_event(e_run); e_out0.record(pipe.s_out)
            oh[sl].copy_(pipe.out_dev[sl],non_blocking=True); e_out.record(pipe.s_out)
        evs.append((e_in,e_run0,e_run,e_out0,e_out))
    for s in [pipe.s_out,pipe.s_in]+pipe.s_run: cur.wait_stream(s)
    return ev0,evs
for _ in range(3): enqueue_timed()
torch.cuda.synchronize()
ev0,evs=enqueue_timed(); torch.cuda.synchronize()
print('chunk: H2D done | run start .. run done | D2H start .. D2H done   (ms from start)')
for c,(a,b0,b,c0,d) in enumerate(evs):
    print(f'{c:2d}: {ev0.elapsed_time(a):.3f} | {ev0.elapsed_time(b0):.3f} .. {ev0.elapsed_time(b):.3f} | {ev0.elapsed_time(c0):.3f} .. {ev0.elapsed_time(d):.3f}')
