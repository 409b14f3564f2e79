import torch, sys, ctypes as C
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
dev=torch.device('cuda:0'); lib=_lib.load()
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T)
x0=xh.to(dev); xo=torch.empty_like(x0)
prm=ops.make_params(all_terms_parameters()); ob=problem.obstacle_tables; rid=robot.robot_id
cu,tc,no=ops._obs(ob)
tp=_lib.ptr(problem.target_path)
def run(nchunks, nsteps=60):
    streams=[torch.cuda.Stream(dev) for _ in range(nchunks)]
    pc=P//nchunks
    wss=[torch.empty((lib.cppflow_lm_full_workspace_bytes(rid,pc,T),),device=dev,dtype=torch.uint8) for _ in range(nchunks)]
    def go():
        for s in range(nsteps):
            for c in range(nchunks):
                sl=slice(c*pc*T,(c+1)*pc*T)
                st=C.c_void_p(streams[c].cuda_stream)
                _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0[sl]), None, tp, pc, T, cu, tc, no, _lib.ptr(wss[c]), wss[c].numel(), st))
                _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0[sl]), pc, T, 1, _lib.ptr(wss[c]), wss[c].numel(), _lib.ptr(xo[sl]), st))
    go(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    cur=torch.cuda.current_stream(dev)
    e0.record(cur)
    for s_ in streams: s_.wait_event(e0)
    go()
    for s_ in streams: cur.wait_stream(s_)
    e1.record(cur); torch.cuda.synchronize()
    print(f"chunks/streams={nchunks}: {e0.elapsed_time(e1)/nsteps:.3f} ms per full-P step")
for n in (1,2,4,8): run(n)
