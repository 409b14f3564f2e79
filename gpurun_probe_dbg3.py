import torch, os, sys, ctypes as C
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
dev=torch.device('cuda:0'); lib=_lib.load()
robot=get_robot('fetch'); P,T,D=8192,300,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T)
x0=xh.to(dev)
ob=problem.obstacle_tables; rid=robot.robot_id
cu,tc,no=ops._obs(ob); st=_lib.stream_ptr(dev)
nb=lib.cppflow_lm_full_workspace_bytes(rid,P,T)
ws=torch.empty(nb,dtype=torch.uint8,device=dev)
prm=ops.make_params(all_terms_parameters())
def run(flags, var=None):
    if var: os.environ['CPPFLOW_DEBUG_SOLVE_VARIANT']=var
    else: os.environ.pop('CPPFLOW_DEBUG_SOLVE_VARIANT',None)
    out=torch.empty_like(x0)
    _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(ws), ws.numel(), st))
    _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, flags, _lib.ptr(ws), ws.numel(), _lib.ptr(out), st))
    torch.cuda.synchronize(); return out
o1=run(1)
o2=run(1,'n')
print('x equal', bool(torch.equal(o1,o2)))
cnt=C.c_uint(0); buf=(C.c_float*(4096*8))()
lib.cppflow_debug_read.argtypes=[C.POINTER(C.c_uint), C.POINTER(C.c_float)]
lib.cppflow_debug_read(C.byref(cnt), buf)
print('mismatch records', cnt.value)
import numpy as np
a=np.ctypeslib.as_array(buf).reshape(4096,8)[:min(cnt.value,4096)]
# group by (g,t)
from collections import defaultdict
dd=defaultdict(list)
for r in a: dd[(int(r[0]),int(r[1]))].append(r)
print('distinct (group,t):', len(dd))
for (g,t),rs in list(sorted(dd.items()))[:25]:
    print('g',g,'t',t,'lanes',sorted(int(r[2]) for r in rs)[:32],'nbad',[int(r[3]) for r in rs][:6],'firstbad',[int(r[4]) for r in rs][:6],'smem/glob',[(float(r[5]),float(r[6])) for r in rs][:2],'same_prev',[int(r[7]) for r in rs][:6])
