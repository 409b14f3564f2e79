import torch, os, sys
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
    wsA=torch.empty_like(ws)
    _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(wsA), ws.numel(), st))
    torch.cuda.synchronize(); A=wsA
    _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(ws), ws.numel(), st))
    _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, flags, _lib.ptr(ws), ws.numel(), _lib.ptr(out), st))
    torch.cuda.synchronize(); return out, A.view(torch.float32).reshape(P//16,T,11,16,4), ws.clone().view(torch.float32).reshape(P//16,T,11,16,4)
o1,A,W1=run(1)
o2,A2,W2=run(1,'d')
print('A equal', bool(torch.equal(A,A2)))
neq=(W1!=W2).reshape(P//16,T,-1).any(dim=2)  # [G,T]
print('bad blocks', int(neq.sum()), 'groups with bad blocks', int(neq.any(dim=1).sum()))
gs=neq.any(dim=1).nonzero().flatten()[:6].tolist()
for g in gs:
    ts=neq[g].nonzero().flatten().tolist()
    print('group',g,'cta',g//4,'warp',g%4,'bad t:',ts[:20], '... n=',len(ts))
    for t in ts[:4]:
        bad=W2[g,t]; ref=W1[g,t]
        side0 = t<150
        prev = t-1 if side0 else t+1
        nxt = t+1 if side0 else t-1
        def eq(a,b): return bool(torch.equal(a,b))
        info={'==A(t)':eq(bad,A[g,t]),'==ref(prev)':eq(bad,W1[g,prev]) if 0<=prev<T else None,'==bad(prev)':eq(bad,W2[g,prev]) if 0<=prev<T else None,'==ref(next)':eq(bad,W1[g,nxt]) if 0<=nxt<T else None}
        # partial: which float4 rows (k) differ, which lanes differ
        d=(bad!=ref)
        info['k rows differing']=d.any(dim=2).any(dim=1).nonzero().flatten().tolist()
        info['paths differing']=d.any(dim=2).any(dim=0).nonzero().flatten().tolist()
        # does the bad block match ref of the mirrored side block?
        mt=T-1-t
        info['==ref(mirror)']=eq(bad,W1[g,mt]); info['==bad(mirror)']=eq(bad,W2[g,mt])
        # search other groups same t
        m=(W1[:,t].reshape(P//16,-1)==bad.reshape(1,-1)).all(dim=1).nonzero().flatten().tolist()
        info['matches ref of groups at t']=m[:4]
        print('   t',t,info)
print('x equal', bool(torch.equal(o1,o2)))
d=(o1-o2).abs().reshape(P,T,D)
badp=(d>0).any(dim=2).any(dim=1).nonzero().flatten()
print('bad paths', len(badp), 'distinct groups', len(set((badp//16).tolist())))
for p in badp[:10].tolist():
    ts=(d[p]>0).any(dim=1).nonzero().flatten().tolist()
    print('path',p,'group',p//16,'lane',p%16,'n bad t',len(ts),'range',ts[0],ts[-1],'contiguous',ts==list(range(ts[0],ts[-1]+1)),'err at first',d[p,ts[0]].max().item(),'max',d[p].max().item(), 'err profile', [round(d[p,t].max().item(),4) for t in ts[:12]])
# is x_bad[t] == q[t'] + dx_ref[t] for some other t' (q slot stale)?
q=x0.reshape(P,T,D)
for p in badp[:5].tolist():
    ts=(d[p]>0).any(dim=1).nonzero().flatten().tolist()
    t=ts[0]
    dxref=o1.reshape(P,T,D)[p,t]-q[p,t]
    cand=o2.reshape(P,T,D)[p,t]-dxref   # the q that was used if dx right
    m=((q[p]-cand).abs().max(dim=1).values<1e-5).nonzero().flatten().tolist()
    print('path',p,'t',t,'q row used (if dx ok):',m)
