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
ws=torch.empty(nb,dtype=torch.uint8,device=dev); wsA=torch.empty_like(ws)
pm=all_terms_parameters(); prm=ops.make_params(pm)
_lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(wsA), ws.numel(), st))
torch.cuda.synchronize()
A=wsA.view(torch.float32).reshape(P//16,T,11,16,4)
def run(flags, var=None):
    if var: os.environ['CPPFLOW_DEBUG_SOLVE_VARIANT']=var
    else: os.environ.pop('CPPFLOW_DEBUG_SOLVE_VARIANT',None)
    out=torch.empty_like(x0)
    _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(ws), ws.numel(), st))
    _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, flags, _lib.ptr(ws), ws.numel(), _lib.ptr(out), st))
    torch.cuda.synchronize(); return out, ws.clone().view(torch.float32).reshape(P//16,T,11,16,4)
o1,W1=run(1)
o2,W2=run(1,'d')
print('x equal',bool(torch.equal(o1,o2)),'W equal',bool(torch.equal(W1,W2)))
q=x0.reshape(P,T,D); X1=o1.reshape(P,T,D); X2=o2.reshape(P,T,D)
d=(X1-X2).abs()
beta=float(pm.alpha_differencing)**2
bvec=torch.full((D,),beta,device=dev); bvec[0]=(float(pm.alpha_differencing)*float(pm.alpha_differencing_prismatic_scaling))**2
def blockvals(Wt,g,t,l):  # 44 floats of path lane l
    return Wt[g,t,:,l,:].reshape(-1)
def tri(i,j): return i*(i+1)//2+j
def backsub(blk,dx_inner):
    z=bvec*dx_inner
    out=torch.empty(D,device=dev)
    for i in range(D):
        s=blk[36+i].clone()
        for j in range(D):
            a=blk[tri(i,j)] if j<=i else blk[tri(j,i)]
            s=s-a*z[j]
        out[i]=s
    return out
badp=(d>0).any(dim=2)
groups=sorted(set((badp.any(dim=1).nonzero().flatten()//16).tolist()))[:12]
for g in groups:
    p=g*16
    ts=badp[p].nonzero().flatten().tolist()
    if not ts: continue
    # side 0 event: highest bad t below 150; side 1 event: lowest bad t above 150
    for side,tt in ((0,max([t for t in ts if t<150],default=None)),(1,min([t for t in ts if t>150],default=None))):
        if tt is None: continue
        inner=tt+1 if side==0 else tt-1
        dx_inner=X1[p,inner]-q[p,inner]
        obs=X2[p,tt]-q[p,tt]
        res={}
        for name,(Wt,t2) in {'ref':(W1,tt),'A':(A,tt),'slotprev(+3)':(W1,tt+3 if side==0 else tt-3),'slotprev(+6)':(W1,tt+6 if side==0 else tt-6),'inner':(W1,inner),'outer':(W1,tt-1 if side==0 else tt+1),'A slotprev':(A,tt+3 if side==0 else tt-3)}.items():
            if 0<=t2<T:
                pred=backsub(blockvals(Wt,g,t2,0),dx_inner)
                res[name]=float((pred-obs).abs().max())
        print('group',g,'side',side,'event t',tt,'err',float(d[p,tt].max()),' |pred-obs| by candidate block:',{k:('%.2e'%v) for k,v in res.items()})
import itertools, numpy as np
print('--- mixture test: each float4 row from ref(t) or from the slot previous content (t+-3)')
bv=bvec.cpu().numpy().astype(np.float64)
def backsub_np(blk,dxin):
    z=bv*dxin; out=np.zeros(D)
    for i in range(D):
        s=blk[36+i]
        for j in range(D):
            a=blk[tri(i,j)] if j<=i else blk[tri(j,i)]
            s-=a*z[j]
        out[i]=s
    return out
cnt=0
for g in groups:
    p=g*16
    ts=badp[p].nonzero().flatten().tolist()
    for side,tt in ((0,max([t for t in ts if t<150],default=None)),(1,min([t for t in ts if t>150],default=None))):
        if tt is None: continue
        inner=tt+1 if side==0 else tt-1
        for off in (3,-3,1,-1):
         prev=tt+off if side==0 else tt-off
         if not (0<=prev<T): continue
         for l in (0,5):
          if True:
            pp=p+l
            dxin=(X1[pp,inner]-q[pp,inner]).double().cpu().numpy()
            obs=(X2[pp,tt]-q[pp,tt]).double().cpu().numpy()
            bref=blockvals(W1,g,tt,l).double().cpu().numpy(); bprev=blockvals(W1,g,prev,l).double().cpu().numpy()
            best=(1e9,None)
            for mask in range(2048):
                blk=bref.copy()
                for k in range(11):
                    if mask>>k&1: blk[4*k:4*k+4]=bprev[4*k:4*k+4]
                e=np.abs(backsub_np(blk,dxin)-obs).max()
                if e<best[0]: best=(e,mask)
            print('group',g,'lane',l,'side',side,'t',tt,'other block offset (in sweep order, + = read earlier)',off,'best mixture err %.2e'%best[0],'mask',bin(best[1]), 'no-mix err %.2e'%np.abs(backsub_np(bref,dxin)-obs).max())
        cnt+=1
        if cnt>=4: break
    if cnt>=4: break
