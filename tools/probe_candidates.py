import torch, time, sys, dataclasses
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops
from cppflow_b200.collision_detection import qpaths_batched_collisions
from cppflow_b200.data_type_utils import problem_from_filename, ALL_PROBLEM_FILENAMES
from cppflow_b200.optimization import run_lm_optimization
from cppflow_b200.optimization_utils import path_metrics
from cppflow_b200.search import dp_search
from cppflow_b200.lm_hyper_parameters import ALT_LOSS_V2_1_POSE
dev=torch.device('cuda:0')
def gen(problem,k,lams,seed=1,spread=0.6):
    robot,T=problem.robot,problem.n_timesteps
    g=torch.Generator().manual_seed(seed)
    lim=torch.tensor(robot.actuated_joints_limits,dtype=torch.float32)
    mid,half=lim.mean(dim=1),(lim[:,1]-lim[:,0])/2
    base=mid+spread*half*(2*torch.rand((k,1,robot.ndof),generator=g)-1)
    x=base.expand(k,T,robot.ndof).reshape(k*T,robot.ndof).contiguous().to(dev)
    for lam in lams:
        prm=ops.make_params(dataclasses.replace(ALT_LOSS_V2_1_POSE,lm_lambda=lam))
        x=ops.lm_pose_step(robot.robot_id,robot.ndof,prm,x,problem.target_path,True)
    return x.reshape(k,T,robot.ndof)
variants={'6x1e-6 (current)':[1e-6]*6,'12x1e-6':[1e-6]*12,'damped 12':[1e-1,1e-1,3e-2,3e-2,1e-2,1e-2,1e-3,1e-3,1e-4,1e-5,1e-6,1e-6],'damped 20':[1e-1]*4+[3e-2]*4+[1e-2]*4+[1e-3]*3+[1e-4,1e-5,1e-6,1e-6,1e-6]}
for name in ALL_PROBLEM_FILENAMES:
    problem=problem_from_filename(None,name,device=dev); rob=problem.robot; T=problem.n_timesteps
    out=[]
    for vn,lams in variants.items():
        qs=gen(problem,175,lams).contiguous()
        err,_=ops.pose_errors(rob.robot_id,rob.ndof,qs.reshape(-1,rob.ndof),problem.target_path)
        ok=((err[:,3:].norm(dim=1)<1e-4)&(err[:,:3].norm(dim=1)<1.7e-3)).reshape(175,T)
        sv,ev=qpaths_batched_collisions(problem,qs)
        best=dp_search(rob,qs,sv,ev,verbosity=0).to(dev).contiguous()
        m=path_metrics(problem,best,1).cpu()[0].tolist()
        res=run_lm_optimization(problem,best,max_n_steps=20,tmax_sec=30.0,return_if_valid_after_n_steps=0,convergence_threshold=1e6,verbosity=0)
        m2=path_metrics(problem,res.x_opt.contiguous(),1).cpu()[0].tolist()
        out.append(f"[{vn}] conv {ok.float().mean()*100:.0f}% fullpaths {int(ok.all(dim=1).sum())} coll {float((sv|ev).float().mean())*100:.0f}% | dp: pos {m[0]:.2g}cm mjac {m[2]:.1f}deg {m[3]:.1f}cm | LM {res.n_steps_taken+1} steps valid={res.is_valid} pos {m2[0]:.2g} rot {m2[1]:.2g} mjac {m2[2]:.1f}/{m2[3]:.1f} minself {m2[5]:.3f} minenv {m2[6]:.3f}")
    print(name,T); [print('   ',o) for o in out]; sys.stdout.flush()
