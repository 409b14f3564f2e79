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
prm=ops.make_params(all_terms_parameters())
ref=ops.lm_full_step(rid,D,prm,x0,None,problem.target_path,P,T,ob,True)
for i in range(6):
    o=ops.lm_full_step(rid,D,prm,x0,None,problem.target_path,P,T,ob,True,overlap=(i%2==0))
    print('overlap' if i%2==0 else 'deep', bool(torch.equal(o,ref)), float((o-ref).abs().max()))
