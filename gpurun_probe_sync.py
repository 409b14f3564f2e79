import torch, sys
sys.path.insert(0,'/root/repo')
from cppflow_b200 import ops
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
dev=torch.device('cuda:0')
robot=get_robot('fetch'); P,T,D=int(sys.argv[1]),40,8
problem=synthetic_problem(robot,T,device=dev)
_,xh=synthetic_seeds_host(robot,P,T)
x0=xh.to(dev)
prm=ops.make_params(all_terms_parameters())
o=ops.lm_full_step(robot.robot_id,D,prm,x0,None,problem.target_path,P,T,problem.obstacle_tables,True)
torch.cuda.synchronize(); print('ok',P,float(o.abs().max()))
