import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops
from cppflow_b200.robot import get_robot
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from tests.helpers import OBSTACLES, cuboid_tensors, synthetic_problem
dev = torch.device("cuda:0")
r = "fetch"; rob = get_robot(r)
cub, Tc = cuboid_tensors(OBSTACLES[r]); ob = ops.Obstacles(cub, Tc)
out = {}
for P, T, S in ((48, 37, 5), (48, 37, 3), (1, 37, 5), (16, 37, 5), (17, 37, 5), (48, 40, 5), (48, 300, 16)):
    m, target, x0 = synthetic_problem(r, P, T, seed=P + T)
    x, tg = x0.to(dev), target.to(dev)
    prm = ops.make_params(all_terms_parameters())
    got = ops.lm_full_step(rob.robot_id, rob.ndof, prm, x, None, tg, P, T, ob, False, segments=S)
    out[(P, T, S)] = got.cpu()
tag = "nostage" if os.environ.get("CPPFLOW_SEG_NO_STAGE") else "stage"
torch.save(out, f"/tmp/seg_{tag}.pt")
if tag == "stage" and os.path.exists("/tmp/seg_nostage.pt"):
    ref = torch.load("/tmp/seg_nostage.pt")
    for key, g in out.items():
        P, T, S = key
        d = (g - ref[key]).abs().max(dim=1).values.view(P, T)
        bad = (d > 1e-4).nonzero()
        print(key, "max", float(d.max()), "n_bad", len(bad), "first", bad[:8].tolist(), "paths", sorted(set(bad[:, 0].tolist()))[:20], "ts", sorted(set(bad[:, 1].tolist())))
