import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops
from cppflow_b200.collision_detection import qpaths_batched_collisions
from cppflow_b200.data_type_utils import problem_from_filename, ALL_PROBLEM_FILENAMES
from cppflow_b200.optimization import run_lm_optimization
from cppflow_b200.optimization_utils import path_metrics
from cppflow_b200.search import dp_search
from cppflow_b200.planners import LmIkCandidateGenerator, LatentIkCandidateGenerator
dev = torch.device("cuda:0")
gens = {"latent stride 16": lambda: LatentIkCandidateGenerator(seed=1),
        "latent stride 8": lambda: LatentIkCandidateGenerator(seed=1, stride=8)}
tot = {g: [0, 0.0, 0.0] for g in gens}
for name in ALL_PROBLEM_FILENAMES:
    problem = problem_from_filename(None, name, device=dev); rob = problem.robot; T = problem.n_timesteps
    print(name, T)
    for gname, mk in gens.items():
        g = mk()
        qs = g(problem, 175).contiguous(); torch.cuda.synchronize()
        g = mk()
        t0 = time.perf_counter(); qs = g(problem, 175).contiguous(); torch.cuda.synchronize(); t_gen = (time.perf_counter() - t0) * 1e3
        err, _ = ops.pose_errors(rob.robot_id, rob.ndof, qs.reshape(-1, rob.ndof), problem.target_path)
        ok = ((err[:, 3:].norm(dim=1) < 1e-4) & (err[:, :3].norm(dim=1) < 1.7e-3)).reshape(175, T)
        sv, ev = qpaths_batched_collisions(problem, qs)
        from cppflow_b200.planners import _with_unreached_waypoints
        ev = _with_unreached_waypoints(ev, g)
        best = dp_search(rob, qs, sv, ev, verbosity=0).to(dev).contiguous()
        m = path_metrics(problem, best, 1).cpu()[0].tolist()
        res = run_lm_optimization(problem, best, max_n_steps=20, tmax_sec=30.0, return_if_valid_after_n_steps=0, convergence_threshold=1e6, verbosity=0)
        m2 = path_metrics(problem, res.x_opt.contiguous(), 1).cpu()[0].tolist()
        tot[gname][0] += int(res.is_valid); tot[gname][1] += float(ok.float().mean()); tot[gname][2] += t_gen
        print(f"   [{gname}] gen {t_gen:.2f} ms conv {ok.float().mean()*100:.0f}% fullpaths {int(ok.all(dim=1).sum())} coll {float((sv|ev).float().mean())*100:.0f}% | dp: pos {m[0]:.2g}cm mjac {m[2]:.1f}deg {m[3]:.1f}cm | LM {res.n_steps_taken+1} steps {res.schedule} valid={res.is_valid} pos {m2[0]:.2g} rot {m2[1]:.2g} mjac {m2[2]:.1f}/{m2[3]:.1f} minself {m2[5]:.3f} minenv {m2[6]:.3f} TL {m2[4]:.2f}")
    sys.stdout.flush()
for g, (v, c, t) in tot.items():
    print(f"{g}: valid {v}/13, mean convergence {c/13*100:.1f}%, mean gen {t/13:.2f} ms")
