"""Iteration time of the resident pipeline at strong-scaling path counts: twisted solve with the default chunking against
the segmented solve over chunk counts.  One JSON line per path count."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from cppflow_b200.pipeline import ResidentPipeline

dev = torch.device("cuda:0")
robot = get_robot("fetch"); T = 300
problem = synthetic_problem(robot, T, device=dev)
_, xh = synthetic_seeds_host(robot, 4096, T)
x0 = xh.to(dev)
K = int(os.environ.get("K", "20"))

def step_ms(rp, xs, xo):
    def run(n):
        rp.begin()
        a, b = xs, xo
        for _ in range(n):
            rp.enqueue_step(a, b)
        rp.end()
    run(5); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(K); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    return round(best, 4)

for P in (256, 512, 1024, 2048, 4096):
    xs = x0[: P * T].contiguous(); xo = torch.empty_like(xs)
    out = {"paths": P, "twisted_default": step_ms(ResidentPipeline(problem, P, all_terms_parameters()), xs, xo)}
    for S in (8, 16):
        for chunks in (1, 2, 4, 8):
            out[f"S{S}_chunks{chunks}"] = step_ms(ResidentPipeline(problem, P, all_terms_parameters(), n_chunks=chunks, segments=S), xs, xo)
    print(json.dumps(out), flush=True)
