"""HostPipeline sweep (chunks x run streams x depth, jobs submitted with refine_async) and the raw PCIe copy rates."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from cppflow_b200.pipeline import HostPipeline

dev = torch.device("cuda:0")
robot = get_robot("fetch"); P, T, D = 8192, 300, 8
problem = synthetic_problem(robot, T, device=dev)
_, xh = synthetic_seeds_host(robot, P, T, pin=True)
ohs = [torch.empty_like(xh).pin_memory() for _ in range(3)]
import ast
CASES = ast.literal_eval(os.environ.get("E2E_CASES", "[(16, 4, 2, True), (8, 4, 2, True), (6, 3, 2, True), (4, 4, 2, True), (8, 4, 3, True), (16, 4, 1, True)]"))
for nch, nrs, depth, graph in CASES:
    pipe = HostPipeline(problem, P, all_terms_parameters(), n_chunks=nch, n_run_streams=nrs, depth=depth, use_graph=graph)
    for i in range(2 * depth):
        pipe.refine_async(xh, ohs[i % depth])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = [pipe.refine_async(xh, ohs[i % depth]) for i in range(20)]
    for ev in done:
        torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    print(f"chunks={nch} run_streams={nrs} depth={depth} graph={graph}: {e0.elapsed_time(e1) / 20:.3f} ms/step", flush=True)
    del pipe

x = xh; d = torch.empty_like(xh, device=dev); o = ohs[0]
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("H2D alone ms", t(lambda: d.copy_(x, non_blocking=True)), "D2H alone ms", t(lambda: o.copy_(d, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): o.copy_(d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
print("H2D + D2H concurrently ms", t(both))
