"""Per-kernel spans inside the chunk pipeline: K iterations of (assemble, solve) per chunk on its own stream, a CUDA
event before and after every kernel.  Prints the mean span of a chunk's assembly and solve and the idle gap between a
chunk's solve and its next assembly, in steady state."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from cppflow_b200.pipeline import split_paths

dev = torch.device("cuda:0"); lib = _lib.load()
robot = get_robot("fetch"); P, T, D = int(os.environ.get("P", 8192)), 300, 8
problem = synthetic_problem(robot, T, device=dev)
x0 = synthetic_seeds_host(robot, P, T)[1].to(dev); xo = torch.empty_like(x0)
rid = robot.robot_id; cu, tc, no = ops._obs(problem.obstacle_tables)
prm = ops.make_params(all_terms_parameters())
K = int(os.environ.get("K", 20))
for nch in [int(v) for v in os.environ.get("CHUNKS", "4,6,2").split(",")]:
    chunks = split_paths(P, nch)
    streams = [torch.cuda.Stream() for _ in chunks]
    wss = [torch.empty((lib.cppflow_lm_full_workspace_bytes(rid, n, T),), device=dev, dtype=torch.uint8) for _, n in chunks]
    def run(record):
        evs = [[[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)] for _ in chunks] if record else None
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams: s.wait_stream(cur)
        for i in range(K):
            for c, ((p0, n), s, ws) in enumerate(zip(chunks, streams, wss)):
                sl = slice(p0 * T, (p0 + n) * T)
                with torch.cuda.stream(s):
                    if record: evs[c][i][0].record(s)
                    _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0[sl]), None, _lib.ptr(problem.target_path), n, T, cu, tc, no, _lib.ptr(ws), ws.numel(), s.cuda_stream))
                    if record: evs[c][i][1].record(s)
                    _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0[sl]), n, T, 3, _lib.ptr(ws), ws.numel(), _lib.ptr(xo[sl]), s.cuda_stream))
                    if record: evs[c][i][2].record(s)
        for s in streams: cur.wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K, evs
    run(False)
    t_plain, _ = run(False)
    t_rec, evs = run(True)
    asm = [evs[c][i][0].elapsed_time(evs[c][i][1]) for c in range(len(chunks)) for i in range(5, K - 2)]
    sol = [evs[c][i][1].elapsed_time(evs[c][i][2]) for c in range(len(chunks)) for i in range(5, K - 2)]
    cyc = [evs[c][i][0].elapsed_time(evs[c][i + 1][0]) for c in range(len(chunks)) for i in range(5, K - 2)]
    m = lambda v: sum(v) / len(v)
    print(f"chunks={len(chunks)} K={K}: {t_plain:.3f} ms/step ({t_rec:.3f} with events) | per chunk: assembly span {m(asm):.3f} ms, "
          f"solve span {m(sol):.3f} ms, cycle {m(cyc):.3f} ms (span = queued + running)", flush=True)
    # where the chunks are relative to each other at iteration 10
    base = evs[0][10][0]
    print("   iteration 10 offsets (ms) [assembly start, assembly end, solve end] per chunk:",
          [[round(base.elapsed_time(evs[c][10][j]), 3) for j in range(3)] for c in range(len(chunks))], flush=True)
