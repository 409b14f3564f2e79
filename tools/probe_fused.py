"""CPPFLOW_LM_FUSED against the two-kernel step: single stream and chunk-pipelined (K iterations per chunk and stream),
P = 8192 (env P), T = 300."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters, ALT_LOSS_V2_1_DIFF
from cppflow_b200.pipeline import split_paths

dev = torch.device("cuda:0"); lib = _lib.load()
robot = get_robot("fetch"); P, T, D = int(os.environ.get("P", 8192)), 300, 8
problem = synthetic_problem(robot, T, device=dev)
x0 = synthetic_seeds_host(robot, P, T)[1].to(dev); xo = torch.empty_like(x0); xr = torch.empty_like(x0)
rid = robot.robot_id; cu, tc, no = ops._obs(problem.obstacle_tables)
K = int(os.environ.get("K", 20))


def run(prm, nch, flags, out):
    chunks = split_paths(P, nch)
    streams = [torch.cuda.Stream() for _ in chunks]
    wss = [torch.empty((lib.cppflow_lm_full_workspace_bytes(rid, n, T),), device=dev, dtype=torch.uint8) for _, n in chunks]

    def go():
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams: s.wait_stream(cur)
        for i in range(K):
            for (p0, n), s, ws in zip(chunks, streams, wss):
                sl = slice(p0 * T, (p0 + n) * T)
                _lib.check(lib.cppflow_lm_full_step(rid, prm, _lib.ptr(x0[sl]), None, _lib.ptr(problem.target_path), n, T, cu, tc, no, flags,
                                                    _lib.ptr(ws), ws.numel(), _lib.ptr(out[sl]), s.cuda_stream))
        for s in streams: cur.wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K
    go()
    return min(go(), go())


for name, pm in (("all terms", all_terms_parameters()), ("differencing step", ALT_LOSS_V2_1_DIFF)):
    prm = ops.make_params(pm)
    for nch in [int(v) for v in os.environ.get("CHUNKS", "1,2,4,6").split(",")]:
        ov = 2 if nch > 1 else 0
        t2 = run(prm, nch, 1 | ov, xr)
        tf = run(prm, nch, 1 | ov | 4, xo)
        print(f"{name}: chunks={nch}: two kernels {t2:.3f} ms/step, fused {tf:.3f} ms/step, equal={bool(torch.equal(xo, xr))}", flush=True)
