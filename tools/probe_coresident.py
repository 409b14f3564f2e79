"""What slows the block solve next to another kernel?  The chunk pipeline of probe_timeline.py with the assembly
replaced by a stand-in of about the same duration: FILLER=assemble (the real thing), pose (lm_pose_step_kernel x 3: FP32,
96 registers, NO shared memory), flags (collision_flags_kernel x 2: shared-memory column like the assembly, 72 registers).
Prints the solve's span per chunk."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters, ALT_LOSS_V2_1_POSE
from cppflow_b200.pipeline import split_paths

dev = torch.device("cuda:0"); lib = _lib.load()
robot = get_robot("fetch"); P, T, D = 8192, 300, 8
problem = synthetic_problem(robot, T, device=dev)
x0 = synthetic_seeds_host(robot, P, T)[1].to(dev); xo = torch.empty_like(x0); xs = torch.empty_like(x0)
rid = robot.robot_id; ob = problem.obstacle_tables; cu, tc, no = ops._obs(ob)
prm = ops.make_params(all_terms_parameters()); prm_pose = ops.make_params(ALT_LOSS_V2_1_POSE)
scratch = torch.zeros(16, device=dev)
K, nch = 20, int(os.environ.get("CHUNKS", 6))
chunks = split_paths(P, nch)
streams = [torch.cuda.Stream() for _ in chunks]
wss = [torch.empty((lib.cppflow_lm_full_workspace_bytes(rid, n, T),), device=dev, dtype=torch.uint8) for _, n in chunks]
for (p0, n), ws in zip(chunks, wss):
    _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0[p0 * T:(p0 + n) * T]), None, _lib.ptr(problem.target_path), n, T, cu, tc, no, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
torch.cuda.synchronize()
for filler in os.environ.get("FILLER", "assemble,pose,flags,none").split(","):
    def fill(sl, n, ws, s):
        if filler == "assemble":
            _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0[sl]), None, _lib.ptr(problem.target_path), n, T, cu, tc, no, _lib.ptr(ws), ws.numel(), s.cuda_stream))
        elif filler == "pose":
            for _ in range(3):
                ops.lm_pose_step(rid, D, prm_pose, x0[sl], problem.target_path, True, out=xs[sl])
        elif filler == "flags":
            for _ in range(2):
                ops.collision_flags(rid, D, x0[sl], ob)
        elif filler.startswith("nb"):  # nb<KB>t<touch>: stand-in neighbour with KB of shared memory per CTA, `touch` accesses per iteration
            kb, touch = filler[2:].split("t")
            _lib.check(lib.cppflow_neighbour_probe(n // 256 * T, 120 if int(touch) == 0 else 60, int(kb) * 1024, int(touch), _lib.ptr(scratch), s.cuda_stream))

    def run(record):
        evs = [[[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)] for _ in chunks]
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams: s.wait_stream(cur)
        for i in range(K):
            for c, ((p0, n), s, ws) in enumerate(zip(chunks, streams, wss)):
                sl = slice(p0 * T, (p0 + n) * T)
                with torch.cuda.stream(s):
                    evs[c][i][0].record(s)
                    fill(sl, n, ws, s)
                    evs[c][i][1].record(s)
                    _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0[sl]), n, T, 3, _lib.ptr(ws), ws.numel(), _lib.ptr(xo[sl]), s.cuda_stream))
                    evs[c][i][2].record(s)
        for s in streams: cur.wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K, evs
    run(False)
    t, evs = run(True)
    m = lambda v: sum(v) / len(v)
    fl = [evs[c][i][0].elapsed_time(evs[c][i][1]) for c in range(len(chunks)) for i in range(5, K - 2)]
    so = [evs[c][i][1].elapsed_time(evs[c][i][2]) for c in range(len(chunks)) for i in range(5, K - 2)]
    print(f"filler={filler:9s} chunks={len(chunks)}: {t:.3f} ms/step | per chunk: filler span {m(fl):.3f} ms, solve span {m(so):.3f} ms", flush=True)
