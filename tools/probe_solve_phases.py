"""Where a solve warp's cycles go, alone and under the assembly of other chunks (debug build:
CPPFLOW_SOLVE_TIMING=1 python -c "import __graft_entry__ as g; g.build()").  clock64 of lane 0 of every warp around the
phases of the forward loop (mbarrier wait | LDS + token + TMA issue | S update + sweep | stores) and of the
back-substitution loop (wait | rest)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from cppflow_b200.pipeline import split_paths

dev = torch.device("cuda:0"); lib = _lib.load()
dbg = lib._lib if hasattr(lib, "_lib") else lib
fn = ctypes.CDLL(_lib.LIB_PATH if hasattr(_lib, "LIB_PATH") else os.path.join(os.path.dirname(_lib.__file__), "libcppflow_b200.so")).cppflow_debug_solve_cycles
robot = get_robot("fetch"); P, T, D = 8192, 300, 8
problem = synthetic_problem(robot, T, device=dev)
x0 = synthetic_seeds_host(robot, P, T)[1].to(dev); xo = torch.empty_like(x0)
rid = robot.robot_id; cu, tc, no = ops._obs(problem.obstacle_tables)
prm = ops.make_params(all_terms_parameters())
NAMES = ["fwd wait", "fwd LDS+token+issue", "fwd update+sweep", "fwd stores", "back wait", "back rest", "fence+middle+back total"]


def report(label, n_warp_steps_fwd, n_warp_steps_back, n_warps):
    out = (ctypes.c_ulonglong * 8)()
    torch.cuda.synchronize(); fn(out, 0)
    v = list(out)
    fwd = [v[i] / n_warp_steps_fwd for i in range(4)]
    back = [v[i] / n_warp_steps_back for i in (4, 5)]
    print(f"{label}: cycles per forward step: " + ", ".join(f"{n} {c:.0f}" for n, c in zip(NAMES[:4], fwd)) + f" = {sum(fwd):.0f};  per back step: "
          + ", ".join(f"{n} {c:.0f}" for n, c in zip(NAMES[4:6], back)) + f" = {sum(back):.0f};  fence+middle+back per warp {v[6] / n_warps:.0f}", flush=True)


def run(nch, K, with_assembly):
    chunks = split_paths(P, nch)
    streams = [torch.cuda.Stream() for _ in chunks]
    wss = [torch.empty((lib.cppflow_lm_full_workspace_bytes(rid, n, T),), device=dev, dtype=torch.uint8) for _, n in chunks]
    for (p0, n), ws in zip(chunks, wss):  # valid blocks in every workspace
        _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0[p0 * T:(p0 + n) * T]), None, _lib.ptr(problem.target_path), n, T, cu, tc, no, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    fn(None, 1)
    cur = torch.cuda.current_stream()
    for s in streams: s.wait_stream(cur)
    for i in range(K):
        for (p0, n), s, ws in zip(chunks, streams, wss):
            sl = slice(p0 * T, (p0 + n) * T)
            if with_assembly:
                _lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0[sl]), None, _lib.ptr(problem.target_path), n, T, cu, tc, no, _lib.ptr(ws), ws.numel(), s.cuda_stream))
            _lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0[sl]), n, T, 3, _lib.ptr(ws), ws.numel(), _lib.ptr(xo[sl]), s.cuda_stream))
    for s in streams: cur.wait_stream(s)
    n_warps = (P // 16) * K
    report(f"chunks={nch} assembly={'yes' if with_assembly else 'no '}", n_warps * 150, n_warps * 150, n_warps)


run(1, 3, False)   # full-size solve alone (HBM-bound)
run(6, 3, False)   # six chunk solves at once, no assembly
run(6, 10, True)   # the pipelined step
run(4, 10, True)
