"""Compare the segmented solve's buffers (factors, corners, separator solutions, x) on the GPU with the fp64 model."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cppflow_b200 import ops, _lib
from cppflow_b200.robot import get_robot
from cppflow_b200.synthetic import synthetic_problem, synthetic_seeds_host
from cppflow_b200.lm_hyper_parameters import all_terms_parameters
from segsolve_model import solve_segmented, dense
dev = torch.device("cuda:0")
robot = get_robot("fetch"); T, D, NT, NW = 300, 8, 36, 44
S = int(os.environ.get("S", "4")); P = 1
problem = synthetic_problem(robot, T, device=dev)
_, xh = synthetic_seeds_host(robot, 16, T)
x0 = xh.to(dev)[:P * T].contiguous()
pm = all_terms_parameters()
prm = ops.make_params(pm)
lib = _lib.load(); cu, tc, no = ops._obs(problem.obstacle_tables); st = _lib.stream_ptr(dev); rid = robot.robot_id
flags = ops.lm_segments(S)
nbytes = lib.cppflow_lm_full_workspace_bytes_ex(rid, P, T, flags)
ws = torch.zeros((nbytes,), device=dev, dtype=torch.uint8)
_lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(ws), ws.numel(), st))
torch.cuda.synchronize()
wsf = ws.view(torch.float32).cpu().numpy().copy()
blk = wsf[: T * 11 * 16 * 4].reshape(T, 11, 16, 4)[:, :, 0, :].reshape(T, 44)  # path 0 of group 0
tri = lambda i, j: i * (i + 1) // 2 + j
A = np.zeros((T, D, D)); b = blk[:, NT:NT + D].astype(np.float64)
for i in range(D):
    for j in range(i + 1):
        A[:, i, j] = A[:, j, i] = blk[:, tri(i, j)]
a = pm.alpha_differencing
beta = np.array([(a * (pm.alpha_differencing_prismatic_scaling if d == 0 else 1.0)) ** 2 for d in range(D)])
print("beta", beta)
xo = torch.empty_like(x0)
_lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, flags, _lib.ptr(ws), ws.numel(), _lib.ptr(xo), st))
torch.cuda.synchronize()
ws_t = torch.zeros((lib.cppflow_lm_full_workspace_bytes(rid, P, T),), device=dev, dtype=torch.uint8)
xt = torch.empty_like(x0)
_lib.check(lib.cppflow_lm_full_assemble(rid, prm, _lib.ptr(x0), None, _lib.ptr(problem.target_path), P, T, cu, tc, no, _lib.ptr(ws_t), ws_t.numel(), st))
_lib.check(lib.cppflow_lm_full_solve(rid, prm, _lib.ptr(x0), P, T, 0, _lib.ptr(ws_t), ws_t.numel(), _lib.ptr(xt), st))
torch.cuda.synchronize()
dump = {}
xm = solve_segmented(A, b, beta, S, dump)
xd = dense(A, b, beta)
print("model vs dense", np.abs(xm - xd).max())
dx_gpu = (xo - x0).cpu().numpy().astype(np.float64)
err = np.abs(dx_gpu - xd).max(axis=1)
dx_tw = (xt - x0).cpu().numpy().astype(np.float64)
err_tw = np.abs(dx_tw - xd).max(axis=1)
print("TWISTED gpu vs dense: max", err_tw.max(), "at", int(err_tw.argmax()), "mean", err_tw.mean(), "| segmented mean", np.abs(dx_gpu - xd).max(axis=1).mean(), "| step max", np.abs(xd).max())
M = np.zeros((T * D, T * D))
for t in range(T):
    M[t*D:(t+1)*D, t*D:(t+1)*D] = A[t]
    if t + 1 < T:
        M[t*D:(t+1)*D, (t+1)*D:(t+2)*D] = -np.diag(beta); M[(t+1)*D:(t+2)*D, t*D:(t+1)*D] = -np.diag(beta)
for name, dxv in (("twisted", dx_tw), ("segmented", dx_gpu), ("dense", xd)):
    r = M @ dxv.reshape(-1) - b.reshape(-1)
    print(name, "residual max", np.abs(r).max(), "rel", np.linalg.norm(r) / np.linalg.norm(b), "energy err", float((dxv - xd).reshape(-1) @ M @ (dxv - xd).reshape(-1)))
print("cond", np.linalg.cond(M), "lambda", pm.lm_lambda)
print("gpu vs dense: max", err.max(), "at", int(err.argmax()), "seps", dump["seps"], "segs", dump["segs"])
wsf2 = ws.view(torch.float32).cpu().numpy()
base = lib.cppflow_lm_full_workspace_bytes(rid, P, T)
off = (base + 255) // 256 * 256 // 4
blkf = 16 * T * 44
fac = wsf2[off: off + blkf].reshape(T, 11, 16, 4)[:, :, 0, :].reshape(T, 44)
off += blkf
CV = 27
cor = wsf2[off: off + S * 2 * CV * 16 * 4].reshape(S, 2, CV, 16, 4)[:, :, :, 0, :].reshape(S, 2, CV * 4)
off += S * 2 * CV * 16 * 4
sepx = wsf2[off: off + (S - 1) * 2 * 16 * 4].reshape(S - 1, 2, 16, 4)[:, :, 0, :].reshape(S - 1, 8)
def unpack(v):
    M = np.zeros((D, D))
    for i in range(D):
        for j in range(i + 1):
            M[i, j] = M[j, i] = v[tri(i, j)]
    return M
for j, s in enumerate(dump["seps"]):
    print("sep", s, "sepx err", np.abs(sepx[j] - xd[s]).max(), "model", np.abs(xm[s] - xd[s]).max())
for si, (fl, bk) in enumerate(dump["corners"]):
    nS, u, Q = fl
    print("seg", si, "corner up nS err", np.abs(unpack(cor[si, 0]) - nS).max(), "u err", np.abs(cor[si, 0, NT:NT + D] - u).max(),
          "Q err", np.abs(cor[si, 0, NW:NW + 64].reshape(8, 8) - Q).max(), "| Q max", np.abs(Q).max())
for t in (0, 10, 36, 37, 38, 40, 74, 75, 76, 100, 149, 150, 151):
    print("t", t, "err", err[t])
