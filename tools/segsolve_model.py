"""fp64 model of the segmented block-tridiagonal solve (csrc/lm_segsolve.cuh): the three passes with numpy, checked
against a dense solve when run as a script; tools/debug_seg.py compares the GPU's intermediate buffers with it."""
import numpy as np

def segments(T, S):
    """separators s_j = (j*T)//S for j=1..S-1 ; segments between them"""
    seps = [(j * T) // S for j in range(1, S)]
    bounds = [-1] + seps + [T]
    segs = [(bounds[i] + 1, bounds[i + 1] - 1) for i in range(S)]
    return seps, segs

def solve_segmented(A, b, beta, S, dump=None):
    T, D, _ = A.shape
    Bm = np.diag(beta)
    seps, segs = segments(T, S)
    # pass 1
    F = [None] * T; Bk = [None] * T
    corners = []
    for (a, e) in segs:
        nS = np.zeros((D, D)); u = np.zeros(D); Q = np.eye(D)
        for t in range(a, e + 1):
            Sm = A[t] + Bm @ nS @ Bm; y = b[t] + beta * u
            Si = np.linalg.inv(Sm); nS = -Si; u = Si @ y
            F[t] = (nS.copy(), u.copy())
            Q = Q @ (Si @ Bm)
        fl = (nS, u, Q)
        nS = np.zeros((D, D)); u = np.zeros(D)
        for t in range(e, a - 1, -1):
            Sm = A[t] + Bm @ nS @ Bm; y = b[t] + beta * u
            Si = np.linalg.inv(Sm); nS = -Si; u = Si @ y
            Bk[t] = (nS.copy(), u.copy())
        corners.append((fl, (nS, u)))
    # reduced
    n = S - 1
    Ah = []; bh = []; K = []
    for j in range(n):
        s = seps[j]
        (nSl, ul, _), _ = corners[j]          # left segment: fwd end corner (last,last)
        (_, _, Qr), (nSr, ur) = corners[j + 1]  # right segment: bwd end corner (first,first) ; cross
        Ah.append(A[s] + Bm @ (nSl + nSr) @ Bm)
        bh.append(b[s] + beta * (ul + ur))
        K.append(-Bm @ Qr)  # couples s_j with s_{j+1} (through right segment)
    x = np.zeros((T, D))
    Sh = [None] * n; uh = [None] * n
    for j in range(n):
        Sm = Ah[j].copy(); y = bh[j].copy()
        if j > 0:
            Sm -= K[j - 1].T @ Sh[j - 1] @ K[j - 1]
            y -= K[j - 1].T @ uh[j - 1]
        Sh[j] = np.linalg.inv(Sm); uh[j] = Sh[j] @ y
    for j in range(n - 1, -1, -1):
        xs = uh[j].copy()
        if j < n - 1:
            xs -= Sh[j] @ (K[j] @ x[seps[j + 1]])
        x[seps[j]] = xs
    # pass 3
    for i, (a, e) in enumerate(segs):
        xp = x[seps[i - 1]] if i > 0 else np.zeros(D)
        xn = x[seps[i]] if i < S - 1 else np.zeros(D)
        L = e - a + 1
        mid = a + L // 2
        # side 0
        nS0 = np.zeros((D, D)); u0 = xp.copy(); du = xp.copy(); ut = {}
        for t in range(a, mid):
            nS_t, u_t = F[t]
            du = -nS_t @ (beta * du)
            ut[t] = u_t + du; nS0 = nS_t; u0 = ut[t]
        nS1 = np.zeros((D, D)); u1 = xn.copy(); du = xn.copy()
        for t in range(e, mid, -1):
            nS_t, u_t = Bk[t]
            du = -nS_t @ (beta * du)
            ut[t] = u_t + du; nS1 = nS_t; u1 = ut[t]
        Sm = A[mid] + Bm @ (nS0 + nS1) @ Bm
        y = b[mid] + beta * (u0 + u1)
        x[mid] = np.linalg.solve(Sm, y)
        dx = x[mid]
        for t in range(mid - 1, a - 1, -1):
            dx = ut[t] - F[t][0] @ (beta * dx); x[t] = dx
        dx = x[mid]
        for t in range(mid + 1, e + 1):
            dx = ut[t] - Bk[t][0] @ (beta * dx); x[t] = dx
    if dump is not None:
        dump.update(F=F, Bk=Bk, corners=corners, seps=seps, segs=segs, Sh=Sh, uh=uh, K=K)
    return x

def dense(A, b, beta):
    T, D, _ = A.shape
    M = np.zeros((T * D, T * D))
    for t in range(T):
        M[t*D:(t+1)*D, t*D:(t+1)*D] = A[t]
        if t + 1 < T:
            M[t*D:(t+1)*D, (t+1)*D:(t+2)*D] = -np.diag(beta)
            M[(t+1)*D:(t+2)*D, t*D:(t+1)*D] = -np.diag(beta)
    return np.linalg.solve(M, b.reshape(-1)).reshape(T, D)

if __name__ == '__main__':
  rng = np.random.default_rng(0)
  for T, S, D in [(300, 16, 8), (37, 5, 7), (9, 4, 8), (300, 2, 8), (20, 1, 7), (8, 4, 8)]:
      J = rng.normal(size=(T, 12, D)); beta = np.abs(rng.normal(size=D)) + 0.5
      A = np.einsum('tki,tkj->tij', J, J) + 0.01 * np.eye(D) + 2 * np.diag(beta**2)
      b = rng.normal(size=(T, D))
      x = solve_segmented(A, b, beta, S)
      xr = dense(A, b, beta)
      print(T, S, D, np.abs(x - xr).max(), segments(T, S)[1][:3])
