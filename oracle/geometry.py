"""Capsule-capsule and capsule-cuboid signed distances with analytic joint gradients.
TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates what the reference obtains from jrl (PARITY UNPINNED - jrl not installable offline):
  * `robot.self_collision_distances(x)` [n,S]        (collision_detection.py:65, optimization_utils.py:652)
  * `robot.env_collision_distances(x, cuboid, Tcuboid)` [n,C]  (collision_detection.py:40, optimization_utils.py:690)
  * `robot.self_collision_distances_jacobian(x)` [n,S,D], `robot.env_collision_distances_jacobian` [n,C,D]
    (optimization_utils.py:670,710)
Negative distance = overlap (collision_detection.py:43,67; optimization_utils.py:644-647).
jrl solves a small box-constrained QP per pair for the closest points; the closed forms below
(clamped segment-segment closest points; exact piecewise-linear root for segment-vs-box) are
the exact minimisers of those QPs.  Cuboid = [-sx/2,-sy/2,-sz/2, sx/2,sy/2,sz/2] in the frame
`Tcuboid` (data_type_utils.py:109-127; only Tcuboid[:3,:4] is meaningful, [3,3] is left 0).
"""
from typing import Tuple

import torch

from .robots import RobotModel
from .kinematics import capsule_world_endpoints, joints_moving_frame

_EPS = 1e-12


def _dot(a, b):
    return (a * b).sum(-1)


def segment_segment_closest(P1, Q1, P2, Q2) -> Tuple[torch.Tensor, torch.Tensor]:
    """Parameters (s,t) in [0,1]^2 of the closest points of segments P1Q1 and P2Q2 (any leading shape)."""
    d1, d2, r = Q1 - P1, Q2 - P2, P1 - P2
    a, e, f = _dot(d1, d1), _dot(d2, d2), _dot(d2, r)
    c, b = _dot(d1, r), _dot(d1, d2)
    denom = a * e - b * b
    a_s = torch.clamp(a, min=_EPS)
    e_s = torch.clamp(e, min=_EPS)
    s = torch.where(denom > _EPS, torch.clamp((b * f - c * e) / torch.clamp(denom, min=_EPS), 0.0, 1.0),
                    torch.zeros_like(a))
    t = (b * s + f) / e_s
    s_lo = torch.clamp(-c / a_s, 0.0, 1.0)
    s_hi = torch.clamp((b - c) / a_s, 0.0, 1.0)
    s = torch.where(t < 0.0, s_lo, torch.where(t > 1.0, s_hi, s))
    t = torch.clamp(t, 0.0, 1.0)
    return s, t


def segment_box_closest(A, B, lo, hi) -> torch.Tensor:
    """Parameter t in [0,1] of the point of segment AB closest to the axis-aligned box [lo,hi].

    f(t) = |P(t) - clamp(P(t))|^2 is convex and C1; h(t) = 0.5 f'(t) = d . (P - clamp(P)) is monotone
    piecewise linear with breakpoints where P(t) crosses a face plane.  Bracket the root between the
    breakpoints and interpolate linearly: exact.
    """
    d = B - A

    def h(t):
        P = A + t[..., None] * d
        return _dot(d, P - torch.minimum(torch.maximum(P, lo), hi))

    zero = torch.zeros_like(A[..., 0])
    one = torch.ones_like(zero)
    h0, h1 = h(zero), h(one)
    t_lo, h_lo, t_hi, h_hi = zero, h0, one, h1
    for k in range(3):
        dk = d[..., k]
        safe = torch.where(dk.abs() > _EPS, dk, torch.ones_like(dk))
        for face in (lo[..., k], hi[..., k]):
            tb = torch.where(dk.abs() > _EPS, (face - A[..., k]) / safe, -one)
            inside = (tb > 0.0) & (tb < 1.0)
            hb = h(torch.clamp(tb, 0.0, 1.0))
            up_lo = inside & (hb <= 0.0) & (tb > t_lo)
            up_hi = inside & (hb > 0.0) & (tb < t_hi)
            t_lo, h_lo = torch.where(up_lo, tb, t_lo), torch.where(up_lo, hb, h_lo)
            t_hi, h_hi = torch.where(up_hi, tb, t_hi), torch.where(up_hi, hb, h_hi)
    dh = h_hi - h_lo
    t_mid = t_lo + (t_hi - t_lo) * torch.where(dh > _EPS, -h_lo / torch.clamp(dh, min=_EPS), zero)
    t = torch.where(h0 >= 0.0, zero, torch.where(h1 <= 0.0, one, t_mid))
    return torch.clamp(t, 0.0, 1.0)


def self_collision_distances(model: RobotModel, x: torch.Tensor, with_jacobian: bool = False):
    """[n,S] signed capsule-capsule distances (and optionally d/dq [n,S,D])."""
    P1, P2, radii, axes, origins = capsule_world_endpoints(model, x)
    ia = torch.tensor([p[0] for p in model.pairs])
    ib = torch.tensor([p[1] for p in model.pairs])
    A1, B1, A2, B2 = P1[:, ia], P2[:, ia], P1[:, ib], P2[:, ib]
    s, t = segment_segment_closest(A1, B1, A2, B2)
    C1 = A1 + s[..., None] * (B1 - A1)
    C2 = A2 + t[..., None] * (B2 - A2)
    diff = C1 - C2
    dist = torch.sqrt(_dot(diff, diff))
    d = dist - radii[ia] - radii[ib]
    if not with_jacobian:
        return d
    nrm = diff / torch.clamp(dist, min=_EPS)[..., None]
    nrm = torch.where((dist > _EPS)[..., None], nrm, torch.zeros_like(nrm))
    J = torch.zeros((x.shape[0], len(model.pairs), model.ndof), dtype=x.dtype)
    for k, (ca, cb) in enumerate(model.pairs):
        for sign, cap, C in ((1.0, ca, C1[:, k]), (-1.0, cb, C2[:, k])):
            for dj in joints_moving_frame(model, model.capsules[cap].frame):
                ci = model.actuated[dj]
                if model.chain[ci].jtype == "revolute":
                    v = torch.cross(axes[dj], C - origins[dj], dim=1)
                else:
                    v = axes[dj]
                J[:, k, dj] += sign * _dot(nrm[:, k], v)
    return d, J


def env_collision_distances(model: RobotModel, x: torch.Tensor, cuboid: torch.Tensor, Tcuboid: torch.Tensor,
                            with_jacobian: bool = False):
    """[n,C] signed capsule-cuboid distances for one obstacle (and optionally d/dq [n,C,D])."""
    P1, P2, radii, axes, origins = capsule_world_endpoints(model, x)
    cuboid = cuboid.to(x.dtype)
    Rb = Tcuboid[:3, :3].to(x.dtype)
    tb = Tcuboid[:3, 3].to(x.dtype)
    lo, hi = cuboid[0:3], cuboid[3:6]
    # world -> box frame: p_b = Rb^T (p - tb)
    A = (P1 - tb) @ Rb
    B = (P2 - tb) @ Rb
    t = segment_box_closest(A, B, lo, hi)
    Cb = A + t[..., None] * (B - A)
    Q = torch.minimum(torch.maximum(Cb, lo), hi)
    diff = Cb - Q
    dist = torch.sqrt(_dot(diff, diff))
    d = dist - radii
    if not with_jacobian:
        return d
    nrm_b = torch.where((dist > _EPS)[..., None], diff / torch.clamp(dist, min=_EPS)[..., None], torch.zeros_like(diff))
    nrm = nrm_b @ Rb.T  # back to world
    Cw = P1 + t[..., None] * (P2 - P1)
    J = torch.zeros((x.shape[0], len(model.capsules), model.ndof), dtype=x.dtype)
    for k, cap in enumerate(model.capsules):
        for dj in joints_moving_frame(model, cap.frame):
            ci = model.actuated[dj]
            if model.chain[ci].jtype == "revolute":
                v = torch.cross(axes[dj], Cw[:, k] - origins[dj], dim=1)
            else:
                v = axes[dj]
            J[:, k, dj] = _dot(nrm[:, k], v)
    return d, J
