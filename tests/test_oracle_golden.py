"""CPU: the oracle restatement against vectors produced by the REAL reference code
(tests/golden/make_golden.py) and against the known-answer vectors of the reference's own tests."""
import math

import numpy as np
import pytest
import torch

from oracle import robots as R, kinematics as K, geometry as G, lm as L, search as S
from oracle.math_utils import geodesic_distance_between_quaternions

ROBOTS = ["fetch", "fetch_arm", "panda"]


def T(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype)


def cub(golden, r):
    c, t = golden[f"{r}/lm/cuboids"], golden[f"{r}/lm/Tcuboids"]
    return [T(x) for x in t], [T(x) for x in c]


@pytest.mark.parametrize("r", ROBOTS)
def test_dp_search_bit_exact(golden, r):
    m = R.get_model(r)
    q = T(golden[f"{r}/dp/q"])
    best, memo, costs, _ = S.dp_search(m, q, torch.tensor(golden[f"{r}/dp/self_v"]), torch.tensor(golden[f"{r}/dp/env_v"]))
    assert np.array_equal(memo.numpy(), golden[f"{r}/dp/memo"])
    assert np.array_equal(costs.numpy(), golden[f"{r}/dp/costs"])
    assert np.array_equal(best.numpy(), golden[f"{r}/dp/best_path"])
    assert np.array_equal(S.get_mjacs(q, m).numpy(), golden[f"{r}/dp/mjacs"])
    assert np.array_equal(S.joint_limit_almost_violations_3d(m, q).numpy(), golden[f"{r}/dp/jlim"])


@pytest.mark.parametrize("r", ["fetch_arm", "panda"])
def test_dp_search_matches_slow_variant(golden, r):
    # search.py:55-97 is a second implementation (without the prismatic x5): equal for all-revolute robots
    m = R.get_model(r)
    q = T(golden[f"{r}/dp/q"])
    sv, ev = torch.tensor(golden[f"{r}/dp/self_v"]), torch.tensor(golden[f"{r}/dp/env_v"])
    assert torch.equal(S.dp_search(m, q, sv, ev)[0], S.dp_search_slow(m, q, sv, ev))


@pytest.mark.parametrize("r", ROBOTS)
def test_pose_error_and_pose_step(golden, r):
    m = R.get_model(r)
    x, tgt = T(golden[f"{r}/lm/x"]), T(golden[f"{r}/lm/target"])
    e, cur = L.get_6d_pose_errors(m, x, tgt)
    np.testing.assert_allclose(e.numpy(), golden[f"{r}/lm/pose_err"], atol=1e-6)
    np.testing.assert_allclose(cur.numpy(), golden[f"{r}/lm/cur_pose"], atol=1e-6)
    xn, J, e = L.levenberg_marquardt_only_pose(m, x, tgt, L.ALT_LOSS_V2_1_POSE, return_residual=True)
    np.testing.assert_allclose(J.numpy(), golden[f"{r}/lm/pose_step_J"], atol=1e-6)
    np.testing.assert_allclose(e.numpy(), golden[f"{r}/lm/pose_step_e"], atol=1e-6)
    np.testing.assert_allclose(xn.numpy(), golden[f"{r}/lm/pose_step_x"], atol=2e-4)


@pytest.mark.parametrize("r", ROBOTS)
def test_differencing_step(golden, r):
    m = R.get_model(r)
    x, tgt = T(golden[f"{r}/lm/x"]), T(golden[f"{r}/lm/target"])
    Tc, c = cub(golden, r)
    pms = L.LmParams(virtual_configs=x.clone())
    Jd, rd = L.get_r_and_J(pms, m, x, tgt, Tc, c)
    J, rr = L.stack_rows(Jd), L.stack_rows(rd)
    assert J.shape == golden[f"{r}/lm/diff_step_J"].shape
    np.testing.assert_allclose(J.numpy(), golden[f"{r}/lm/diff_step_J"], atol=1e-7)
    np.testing.assert_allclose(rr.numpy(), golden[f"{r}/lm/diff_step_r"], atol=1e-7)
    xn = L.lm_full_step(J, rr, x, pms.lm_lambda)
    np.testing.assert_allclose(xn.numpy(), golden[f"{r}/lm/diff_step_x"], atol=1e-5)


@pytest.mark.parametrize("r", ROBOTS)
def test_all_terms_r_and_J(golden, r):
    m = R.get_model(r)
    x, xv, tgt = T(golden[f"{r}/lm/all_x"]), T(golden[f"{r}/lm/all_xv"]), T(golden[f"{r}/lm/target"])
    Tc, c = cub(golden, r)
    pms = L.LmParams(use_pose=True, virtual_configs=xv)
    Jd, rd = L.get_r_and_J(pms, m, x, tgt, Tc, c)
    assert rd["self_collisions"].shape[0] == int(golden[f"{r}/lm/all_n_self"]) > 0
    assert rd["env_collisions"].shape[0] == int(golden[f"{r}/lm/all_n_env"]) > 0
    J, rr = L.stack_rows(Jd), L.stack_rows(rd)
    np.testing.assert_allclose(J.numpy(), golden[f"{r}/lm/all_J"], atol=2e-6)
    np.testing.assert_allclose(rr.numpy(), golden[f"{r}/lm/all_r"], atol=2e-6)
    xn = L.lm_full_step(J, rr, x, pms.lm_lambda)
    np.testing.assert_allclose(xn.numpy(), golden[f"{r}/lm/all_step_x"], atol=5e-3)


@pytest.mark.parametrize("r", ROBOTS)
def test_collision_flags_and_metrics(golden, r):
    m = R.get_model(r)
    q = T(golden[f"{r}/cd/q"])
    Tc, c = cub(golden, r)
    assert np.array_equal(S.qpaths_batched_self_collisions(m, q).numpy(), golden[f"{r}/cd/self"])
    assert np.array_equal(S.qpaths_batched_env_collisions(m, q, c, Tc).numpy(), golden[f"{r}/cd/env"])
    assert golden[f"{r}/cd/self"].any() and golden[f"{r}/cd/env"].any()
    x, tgt = T(golden[f"{r}/lm/x"]), T(golden[f"{r}/lm/target"])
    ecm, edeg = L.calculate_pose_error_cm_deg(m, x, tgt)
    np.testing.assert_allclose(ecm.numpy(), golden[f"{r}/ev/err_cm"], atol=1e-5)
    np.testing.assert_allclose(edeg.numpy(), golden[f"{r}/ev/err_deg"], atol=1e-4)
    np.testing.assert_array_equal(L.angular_changes(x).numpy(), golden[f"{r}/ev/angular_changes"])
    np.testing.assert_array_equal(L.clamp_to_joint_limits(m, T(golden[f"{r}/lm/clamp_in"])).numpy(),
                                  golden[f"{r}/lm/clamp_out"])


# ------------------------------------------------------------------------------------------------------------
# known-answer vectors held by the reference's own tests


def test_angular_changes_known_answers():
    """tests/evaluation_utils_test.py:17-124"""
    P2 = 2 * math.pi
    cases = [
        ([[0, 0, 0], [0, 0, 0], [0, 0, 0]], [[0, 0, 0], [0, 0, 0]]),
        ([[0, 0, 0], [0, 0, 0], [0, 0, 0.1]], [[0, 0, 0], [0, 0, 0.1]]),
        ([[0, 0, 0], [0, 0, 0.1], [0, 0, -0.1]], [[0, 0, 0.1], [0, 0, -0.2]]),
        ([[0, -0.05, 0], [0, 0, 0.1], [0, 0, -0.1]], [[0, 0.05, 0.1], [0, 0, -0.2]]),
        ([[0, 0, 0], [0, 0, P2 - 0.1], [0, 0, 0]], [[0, 0, -0.1], [0, 0, 0.1]]),
        ([[0, 0, 0], [0, 0, P2 - 0.1], [-0.5, 0, 0.2]], [[0, 0, -0.1], [-0.5, 0, 0.3]]),
    ]
    for qpath, expected in cases:
        torch.testing.assert_close(T(expected), L.angular_changes(T(qpath)))


def test_joint_limit_almost_violations_known_answer():
    """tests/search_test.py:22-57 (Fetch limits spelled out at :35-42)"""
    m = R.get_model("fetch")
    assert m.actuated_joints_limits == [(0, 0.38615), (-1.6056, 1.6056), (-1.221, 1.518), (-math.pi, math.pi),
                                        (-2.251, 2.251), (-math.pi, math.pi), (-2.16, 2.16), (-math.pi, math.pi)]
    qs = torch.zeros((2, 3, 8))
    qs[0, 0] = T([0.051, 0, 0, 0, 0, 0, 0, 0])
    qs[0, 1] = T([0.38615 - 0.001, 0, 0, 0, 0, 0, 0, 0])
    qs[0, 2] = T([0.38615 - 0.051, 0, 0, 0, 0, 0, 0, 0])
    qs[1, 0] = T([0.38615 - 0.051, 0, 0, -np.pi, 0, 0, 0, 0])
    qs[1, 1] = T([0.38615 - 0.051, 0, 0, -np.pi + 0.11, 0, 0, 0, 0])
    qs[1, 2] = T([0.38615 - 0.051, 0, 0, -np.pi + 0.11, 0, 0, 0, np.pi - 0.25])
    expected = T([[0, 1, 0], [1, 0, 0]])
    torch.testing.assert_close(S.joint_limit_almost_violations_3d(m, qs, eps_revolute=0.1, eps_prismatic=0.05), expected)


def test_row_conventions():
    """tests/optimization_utils_test.py:67-119: Fetch prismatic idx 0, Panda all revolute."""
    assert R.get_model("fetch").prismatic_joint_idxs == [0]
    assert R.get_model("panda").prismatic_joint_idxs == []
    assert R.get_model("fetch_arm").ndof == 7 and R.get_model("fetch_arm").prismatic_joint_idxs == []


def test_fetch_torso_moves_ee_z_only():
    """tests/optimization_utils_test.py:377-402: moving joint 0 by d changes EE z by d, no rotation."""
    m = R.get_model("fetch")
    x = torch.zeros((1, 8), dtype=torch.float64)
    x[0, 1:] = T([0.3, -0.2, 0.5, 1.0, -0.4, 0.6, 0.1], torch.float64)
    x2 = x.clone()
    x2[0, 0] += 0.1
    p1, p2 = K.forward_kinematics(m, x), K.forward_kinematics(m, x2)
    torch.testing.assert_close(p2[:, :3] - p1[:, :3], T([[0, 0, 0.1]], torch.float64))
    torch.testing.assert_close(p2[:, 3:], p1[:, 3:])
    # and the pose residual is alpha_position * delta on the z row, zero rotation (row order [rot x3, pos x3])
    e, _ = L.get_6d_pose_errors(m, x, p2)
    torch.testing.assert_close(e[0, :, 0], T([0, 0, 0, 0, 0, 0.1], torch.float64), atol=1e-12, rtol=0)


def test_panda_fk_known_point():
    """tests/planners_test.py:282-309: q0 -> [0.45, 0.5422, 0.7885 | 1,0,0,0] within 1e-3."""
    m = R.get_model("panda")
    q0 = T([[1.267967, 0.711829, -0.811080, -0.810924, -2.637594, 1.767759, 0.083284]], torch.float64)
    pose = K.forward_kinematics(m, q0)[0]
    np.testing.assert_allclose(pose[:3].numpy(), [0.45, 0.5421984559194368, 0.7885155964931997], atol=1e-3)
    np.testing.assert_allclose(pose[3:].numpy(), [1, 0, 0, 0], atol=1e-3)


def test_differencing_residual_is_row_major_delta():
    """tests/optimization_utils_test.py:590-637"""
    m = R.get_model("panda")
    x = torch.arange(21, dtype=torch.float32).reshape(3, 7) * 0.01
    _, r = L.get_r_and_J(L.LmParams(use_virtual_configs=False, use_self_collisions=False, use_env_collisions=False,
                                    alpha_differencing=1.0), m, x, None)
    torch.testing.assert_close(r["differencing"][:, 0], (x[1:] - x[:-1]).reshape(-1))


@pytest.mark.parametrize("r", ROBOTS)
def test_batched_pose_lm_equals_dense_lm(golden, r):
    """tests/optimization_test.py:74-100: batched pose step == dense full step on pose-only params
    (J, r atol 1e-5; x atol 5e-3).  J and r are compared in the reference's fp32; x in fp64, because in fp32 the
    two solvers only agree up to the null-space noise measured in test_reference_fp32_null_space_noise."""
    m = R.get_model(r)
    x, tgt = T(golden[f"{r}/lm/x"]), T(golden[f"{r}/lm/target"])
    xb, Jb, eb = L.levenberg_marquardt_only_pose(m, x, tgt, L.ALT_LOSS_V2_1_POSE, return_residual=True)
    Jd, rd = L.get_r_and_J(L.ALT_LOSS_V2_1_POSE, m, x, tgt)
    n, D = x.shape
    for i in range(n):
        torch.testing.assert_close(Jd["pose"][6 * i : 6 * i + 6, D * i : D * i + D], Jb[i], atol=1e-5, rtol=0)
    torch.testing.assert_close(rd["pose"].reshape(n, 6, 1), eb, atol=1e-5, rtol=0)
    x64, t64 = x.double(), tgt.double()
    xb = L.levenberg_marquardt_only_pose(m, x64, t64, L.ALT_LOSS_V2_1_POSE)
    Jd, rd = L.get_r_and_J(L.ALT_LOSS_V2_1_POSE, m, x64, t64)
    xd = L.lm_full_step(L.stack_rows(Jd), L.stack_rows(rd), x64, 1e-6)
    torch.testing.assert_close(xd, xb, atol=1e-6, rtol=0)


def test_reference_fp32_null_space_noise(golden):
    """DESIGN.md 'numerics': with lambda = 1e-6 the fp32 normal equations of a 7/8-dof arm carry rounding noise in
    the null space of J that 1/lambda amplifies; the reference's own fp32 step is > 1e-3 rad away from exact
    arithmetic, which is why LM parity is defined against the fp64 oracle."""
    m = R.get_model("fetch")
    x, tgt = T(golden["fetch/lm/x"]), T(golden["fetch/lm/target"])
    x32 = L.levenberg_marquardt_only_pose(m, x, tgt, L.ALT_LOSS_V2_1_POSE)
    x64 = L.levenberg_marquardt_only_pose(m, x.double(), tgt.double(), L.ALT_LOSS_V2_1_POSE)
    assert (x32.double() - x64).abs().max() > 1e-3
    # ... while both reduce the pose error equally well (the noise lives in the null space)
    e32, _ = L.get_6d_pose_errors(m, x32.double(), tgt.double())
    e64, _ = L.get_6d_pose_errors(m, x64, tgt.double())
    assert abs(e32.abs().max() - e64.abs().max()) < 1e-3


def test_batched_equals_per_path_collision_flags(golden):
    """tests/collision_checking_test.py:43-56"""
    m = R.get_model("panda")
    q = T(golden["panda/cd/q"])
    Tc, c = cub(golden, "panda")
    b_self = S.qpaths_batched_self_collisions(m, q)
    b_env = S.qpaths_batched_env_collisions(m, q, c, Tc)
    for i in range(q.shape[0]):
        assert torch.equal(b_self[i], G.self_collision_distances(m, q[i]).min(dim=1).values < 0)
        e = torch.zeros(q.shape[1], dtype=torch.bool)
        for ci, Ti in zip(c, Tc):
            e |= G.env_collision_distances(m, q[i], ci, Ti).min(dim=1).values < 0
        assert torch.equal(b_env[i], e)


def test_geodesic_floor():
    q = T([[1.0, 0, 0, 0]], torch.float64)
    assert abs(geodesic_distance_between_quaternions(q, q).item() - 2 * math.acos(1 - 1e-7)) < 1e-9


def test_geodesic_distance_is_sign_invariant():
    """q and -q are the same rotation (jrl folds 2 acos(dot) into [0, pi])."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn((50, 4), generator=g, dtype=torch.float64)
    b = torch.randn((50, 4), generator=g, dtype=torch.float64)
    a, b = a / a.norm(dim=1, keepdim=True), b / b.norm(dim=1, keepdim=True)
    d = geodesic_distance_between_quaternions(a, b)
    assert torch.allclose(d, geodesic_distance_between_quaternions(a, -b), atol=1e-6)
    assert (d >= 0).all() and (d <= math.pi + 1e-9).all()


LOOP_PATTERNS = {  # planners.py:402-422
    "normal": dict(max_n_steps=20, return_if_valid_after_n_steps=0, convergence_threshold=1e6),
    "anytime": dict(max_n_steps=75, return_if_valid_after_n_steps=int(1e8), convergence_threshold=0.005),
}
LOOP_CASES = ["fetch_smooth", "fetch_noisy", "fetch_arm_smooth", "panda_smooth", "panda_noisy", "fetch_colliding"]


@pytest.mark.parametrize("case", LOOP_CASES)
@pytest.mark.parametrize("dt_name", ["f32", "f64"])
def test_alternating_loop_matches_reference_loop(loop_golden, case, dt_name):
    """oracle.lm.run_lm_alternating_loss == the reference's own run_lm_optimization / run_lm_alternating_loss
    (optimization.py:147-426, run by tests/golden/make_golden_loop.py): same step sequence, n_steps_taken, is_valid
    and returned iterate, in the reference's float32 and in float64, for both of the planner's call patterns."""
    g = loop_golden
    dt = torch.float32 if dt_name == "f32" else torch.float64
    m = R.get_model(str(g[f"{case}/robot"]))
    cuboids = [T(c, dt) for c in g[f"{case}/cuboids"]]
    Tcuboids = [T(t, dt) for t in g[f"{case}/Tcuboids"]]
    for pat, kw in LOOP_PATTERNS.items():
        key = f"{case}/{pat}/{dt_name}"
        x, n, valid, sched = L.run_lm_alternating_loss(m, T(g[f"{case}/x_seed"], dt), T(g[f"{case}/target"], dt),
                                                       (0.01, 0.1, 7.0, 2.0), Tcuboids=Tcuboids, cuboids=cuboids, **kw)
        assert sched == str(g[f"{key}/schedule"]), key
        assert n == int(g[f"{key}/n_steps_taken"]) and valid == bool(g[f"{key}/is_valid"]), key
        np.testing.assert_allclose(x.numpy(), g[f"{key}/x_opt"], atol=1e-6 if dt_name == "f32" else 1e-12)
    # the golden set covers: valid after the first pose steps, convergence exit, never valid
    assert bool(g["fetch_smooth/normal/f32/is_valid"]) and not bool(g["fetch_colliding/anytime/f32/is_valid"])


@pytest.fixture(scope="module")
def rowscale_golden():
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_rowscale_golden.npz"))


@pytest.mark.parametrize("key", ["fetch/unit", "fetch/random", "panda/random"])
def test_differencing_row_scaling_and_filtering(rowscale_golden, key):
    """SURVEY 8c golden items 4 and 6: _scale_down_rows_from_r_J_differencing_below_error (optimization_utils.py:352-397)
    and filter_rows_from_r_J_differencing (:736-768), with and without the shift, against the reference's outputs;
    `fetch/unit` is the input of the reference's own test (tests/optimization_utils_test.py:126-155)."""
    g = rowscale_golden
    m = R.get_model(str(g[f"{key}/robot"]))
    for shift in (0, 1):
        J, r, inv = L.scale_down_rows_differencing_below_error(m, T(g[f"{key}/diff/r_in"]), T(g[f"{key}/diff/J_in"]), 0.25, 1.5,
                                                               0.5, bool(shift))
        assert np.array_equal(r.numpy(), g[f"{key}/diff/scale_shift{shift}/r"])
        assert np.array_equal(J.numpy(), g[f"{key}/diff/scale_shift{shift}/J"])
        assert np.array_equal(inv.numpy(), g[f"{key}/diff/scale_shift{shift}/invalid"])
        r, J = L.filter_rows_from_r_J_differencing(m, T(g[f"{key}/diff/r_in"]), T(g[f"{key}/diff/J_in"]), 1.5, 0.25, bool(shift))
        assert np.array_equal(r.numpy(), g[f"{key}/diff/filter_shift{shift}/r"])
        assert np.array_equal(J.numpy(), g[f"{key}/diff/filter_shift{shift}/J"])
    if key == "fetch/unit":  # the expected values spelled out in the reference's test (:137-186)
        _, r, inv = L.scale_down_rows_differencing_below_error(m, T(g[f"{key}/diff/r_in"]), T(g[f"{key}/diff/J_in"]), 0.25, 1.5, 0.5)
        assert inv.nonzero()[:, 0].tolist() == [0, 2, 8, 9, 10]
        assert r[:, 0].tolist() == pytest.approx([0.5, 0.05, 1.6] + [0.05] * 5 + [-0.4, 1.7, -1.7] + [0.05] * 5 + [0.1, 0.005] + [0.05] * 6)


@pytest.mark.parametrize("key", ["pose/a", "pose/b"])
def test_pose_row_scaling(rowscale_golden, key):
    """SURVEY 8c golden item 5: _scale_down_rows_from_r_J_pose_below_error (optimization_utils.py:288-329)."""
    g = rowscale_golden
    r, J, inv = L.scale_down_rows_pose_below_error(T(g[f"{key}/r_in"]), T(g[f"{key}/J_in"]), 0.01, 0.03, 0.25)
    assert np.array_equal(r.numpy(), g[f"{key}/r"]) and np.array_equal(J.numpy(), g[f"{key}/J"])
    assert np.array_equal(inv.numpy(), g[f"{key}/invalid"])
    assert 0 < int(inv.sum()) < inv.numel()


def test_row_scaling_options_in_dense_get_r_and_J():
    """The options change the dense system the way the reference's get_r_and_J composes them: scaled rows before the
    alpha factors, filtered differencing rows dropped (and no prismatic scaling then, optimization_utils.py:606)."""
    from dataclasses import replace

    m = R.get_model("fetch")
    g = torch.Generator().manual_seed(3)
    x = torch.tensor(np.asarray(m.actuated_joints_limits)).float().mean(dim=1)[None] + 0.2 * torch.randn((12, 8), generator=g)
    target = K.forward_kinematics(m, x.double() + 0.01).float()
    base = L.LmParams(use_pose=True, use_virtual_configs=False, use_self_collisions=False, use_env_collisions=False,
                      alpha_differencing_prismatic_scaling=2.0, constraints=(0.01, 0.1, 7.0, 2.0))
    J0, r0 = L.get_r_and_J(base, m, x, target)
    J1, r1 = L.get_r_and_J(replace(base, differencing_do_scale_satisfied=True, differencing_ignore_satisfied_margin_deg=1.0,
                                   differencing_ignore_satisfied_margin_cm=0.5), m, x, target)
    small = (r0["differencing"].abs() < 1e-9) | (r1["differencing"].abs() <= r0["differencing"].abs() + 1e-12)
    assert small.all() and not torch.equal(r0["differencing"], r1["differencing"])
    J2, r2 = L.get_r_and_J(replace(base, differencing_do_ignore_satisfied=True, differencing_ignore_satisfied_margin_deg=1.0,
                                   differencing_ignore_satisfied_margin_cm=0.5), m, x, target)
    assert 0 < r2["differencing"].shape[0] < r0["differencing"].shape[0] and J2["differencing"].shape[0] == r2["differencing"].shape[0]
    J3, r3 = L.get_r_and_J(replace(base, pose_do_scale_down_satisfied=True, pose_ignore_satisfied_threshold_scale=500.0,
                                   pose_ignore_satisfied_scale_down=0.5), m, x, target)
    ratio = r3["pose"] / r0["pose"]
    assert set(np.round(ratio[torch.isfinite(ratio)].numpy(), 4).tolist()) <= {0.5, 1.0} and (ratio == 0.5).any()
