"""CPU: host-side logic that does not need a GPU - parameter marshalling, validity decisions, problem loading,
sharding, and the multi-rank cost gather over gloo (world_size 2)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from cppflow_b200 import ops
from cppflow_b200.data_types import Constraints, DEFAULT_CONSTRAINTS
from cppflow_b200.lm_hyper_parameters import ALT_LOSS_V2_1_DIFF, ALT_LOSS_V2_1_POSE, OptimizationParameters, all_terms_parameters


def test_make_params_marshalling():
    p = ops.make_params(ALT_LOSS_V2_1_DIFF)
    assert (p.use_pose, p.use_differencing, p.use_virtual_configs, p.n_virtual_configs) == (0, 1, 1, 4)
    assert (p.use_self_collisions, p.use_env_collisions) == (1, 1)
    assert abs(p.alpha_differencing - 0.00375) < 1e-9 and abs(p.lm_lambda - 1e-6) < 1e-12 and p.alpha_position == 0.0
    p = ops.make_params(ALT_LOSS_V2_1_POSE)
    assert (p.use_pose, p.use_differencing, p.use_virtual_configs, p.n_virtual_configs) == (1, 0, 0, 0)
    assert (p.alpha_position, p.alpha_rotation) == (3.5, pytest.approx(0.35))
    p = ops.make_params(all_terms_parameters())
    assert p.use_pose == 1 and p.use_differencing == 1 and p.use_self_collisions == 1
    bad = OptimizationParameters(**{**ALT_LOSS_V2_1_DIFF.__dict__, "differencing_do_scale_satisfied": True,
                                    "differencing_ignore_satisfied_margin_deg": 1.0,
                                    "differencing_ignore_satisfied_margin_cm": 1.0})
    with pytest.raises(NotImplementedError):
        ops.make_params(bad)


def test_obstacle_tables():
    c = torch.tensor([-0.1, -0.2, -0.3, 0.1, 0.2, 0.3])
    T = torch.zeros((4, 4))
    T[:3, :3] = torch.eye(3)
    T[:3, 3] = torch.tensor([1.0, 2.0, 3.0])
    ob = ops.Obstacles([c, c], [T, T])
    assert ob.n == 2 and list(ob.cuboids_ptr)[:6] == pytest.approx(c.tolist())
    assert list(ob.Tcuboids_ptr)[3] == 1.0 and list(ob.single(1).Tcuboids_ptr)[7] == 2.0
    assert ops._obs(None) == (None, None, 0) and ops._obs(ops.Obstacles()) == (None, None, 0)
    with pytest.raises(AssertionError):
        ops.Obstacles([c] * 9, [T] * 9)


class _FakeRobot:
    ndof, robot_id, name, formal_robot_name = 8, 0, "fetch", "Fetch"


def _problem(T=10):
    from cppflow_b200.data_types import Problem

    tp = torch.zeros((T, 7))
    tp[:, 3] = 1.0
    return Problem(DEFAULT_CONSTRAINTS, tp, None, _FakeRobot(), "t", "t", [], [], [], [])


def test_x_is_valid_decisions():
    """optimization_utils.py:836-923: first path passing every threshold and the collision check wins."""
    from cppflow_b200.optimization_utils import x_is_valid

    prob = _problem()
    x = torch.arange(3 * 10 * 8, dtype=torch.float32).reshape(30, 8)
    #            pos_cm rot_deg mjac_deg mjac_cm tl  min_self min_env pad
    metrics = torch.tensor([[0.5, 0.01, 1.0, 0.1, 3.0, 0.1, 0.1, 0],      # position error too large
                            [0.001, 0.01, 1.0, 0.1, 3.0, -0.01, 0.1, 0],  # self collision
                            [0.001, 0.01, 1.0, 0.1, 3.0, 0.02, 0.03, 0]])
    x_sol, idx, flags = x_is_valid(prob, DEFAULT_CONSTRAINTS, None, x, 3, metrics=metrics)
    assert idx == 2 and torch.equal(x_sol, x[20:30]) and flags == (True, True, True, True, False, False)
    x_sol, idx, flags = x_is_valid(prob, DEFAULT_CONSTRAINTS, None, x[:20], 2, metrics=metrics[:2])
    assert x_sol is None and idx is None and flags[4] is True
    x_sol, idx, flags = x_is_valid(prob, DEFAULT_CONSTRAINTS, None, x[:10], 1, metrics=metrics[:1])
    assert x_sol is None and flags[:4] == (False, True, True, True) and flags[4] is None
    # thresholds are strict '<' (evaluation_utils.py:41-58)
    edge = torch.tensor([[0.5, 0.01, 1.0, 0.1, 3.0, 0.1, 0.1, 0]])
    assert x_is_valid(prob, Constraints(0.5, 0.1, 7.0, 2.0), None, x[:10], 1, metrics=edge)[0] is None
    assert x_is_valid(prob, Constraints(0.5001, 0.1, 7.0, 2.0), None, x[:10], 1, metrics=edge)[0] is not None
    # a mesh-level validator (klampt in the reference) overrides the capsule verdict
    x_sol, _, flags = x_is_valid(prob, DEFAULT_CONSTRAINTS, None, x[:20], 2, metrics=metrics[:2],
                                 mesh_validator=lambda p, xi: (False, False))
    assert x_sol is not None and flags[4] is False


def test_problem_loader_matches_reference_files():
    """data_type_utils.py:148-219 / SURVEY 8d config 4: 13 benchmark problems, sum T = 4036."""
    from cppflow_b200.data_type_utils import ALL_PROBLEM_FILENAMES, problem_from_filename

    expected_T = {"hello": 553, "circle": 295, "rot_yz2": 249, "s": 301, "square": 320, "1cube": 200, "2cubes": 200,
                  "flappy_bird": 200}
    total = 0
    for name in ALL_PROBLEM_FILENAMES:
        p = problem_from_filename(None, name, device="cpu")
        assert p.n_timesteps == expected_T[p.name] and p.target_path.shape == (p.n_timesteps, 7)
        assert torch.allclose(p.target_path[:, 3:].norm(dim=1), torch.ones(p.n_timesteps), atol=1e-3)
        assert len(p.obstacles_cuboids) == len(p.obstacles_Tcuboids) == p.obstacle_tables.n
        total += p.n_timesteps
    assert total == 4036
    p = problem_from_filename(None, "fetch__circle", device="cpu")
    # problems/fetch__circle.yaml: offset (0.9, 0.25, 0.46) from torso_lift_link at q = 0
    np.testing.assert_allclose(p.target_path[0, :3].numpy(), [0.9 - 0.086875, 0.25, 0.46 + 0.37743], atol=1e-6)
    assert p.obstacles_cuboids[0].tolist() == pytest.approx([-0.15, -0.025, -0.4, 0.15, 0.025, 0.4])
    assert p.obstacles_Tcuboids[2][:3, 3].tolist() == pytest.approx([0.4, 0.0, 1.225]) and p.obstacles_Tcuboids[0][3, 3] == 0
    p = problem_from_filename(None, "fetch__s", device="cpu")  # obstacle_xyz_offset (0, 0, -0.2)
    assert p.obstacles_Tcuboids[0][2, 3].item() == pytest.approx(0.36 - 0.2)
    with pytest.raises(ValueError):
        from cppflow_b200.data_types import Problem

        Problem(DEFAULT_CONSTRAINTS, torch.zeros((4, 7)), None, _FakeRobot(), "bad", "bad", [], [], [], [])


def test_problem_from_user_yaml_and_csv(tmp_path):
    """data_type_utils.py:167-173 (`filepath_override`, the route tests/planners_test.py:96 takes): a problem yaml in
    the reference's layout gives the same Problem as the packed definition; a path csv beside the yaml is picked up."""
    from cppflow_b200.data_type_utils import _problems, problem_from_filename

    d = _problems()["fetch__s"]
    lines = [f"robot: {d['robot']}", f"path_name: {d['path_name']}", f"path_offset_frame: {d['path_offset_frame']}",
             f"path_xyz_offset: {d['path_xyz_offset']}", "path_R_offset:"]
    lines += [f"  - {row}" for row in d["path_R_offset"]]
    lines += [f"obstacle_xyz_offset: {d['obstacle_xyz_offset']}", "obstacles:"]
    for obs in d["obstacles"]:
        items = list(obs.items())
        lines.append(f"  - - {items[0][0]}: {items[0][1]}")
        lines += [f"    - {k}: {v}" for k, v in items[1:]]
    yaml_path = tmp_path / "my_problem.yaml"
    yaml_path.write_text("\n".join(lines) + "\n")
    packed = problem_from_filename(None, "fetch__s", device="cpu")
    loaded = problem_from_filename(None, "", filepath_override=str(yaml_path), device="cpu")
    assert torch.equal(packed.target_path, loaded.target_path) and loaded.robot.name == packed.robot.name
    assert len(loaded.obstacles_cuboids) == len(packed.obstacles_cuboids) > 0
    for a, b in zip(packed.obstacles_cuboids + packed.obstacles_Tcuboids, loaded.obstacles_cuboids + loaded.obstacles_Tcuboids):
        assert torch.equal(a, b)
    with pytest.raises(AssertionError):  # :184 - a caller-provided robot excludes obstacles
        problem_from_filename(None, "", filepath_override=str(yaml_path), robot=packed.robot, device="cpu")

    # a user path: csv beside the yaml, header + time,x,y,z,qw,qx,qy,qz
    rows = ["time,x,y,z,qw,qx,qy,qz"] + [f"{0.1 * i},{0.01 * i},0.2,0.3,1.0,0.0,0.0,0.0" for i in range(5)]
    (tmp_path / "my_line.csv").write_text("\n".join(rows) + "\n")
    (tmp_path / "line.yaml").write_text(
        "robot: panda\npath_name: my_line\npath_offset_frame: world\npath_xyz_offset: [0.4, 0.0, 0.1]\n"
        "path_R_offset:\n  - [1, 0, 0]\n  - [0, 1, 0]\n  - [0, 0, 1]\n")
    p = problem_from_filename(None, "", filepath_override=str(tmp_path / "line.yaml"), device="cpu")
    assert p.n_timesteps == 5 and p.obstacle_tables.n == 0 and p.robot.name == "panda"
    np.testing.assert_allclose(p.target_path[4].numpy(), [0.44, 0.2, 0.4, 1, 0, 0, 0], atol=1e-6)
    (tmp_path / "missing.yaml").write_text((tmp_path / "line.yaml").read_text().replace("my_line", "nowhere"))
    with pytest.raises(FileNotFoundError):
        problem_from_filename(None, "", filepath_override=str(tmp_path / "missing.yaml"), device="cpu")


def test_shard_range_and_costs():
    from cppflow_b200.distributed import shard_range, path_costs, path_keys, decode_key, INVALID_COST

    for n, w in [(8192, 8), (13, 4), (3, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    m = torch.tensor([[0.001, 0.01, 1.0, 0.1, 3.0, 0.1, 0.1, 0], [0.5, 0.01, 1.0, 0.1, 2.0, 0.1, 0.1, 0],
                      [0.001, 0.01, 1.0, 0.1, 2.5, 0.1, -0.1, 0]])
    c = path_costs(m, DEFAULT_CONSTRAINTS)
    assert c.dtype == torch.float64
    assert c.tolist() == pytest.approx([3.0, 2.0 + INVALID_COST, 2.5 + INVALID_COST], abs=1e-6)
    # the packed keys order like (invalid, TL, index) and survive the round trip
    keys = path_keys(m, DEFAULT_CONSTRAINTS, first_index=100).tolist()
    assert [decode_key(k) for k in keys] == [(True, 3.0, 100), (False, 2.0, 101), (False, 2.5, 102)]
    assert sorted(range(3), key=lambda i: keys[i]) == [0, 1, 2]
    # among invalid paths the trajectory length still decides (float32 TL + 1e9 could not tell 2.0 from 2.5)
    assert keys[1] < keys[2] and float(torch.tensor(2.0) + 1e9) == float(torch.tensor(2.5) + 1e9)
    nan = m.clone()
    nan[0, 4] = float("nan")
    assert decode_key(int(path_keys(nan, DEFAULT_CONSTRAINTS)[0]))[0] is False


def _gather_worker(rank, world, port, out, all_invalid):
    import torch.distributed as dist
    from cppflow_b200.distributed import enqueue_argmin, gather_costs_and_argmin, shard_range

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    metrics = _gather_metrics(all_invalid)
    s, e = shard_range(metrics.shape[0], rank, world)
    best = gather_costs_and_argmin(metrics[s:e], DEFAULT_CONSTRAINTS, rank, world, first_index=s)
    full = enqueue_argmin(metrics[s:e], DEFAULT_CONSTRAINTS, s, world).result()
    out[rank] = (best, (s, e), tuple(full))
    dist.destroy_process_group()


def _gather_metrics(all_invalid):
    g = torch.Generator().manual_seed(0)
    P = 37
    metrics = torch.zeros((P, 8))
    metrics[:, 4] = torch.rand(P, generator=g) + 1.0  # trajectory lengths
    metrics[:, 5:7] = 0.1
    metrics[5, 4] = metrics[29, 4] = 0.5              # tie across shards -> lowest global index wins
    if all_invalid:
        metrics[:, 0] = 1.0                           # every path breaks the position threshold
        metrics[20, 4] = 0.25                         # ... and the shortest of them lives in the second shard
    else:
        metrics[20, 4] = 0.25                         # shorter than the tie but invalid
        metrics[20, 6] = -0.01
    return metrics


@pytest.mark.parametrize("all_invalid", [False, True])
def test_gather_argmin_two_ranks_gloo(all_invalid):
    """the argmin is invariant to the shard count - also when every path is invalid, where the ranking
    falls back to the trajectory length among invalid paths (round 1's float32 `TL + 1e9` lost it to rounding)."""
    from cppflow_b200.distributed import gather_costs_and_argmin, INVALID_COST

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(2, port, out, all_invalid), nprocs=2, join=True)
    (b0, span0, f0), (b1, span1, f1) = out[0], out[1]
    assert b0 == b1 and f0 == f1
    cost, rank, idx = b0
    single = gather_costs_and_argmin(_gather_metrics(all_invalid), DEFAULT_CONSTRAINTS, 0, 1)
    if all_invalid:
        assert cost == pytest.approx(0.25 + INVALID_COST) and rank == 1 and span1[0] + idx == 20
        assert f0[3] == 0 and f0[4] is False  # n_valid, valid
        assert single == (pytest.approx(0.25 + INVALID_COST), 0, 20)
    else:
        assert cost == pytest.approx(0.5) and rank == 0 and span0[0] + idx == 5
        assert f0[3] == 36 and f0[4] is True and f0[6] == 5
        assert single == (pytest.approx(0.5), 0, 5)


def test_split_paths_covers_every_path_once():
    """pipeline.split_paths: contiguous chunks, multiples of the assembly CTA (256 paths) / the solve group (16) when
    the path count allows, no empty chunk, every path in exactly one chunk."""
    from cppflow_b200.pipeline import split_paths

    for n, c in [(8192, 4), (8192, 3), (8192, 1), (1000, 3), (300, 4), (17, 4), (5, 8), (1, 4), (4096, 16), (257, 2)]:
        chunks = split_paths(n, c)
        assert 1 <= len(chunks) <= c
        assert chunks[0][0] == 0 and sum(k for _, k in chunks) == n
        for (a, ka), (b, _) in zip(chunks, chunks[1:]):
            assert a + ka == b and ka > 0
        gran = 256 if n >= 256 * c else 16 if n >= 16 * c else 1
        assert all(k % gran == 0 for _, k in chunks[:-1])
    assert split_paths(8192, 4) == [(0, 2048), (2048, 2048), (4096, 2048), (6144, 2048)]


def test_numa_local_is_a_safe_no_op_without_topology():
    """pipeline.numa_local must never raise and must restore the affinity (no CUDA device / no sysfs entry here)."""
    import os

    from cppflow_b200.pipeline import numa_local

    before = os.sched_getaffinity(0)
    with numa_local("cuda:0"):
        pass
    assert os.sched_getaffinity(0) == before


def test_plan_metrics_match_the_reference_definitions():
    """Plan (data_types.py:86-348) on CPU tensors: derived metrics, validity and the numpy view."""
    from cppflow_b200.data_types import Plan, PlanNp, Constraints
    from cppflow_b200.evaluation_utils import positional_errors, rotational_errors
    from oracle import robots as OR, kinematics as OK, lm as OL

    m = OR.get_model("fetch")
    lim = torch.tensor(m.actuated_joints_limits)
    g = torch.Generator().manual_seed(0)
    T = 20
    steps = 0.02 * torch.randn((T, 8), generator=g)
    steps[:, 0] *= 0.2  # the prismatic torso: millimetres per waypoint
    q = lim.mean(dim=1)[None] + steps.cumsum(dim=0)
    target = OK.forward_kinematics(m, q.double()).float()
    q2 = q.clone()
    q2[7] += 0.004  # a pose error of a few mm and a joint jump
    traced = OK.forward_kinematics(m, q2.double()).float()
    plan = Plan(q_path=q2, q_path_revolute=q2[:, m.revolute_joint_idxs], q_path_prismatic=q2[:, m.prismatic_joint_idxs],
                pose_path=traced, target_path=target, robot_joint_limits=m.actuated_joints_limits,
                self_colliding_per_ts=torch.zeros(T, dtype=torch.bool), env_colliding_per_ts=torch.zeros(T, dtype=torch.bool),
                positional_errors=positional_errors(traced, target), rotational_errors=rotational_errors(traced, target),
                provided_initial_configuration=None, constraints=Constraints(0.01, 0.1, 7.0, 2.0))
    ref = OL.path_metrics(m, q2.double(), target.double())
    assert plan.max_positional_error_cm == pytest.approx(float(ref["max_pos_cm"]), abs=1e-4)
    assert plan.max_positional_error_mm == pytest.approx(10 * plan.max_positional_error_cm)
    assert plan.max_rotational_error_deg == pytest.approx(float(ref["max_rot_deg"]), abs=3e-2)
    assert plan.mjac_deg == pytest.approx(float(ref["mjac_deg"]), abs=1e-4)
    assert plan.mjac_cm == pytest.approx(float(ref["mjac_cm"]), abs=1e-4)
    assert plan.path_length_rad == pytest.approx(float(ref["tl"]), rel=1e-5)
    assert plan.path_length_m == pytest.approx(float((q2[1:, 0] - q2[:-1, 0]).abs().sum()), rel=1e-5)
    assert plan.is_a_prismatic_joint and not plan.joint_limits_violated and plan.initial_q_norm_dist == 0.0
    assert plan.is_valid is False and "errors_are_below_threshold(...): False" in plan.is_valid_(verbose=True)[1]
    assert "max positional error" in str(plan)
    assert isinstance(PlanNp(plan).positional_errors, np.ndarray) and PlanNp(plan).mjac_deg == plan.mjac_deg
    exact = Plan(q_path=q, q_path_revolute=q[:, m.revolute_joint_idxs], q_path_prismatic=q[:, m.prismatic_joint_idxs],
                 pose_path=target, target_path=target, robot_joint_limits=m.actuated_joints_limits,
                 self_colliding_per_ts=torch.zeros(T, dtype=torch.bool), env_colliding_per_ts=torch.zeros(T, dtype=torch.bool),
                 positional_errors=positional_errors(target, target), rotational_errors=rotational_errors(target, target),
                 provided_initial_configuration=q[0:1] + 0.5, constraints=Constraints(0.01, 0.1, 7.0, 2.0))
    assert exact.max_rotational_error_deg == pytest.approx(math.degrees(2 * math.acos(1 - 1e-7)), abs=1e-2)  # the clamp floor (0.056 in fp32)
    assert exact.is_valid is False  # too far from the requested initial configuration (0.5 * sqrt(8) > 0.2)
    exact.provided_initial_configuration = q[0:1]
    assert exact.is_valid is True
    colliding = exact
    colliding.env_colliding_per_ts = colliding.env_colliding_per_ts.clone()
    colliding.env_colliding_per_ts[3] = True
    assert colliding.is_valid is False


def test_problem_rotational_path_length_discounts_the_clamp_floor():
    """data_types.py:403-418: identical consecutive orientations contribute 0, not 2 acos(1 - 1e-7)."""
    from cppflow_b200.data_types import Problem

    T = 10
    target = torch.zeros((T, 7))
    target[:, 3] = 1.0
    target[:, 0] = torch.linspace(0, 0.09, T)
    p = Problem(DEFAULT_CONSTRAINTS, target, None, _FakeRobot(), "line", "fake__line", [], [], [], [])
    assert p.path_length_cumultive_positional_change_cm == pytest.approx(9.0, abs=1e-4)
    assert abs(p.path_length_cumulative_rotational_change_deg) < 1e-3
    half = math.sin(math.radians(5.0))
    target[5:, 3], target[5:, 6] = math.cos(math.radians(5.0)), half  # one 10 degree turn about z
    p = Problem(DEFAULT_CONSTRAINTS, target, None, _FakeRobot(), "line", "fake__line", [], [], [], [])
    assert p.path_length_cumulative_rotational_change_deg == pytest.approx(10.0, abs=2e-2)


def test_segmented_solve_model_equals_dense_solve():
    """The algebra of csrc/lm_segsolve.cuh (segments between separator waypoints eliminated up and down, the separators'
    reduced block-tridiagonal system with full couplings K = -beta Q, right-hand-side correction + twisted
    back-substitution inside the segments), restated in fp64 numpy (tools/segsolve_model.py), against a dense solve of
    the same SPD block-tridiagonal system - including one-block segments, a single segment, and the separator rule
    s_j = floor(j T / S) the kernels use."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from segsolve_model import dense, segments, solve_segmented

    rng = np.random.default_rng(7)
    for T, S, D in [(300, 16, 8), (37, 5, 7), (9, 4, 8), (300, 2, 8), (20, 1, 7), (8, 4, 8), (64, 16, 7)]:
        J = rng.normal(size=(T, 12, D))
        beta = np.abs(rng.normal(size=D)) + 0.5
        A = np.einsum("tki,tkj->tij", J, J) + 0.01 * np.eye(D) + 2 * np.diag(beta ** 2)
        b = rng.normal(size=(T, D))
        seps, segs = segments(T, S)
        assert len(seps) == S - 1 and all(e >= a for a, e in segs) and segs[0][0] == 0 and segs[-1][1] == T - 1
        assert sum(e - a + 1 for a, e in segs) + len(seps) == T
        np.testing.assert_allclose(solve_segmented(A, b, beta, S), dense(A, b, beta), atol=1e-12)


def test_latent_sampling_layout():
    """tests/planners_test.py:219-260 (`_get_fixed_random_latent`, per-k): [k * T, width] latents, one latent per path
    repeated over its T waypoints, k distinct ones; uniform in +-scale / 2 or gaussian; sampling near a centre latent keeps
    the centre as the first path's latent (planners.py:136-153)."""
    from cppflow_b200.planners import LatentIkCandidateGenerator

    k, T, width = 15, 300, 9
    for dist in ("uniform", "gaussian"):
        gen = LatentIkCandidateGenerator(seed=4, latent_distribution=dist, latent_vector_scale=1.5)
        latent = gen._sample_latents(k, T, width)
        assert latent.shape == (k * T, width)
        per_path = latent.reshape(k, T, width)
        assert torch.equal(per_path, per_path[:, :1].expand(k, T, width)), "a path's latent changes along the path"
        assert per_path[:, 0].unique(dim=0).shape[0] == k, "paths share a latent"
        if dist == "uniform":
            assert latent.min() >= -0.75 and latent.max() <= 0.75
            assert latent.min() < -0.6 and latent.max() > 0.6 and abs(float(latent.mean())) < 0.1
    gen = LatentIkCandidateGenerator(seed=4, latent_vector_scale=1.0)
    centre = torch.linspace(-0.3, 0.3, 7)
    near = gen._sample_latents_near(5, 11, centre).reshape(5, 11, 7)
    assert torch.equal(near[0], centre.expand(11, 7))
    assert ((near - centre).abs() <= 0.5 + 1e-6).all() and near[1:, 0].unique(dim=0).shape[0] == 4


def test_mjac_consistency():
    """tests/evaluation_utils_test.py:12-15: the three ways the reference computes the maximum joint-angle change of a
    path agree (evaluation_utils.py:113-141)."""
    from cppflow_b200.evaluation_utils import angular_changes, calculate_mjac_deg, calculate_per_timestep_mjac_deg

    torch.manual_seed(0)
    qpath = torch.randn((10, 3)) * 3.0  # beyond +-pi: the wrap matters
    assert calculate_mjac_deg(qpath) == pytest.approx(float(calculate_per_timestep_mjac_deg(qpath).max()), abs=1e-5)
    assert calculate_mjac_deg(qpath) == pytest.approx(float(torch.rad2deg(angular_changes(qpath).abs().max())), abs=1e-5)
    assert calculate_per_timestep_mjac_deg(qpath).shape == (9,)
