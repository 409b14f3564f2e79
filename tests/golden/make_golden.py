"""Generate golden vectors by running the REAL reference code (run once, in the build container).

    python tests/golden/make_golden.py          # writes tests/golden/*.npz

`/root/reference/cppflow` cannot be imported as-is: it needs `jrl`, `klampt`, `ikflow`, `matplotlib`
(none installed, no network - SURVEY.md 8c).  This script stubs those modules, backs the `jrl.robot.Robot`
stub with the oracle's kinematics/geometry (the jrl side stays "parity unpinned"), and then calls the
reference's own, unmodified functions:
    cppflow.search.dp_search / _get_mjacs / joint_limit_almost_violations_3d / dp_search_slow
    cppflow.optimization.levenberg_marquardt_only_pose / levenberg_marquardt_full / _lm_full_step
    cppflow.optimization_utils.LmResidualFns.get_r_and_J / get_6d_pose_errors / clamp_to_joint_limits
    cppflow.collision_detection.qpaths_batched_self_collisions / qpaths_batched_env_collisions
    cppflow.evaluation_utils.angular_changes / calculate_pose_error_cm_deg / errors_are_below_threshold
so the in-tree logic of the hot path is pinned by the reference itself.  The .npz files hold inputs and
outputs; nothing under /root/reference is read at test time.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import robots as orobots  # noqa: E402
from oracle import kinematics as OK  # noqa: E402
from oracle import geometry as OG  # noqa: E402
from oracle import math_utils as OM  # noqa: E402


# ----------------------------------------------------------------------------------------------------------
# stubs


class _Robot:
    """Stand-in for jrl.robot.Robot exposing what the hot path reads."""

    name = None

    def __init__(self):
        self._m = orobots.get_model(self.name)
        self._klampt_world_model = mock.MagicMock()
        self._collision_capsules_by_link = {c.link: c for c in self._m.capsules}

    ndof = property(lambda s: s._m.ndof)
    formal_robot_name = property(lambda s: s._m.formal_robot_name)
    actuated_joints_limits = property(lambda s: s._m.actuated_joints_limits)
    actuated_joint_names = property(lambda s: s._m.actuated_joint_names)
    prismatic_joint_idxs = property(lambda s: s._m.prismatic_joint_idxs)
    revolute_joint_idxs = property(lambda s: s._m.revolute_joint_idxs)
    has_prismatic_joints = property(lambda s: s._m.has_prismatic_joints)
    end_effector_link_name = property(lambda s: s._m.end_effector_link_name)

    def forward_kinematics(self, x, out_device=None, dtype=None):
        return OK.forward_kinematics(self._m, x)

    def jacobian(self, x):
        return OK.jacobian(self._m, x)

    def self_collision_distances(self, x):
        return OG.self_collision_distances(self._m, x)

    def self_collision_distances_jacobian(self, x):
        return OG.self_collision_distances(self._m, x, with_jacobian=True)[1]

    def env_collision_distances(self, x, cuboid, Tcuboid):
        return OG.env_collision_distances(self._m, x, cuboid, Tcuboid)

    def env_collision_distances_jacobian(self, x, cuboid, Tcuboid):
        return OG.env_collision_distances(self._m, x, cuboid, Tcuboid, with_jacobian=True)[1]

    def split_configs_to_revolute_and_prismatic(self, x):
        return x[:, self.revolute_joint_idxs], x[:, self.prismatic_joint_idxs]


class Fetch(_Robot):
    name = "fetch"


class FetchArm(_Robot):
    name = "fetch_arm"


class Panda(_Robot):
    name = "panda"


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("klampt", "matplotlib", "ikflow", "FrEIA", "pkg_resources", "rclpy", "cppflow_msgs")

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__name__ = spec.name
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


def install_stubs():
    sys.meta_path.insert(0, _StubFinder())
    jrl = types.ModuleType("jrl")
    jrl.__path__ = []
    cfg = types.ModuleType("jrl.config")
    cfg.DEVICE = "cpu"
    cfg.PT_NP_TYPE = object
    robot = types.ModuleType("jrl.robot")
    robot.Robot = _Robot
    robots = types.ModuleType("jrl.robots")
    robots.Fetch, robots.FetchArm, robots.Panda = Fetch, FetchArm, Panda
    robots.get_robot = lambda name: {"fetch": Fetch, "fetch_arm": FetchArm, "panda": Panda}[name]()
    mu = types.ModuleType("jrl.math_utils")
    for fn in ("quaternion_inverse", "quaternion_product", "quaternion_to_rpy", "angular_subtraction",
               "geodesic_distance_between_quaternions", "quaternion_norm", "quaternion_conjugate"):
        setattr(mu, fn, getattr(OM, fn))
    mu.rpy_tuple_to_rotation_matrix = lambda rpy: OM.rpy_to_rotation_matrix(rpy, dtype=torch.float32)
    utils = types.ModuleType("jrl.utils")
    utils.safe_mkdir = lambda p: None
    utils.to_torch = lambda x: torch.tensor(x, dtype=torch.float32)
    utils.set_seed = lambda *a, **k: None
    utils.mm_to_m = lambda x: x / 1000.0
    utils.make_text_green_or_red = lambda t, c: t
    for name, mod in (("jrl", jrl), ("jrl.config", cfg), ("jrl.robot", robot), ("jrl.robots", robots),
                      ("jrl.math_utils", mu), ("jrl.utils", utils)):
        sys.modules[name] = mod
    sys.path.insert(0, "/root/reference")


# ----------------------------------------------------------------------------------------------------------
# synthetic inputs (identical generator to tests/helpers.py)


def smooth_joint_path(model, T, seed, amp=0.25):
    g = np.random.default_rng(seed)
    lim = np.array(model.actuated_joints_limits)
    mid, half = lim.mean(1), (lim[:, 1] - lim[:, 0]) / 2
    t = np.linspace(0, 1, T)[:, None]
    q = mid + half * 0.3 * g.uniform(-1, 1, (1, model.ndof))
    for _ in range(3):
        q = q + half * amp * g.uniform(0.2, 1.0, (1, model.ndof)) * np.sin(
            2 * np.pi * (g.uniform(0.3, 1.5, (1, model.ndof)) * t + g.uniform(0, 1, (1, model.ndof))))
    return np.clip(q, lim[:, 0] + 0.05 * half, lim[:, 1] - 0.05 * half)


def capture_locals(fn, names, *args, **kwargs):
    """Run fn and grab named locals of its frame at return (used to read dp_search's `memo`/`costs`)."""
    got = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code is fn.__code__:
            for n in names:
                got[n] = frame.f_locals.get(n)

    sys.setprofile(prof)
    try:
        out = fn(*args, **kwargs)
    finally:
        sys.setprofile(None)
    return out, got


def main():
    install_stubs()
    torch.manual_seed(0)
    import cppflow  # noqa: F401  (the real reference package)
    from cppflow import search as rsearch, optimization as ropt, optimization_utils as rou
    from cppflow import collision_detection as rcd, evaluation_utils as rev
    from cppflow.data_types import Problem, Constraints
    from cppflow.lm_hyper_parameters import ALT_LOSS_V2_1_DIFF, ALT_LOSS_V2_1_POSE, OptimizationParameters

    assert torch.get_default_dtype() == torch.float32
    out = {}
    constraints = Constraints(0.01, 0.1, 7.0, 2.0)  # scripts/evaluate.py:51-56
    obstacles = {
        # problems/fetch__circle.yaml:13-16 and problems/panda__1cube.yaml
        "fetch": [(0.4, 0.4, 0.825, 0.3, 0.05, 0.8), (0.4, -0.4, 0.825, 0.3, 0.05, 0.8),
                  (0.4, 0.0, 1.225, 0.3, 0.85, 0.05), (0.4, 0.0, 0.425, 0.3, 0.85, 0.05)],
        "fetch_arm": [(0.4, 0.4, 0.825, 0.3, 0.05, 0.8), (0.4, 0.0, 0.425, 0.3, 0.85, 0.05)],
        "panda": [(0.0, 0.2, 0.7, 0.25, 0.25, 0.25)],
    }

    for rname, RC in (("fetch", Fetch), ("fetch_arm", FetchArm), ("panda", Panda)):
        robot = RC()
        model = robot._m
        D = model.ndof
        # ------------------------------------------------------------------ dp_search
        k, T = 14, 18
        g = np.random.default_rng(7)
        base = smooth_joint_path(model, T, seed=11)
        q = torch.tensor(base[None] + 0.35 * g.standard_normal((k, T, D)), dtype=torch.float32)
        q[:, :, 1] += torch.tensor(g.choice([0.0, 2 * np.pi], size=(k, T), p=[0.9, 0.1]), dtype=torch.float32)
        self_v = torch.tensor(g.random((k, T)) < 0.15)
        env_v = torch.tensor(g.random((k, T)) < 0.15)
        (best, got) = capture_locals(rsearch.dp_search, ("memo", "costs", "q_costs_external"), robot, q.clone(),
                                     self_v.clone(), env_v.clone(), verbosity=0)
        out[f"{rname}/dp/q"] = q.numpy()
        out[f"{rname}/dp/self_v"] = self_v.numpy()
        out[f"{rname}/dp/env_v"] = env_v.numpy()
        out[f"{rname}/dp/best_path"] = best.numpy()
        out[f"{rname}/dp/memo"] = got["memo"].numpy()
        out[f"{rname}/dp/costs"] = got["costs"].numpy()
        out[f"{rname}/dp/ext"] = got["q_costs_external"].numpy()
        out[f"{rname}/dp/mjacs"] = rsearch._get_mjacs(q.clone(), robot).numpy()
        out[f"{rname}/dp/jlim"] = rsearch.joint_limit_almost_violations_3d(robot, q).numpy()

        # ------------------------------------------------------------------ LM pieces
        T = 16
        qstar = torch.tensor(smooth_joint_path(model, T, seed=3), dtype=torch.float32)
        target = robot.forward_kinematics(qstar.double()).float()
        gen = torch.Generator().manual_seed(1234)
        x = qstar + 0.05 * torch.randn((T, D), generator=gen)
        x = rou.clamp_to_joint_limits(robot, x.clone())
        cuboids, Tcuboids = [], []
        for (ox, oy, oz, sx, sy, sz) in obstacles[rname]:
            cuboids.append(torch.tensor([-sx / 2, -sy / 2, -sz / 2, sx / 2, sy / 2, sz / 2]))
            Tc = torch.zeros((4, 4))
            Tc[:3, :3] = torch.eye(3)
            Tc[0, 3], Tc[1, 3], Tc[2, 3] = ox, oy, oz
            Tcuboids.append(Tc)
        problem = Problem(constraints, target, None, robot, "synthetic", f"{rname}__synthetic", [], Tcuboids, cuboids,
                          [])
        out[f"{rname}/lm/x"] = x.numpy()
        out[f"{rname}/lm/target"] = target.numpy()
        out[f"{rname}/lm/cuboids"] = np.stack([c.numpy() for c in cuboids]) if cuboids else np.zeros((0, 6), "f4")
        out[f"{rname}/lm/Tcuboids"] = np.stack([t.numpy() for t in Tcuboids]) if cuboids else np.zeros((0, 4, 4), "f4")

        e, cur = rou.get_6d_pose_errors(robot, x, target)
        out[f"{rname}/lm/pose_err"] = e.numpy()
        out[f"{rname}/lm/cur_pose"] = cur.numpy()

        opt_problem = ropt.OptimizationProblem(problem, constraints, x.clone(), target, 0, 1, None)
        opt_state = ropt.OptimizationState(x.clone(), 0, 0.0)
        xn, Jp, ep = ropt.levenberg_marquardt_only_pose(opt_problem, opt_state, ALT_LOSS_V2_1_POSE, return_residual=True)
        out[f"{rname}/lm/pose_step_x"] = xn.numpy()
        out[f"{rname}/lm/pose_step_J"] = Jp.numpy()
        out[f"{rname}/lm/pose_step_e"] = ep.numpy()

        # differencing step exactly as run_lm_alternating_loss does it (optimization.py:253-255); to make the
        # collision terms fire, use a perturbed path that dips into the obstacles / itself
        xd = x.clone()
        params_diff = OptimizationParameters(**ALT_LOSS_V2_1_DIFF.__dict__)
        params_diff.virtual_configs = xd.clone()
        opt_state = ropt.OptimizationState(xd.clone(), 0, 0.0)
        xn, jac, res = ropt.levenberg_marquardt_full(opt_problem, opt_state, params_diff, return_residual=True)
        out[f"{rname}/lm/diff_step_x"] = xn.numpy()
        out[f"{rname}/lm/diff_step_J"] = jac.get_J().numpy()
        out[f"{rname}/lm/diff_step_r"] = res.get_r().numpy()
        out[f"{rname}/lm/diff_n_self"] = np.array(0 if res.self_collisions is None else res.self_collisions.shape[0])
        out[f"{rname}/lm/diff_n_env"] = np.array(0 if res.env_collisions is None else res.env_collisions.shape[0])

        # a collision-heavy path: random configs (many capsule overlaps), virtual configs != x
        lim = torch.tensor(model.actuated_joints_limits)
        xc = lim[:, 0] + torch.rand((T, D), generator=gen) * (lim[:, 1] - lim[:, 0])
        xc = 0.5 * xc + 0.5 * x
        # plant self-colliding configurations (rare among random samples) in every third row
        cand = lim[:, 0] + torch.rand((4000, D), generator=gen) * (lim[:, 1] - lim[:, 0])
        bad = cand[OG.self_collision_distances(model, cand).min(dim=1).values < -0.01]
        assert bad.shape[0] >= T // 3 + 1, bad.shape
        xc[::3] = bad[: xc[::3].shape[0]]
        params_all = OptimizationParameters(**ALT_LOSS_V2_1_DIFF.__dict__)
        params_all.use_pose = True
        params_all.alpha_position = ALT_LOSS_V2_1_POSE.alpha_position
        params_all.alpha_rotation = ALT_LOSS_V2_1_POSE.alpha_rotation
        params_all.virtual_configs = x.clone()
        opt_state = ropt.OptimizationState(xc.clone(), 0, 0.0)
        jac, res = rou.LmResidualFns.get_r_and_J(params_all, robot, xc.clone(), target, Tcuboids=Tcuboids,
                                                 cuboids=cuboids)
        Jall, rall = jac.get_J(), res.get_r()
        out[f"{rname}/lm/all_x"] = xc.numpy()
        out[f"{rname}/lm/all_xv"] = x.numpy()
        out[f"{rname}/lm/all_J"] = Jall.numpy()
        out[f"{rname}/lm/all_r"] = rall.numpy()
        out[f"{rname}/lm/all_n_self"] = np.array(0 if res.self_collisions is None else res.self_collisions.shape[0])
        out[f"{rname}/lm/all_n_env"] = np.array(0 if res.env_collisions is None else res.env_collisions.shape[0])
        out[f"{rname}/lm/all_step_x"] = ropt._lm_full_step(Jall, rall, xc.clone(), params_all.lm_lambda).numpy()

        # ------------------------------------------------------------------ collision flags + validity metrics
        kq = 6
        qs = (lim[:, 0] + torch.rand((kq, T, D), generator=gen) * (lim[:, 1] - lim[:, 0]))
        qs = 0.6 * qs + 0.4 * x[None]
        qs[:, ::4] = bad[: kq * qs[:, ::4].shape[1]].reshape(kq, -1, D)
        out[f"{rname}/cd/q"] = qs.numpy()
        out[f"{rname}/cd/self"] = rcd.qpaths_batched_self_collisions(problem, qs).numpy()
        out[f"{rname}/cd/env"] = rcd.qpaths_batched_env_collisions(problem, qs).numpy()
        ecm, edeg = rev.calculate_pose_error_cm_deg(robot, x, target)
        out[f"{rname}/ev/err_cm"] = ecm.numpy()
        out[f"{rname}/ev/err_deg"] = edeg.numpy()
        out[f"{rname}/ev/angular_changes"] = rev.angular_changes(x).numpy()
        xc2 = xc.clone()
        xc2[0] += 10.0
        out[f"{rname}/lm/clamp_in"] = xc2.numpy()
        out[f"{rname}/lm/clamp_out"] = rou.clamp_to_joint_limits(robot, xc2.clone()).numpy()

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), len(out), "arrays")
    for k_ in sorted(out):
        if k_.endswith(("n_self", "n_env")):
            print(k_, out[k_])


if __name__ == "__main__":
    main()
