"""Golden vectors of the reference's OWN alternating LM loop (run once, in the build container).

    python tests/golden/make_golden_loop.py          # writes tests/golden/reference_loop_golden.npz

Imports the real `/root/reference/cppflow` with the stubs of make_golden.py and calls the unmodified
    cppflow.optimization.run_lm_optimization -> run_lm_alternating_loss   (optimization.py:147-426)
in both of the planner's call patterns (planners.py:402-422: normal = max 20 steps, return once valid; anytime = max 75
steps, run until the trajectory length converges) on seeded synthetic problems of the three robots.  The klampt mesh
checks the loop reaches through `x_is_valid` (`robot.config_self_collides`, `robot.config_collides_with_env`,
collision_detection.py:89-120) are backed by the capsule oracle - exactly the check the CUDA loop performs
(csrc/lm_loop.cu), so the two loops decide on the same quantities.

Each case is run twice: in the reference's dtype (float32) and with torch's default dtype switched to float64 (the
reference's code is dtype-generic; float64 is the exact-arithmetic result of the same algorithm).  Stored per case:
inputs (x_seed, target path, cuboids), and for each dtype the step sequence ('p' = levenberg_marquardt_only_pose,
'd' = levenberg_marquardt_full; recorded by wrapping the two functions, which the loop looks up in its module),
n_steps_taken, is_valid and x_opt.  Nothing under /root/reference is read at test time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import make_golden as MG  # noqa: E402
from oracle import geometry as OG  # noqa: E402

OBSTACLES = {
    "fetch": [(0.4, 0.4, 0.825, 0.3, 0.05, 0.8), (0.4, -0.4, 0.825, 0.3, 0.05, 0.8),
              (0.4, 0.0, 1.225, 0.3, 0.85, 0.05), (0.4, 0.0, 0.425, 0.3, 0.85, 0.05)],
    "fetch_arm": [(0.4, 0.4, 0.825, 0.3, 0.05, 0.8), (0.4, 0.0, 0.425, 0.3, 0.85, 0.05)],
    "panda": [(0.0, 0.2, 0.7, 0.25, 0.25, 0.25)],
}

# (case name, robot, T, seed and amplitude of the smooth joint path, kind of seed perturbation).  With amplitude 0.08 the
# path itself is collision-free and inside the mjac thresholds (a valid plan exists); the amplitude 0.25 path of the
# last case runs through a cuboid, so the loop never finds a valid iterate.
CASES = [
    ("fetch_smooth", "fetch", 40, 3, 0.08, "smooth"),
    ("fetch_noisy", "fetch", 36, 20, 0.08, "noisy"),
    ("fetch_arm_smooth", "fetch_arm", 33, 17, 0.08, "smooth"),
    ("panda_smooth", "panda", 40, 12, 0.08, "smooth"),
    ("panda_noisy", "panda", 30, 34, 0.08, "noisy"),
    ("fetch_colliding", "fetch", 36, 5, 0.25, "noisy"),
]
PATTERNS = {
    # planners.py:402-422
    "normal": dict(max_n_steps=20, return_if_valid_after_n_steps=0, convergence_threshold=1e6),
    "anytime": dict(max_n_steps=75, return_if_valid_after_n_steps=int(1e8), convergence_threshold=0.005),
}


def add_klampt_backing(robot_cls):
    """config_self_collides / config_collides_with_env (klampt in jrl) answered by the capsule oracle."""

    def config_self_collides(self, x):
        q = torch.as_tensor(np.asarray(x), dtype=torch.float64)[None]
        return bool(OG.self_collision_distances(self._m, q).min() < 0)

    def config_collides_with_env(self, x, j):
        q = torch.as_tensor(np.asarray(x), dtype=torch.float64)[None]
        cuboids, Tcuboids = self._env
        return bool(OG.env_collision_distances(self._m, q, cuboids[j].double(), Tcuboids[j].double()).min() < 0)

    robot_cls.config_self_collides = config_self_collides
    robot_cls.config_collides_with_env = config_collides_with_env


def seed_path(model, qstar, kind, gen):
    T, D = qstar.shape
    lim = torch.tensor(model.actuated_joints_limits, dtype=torch.float64)
    if kind == "smooth":
        # low-frequency offset: pose errors of a few cm, joint jumps already inside the mjac thresholds
        t = torch.linspace(0, 1, T, dtype=torch.float64)[:, None]
        amp = 0.06 * torch.rand((1, D), generator=gen, dtype=torch.float64)
        ph = torch.rand((1, D), generator=gen, dtype=torch.float64)
        x = qstar + amp * torch.sin(2 * np.pi * (t + ph))
    else:
        x = qstar + 0.02 * torch.randn((T, D), generator=gen, dtype=torch.float64)
    return torch.minimum(torch.maximum(x, lim[:, 0]), lim[:, 1])


def main():
    MG.install_stubs()
    for cls in (MG.Fetch, MG.FetchArm, MG.Panda):
        add_klampt_backing(cls)
    import cppflow  # noqa: F401
    from cppflow import optimization as ropt
    from cppflow.data_types import Problem, Constraints

    sched = []
    real_pose, real_full = ropt.levenberg_marquardt_only_pose, ropt.levenberg_marquardt_full

    def rec_pose(*a, **k):
        sched.append("p")
        return real_pose(*a, **k)

    def rec_full(*a, **k):
        sched.append("d")
        return real_full(*a, **k)

    ropt.levenberg_marquardt_only_pose, ropt.levenberg_marquardt_full = rec_pose, rec_full

    out = {}
    constraints = Constraints(0.01, 0.1, 7.0, 2.0)  # scripts/evaluate.py:51-56
    robots = {"fetch": MG.Fetch, "fetch_arm": MG.FetchArm, "panda": MG.Panda}
    for name, rname, T, pseed, amp, kind in CASES:
        robot = robots[rname]()
        model = robot._m
        qstar = torch.tensor(MG.smooth_joint_path(model, T, seed=pseed, amp=amp), dtype=torch.float64)
        gen = torch.Generator().manual_seed(4321 + pseed)
        x_seed64 = seed_path(model, qstar, kind, gen)
        cuboids, Tcuboids = [], []
        for (ox, oy, oz, sx, sy, sz) in OBSTACLES[rname]:
            cuboids.append(torch.tensor([-sx / 2, -sy / 2, -sz / 2, sx / 2, sy / 2, sz / 2], dtype=torch.float32))
            Tc = torch.zeros((4, 4), dtype=torch.float32)
            Tc[:3, :3] = torch.eye(3)
            Tc[0, 3], Tc[1, 3], Tc[2, 3] = ox, oy, oz
            Tcuboids.append(Tc)
        robot._env = (cuboids, Tcuboids)
        target64 = robot.forward_kinematics(qstar)
        out[f"{name}/robot"] = np.array(rname)
        out[f"{name}/x_seed"] = x_seed64.float().numpy()
        out[f"{name}/target"] = target64.float().numpy()
        out[f"{name}/cuboids"] = np.stack([c.numpy() for c in cuboids])
        out[f"{name}/Tcuboids"] = np.stack([t.numpy() for t in Tcuboids])
        for dt_name, dt in (("f32", torch.float32), ("f64", torch.float64)):
            torch.set_default_dtype(dt)
            try:
                # both dtypes start from the SAME float32-representable inputs
                x_seed = x_seed64.float().to(dt)
                target = target64.float().to(dt)
                problem = Problem(constraints, target, None, robot, "synthetic", f"{rname}__synthetic", [],
                                  [t.to(dt) for t in Tcuboids], [c.to(dt) for c in cuboids], [None] * len(cuboids))
                for pname, kw in PATTERNS.items():
                    del sched[:]
                    res = ropt.run_lm_optimization(problem, x_seed.clone(), tmax_sec=1e9, verbosity=0, **kw)
                    key = f"{name}/{pname}/{dt_name}"
                    out[f"{key}/schedule"] = np.array("".join(sched))
                    out[f"{key}/n_steps_taken"] = np.array(res.n_steps_taken)
                    out[f"{key}/is_valid"] = np.array(bool(res.is_valid))
                    out[f"{key}/x_opt"] = res.x_opt.detach().numpy()
                    print(key, "".join(sched), res.n_steps_taken, res.is_valid)
            finally:
                torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, "reference_loop_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_loop_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
