"""Golden vectors of the reference's row scaling / filtering helpers of get_r_and_J (run once, in the build container).

    python tests/golden/make_golden_rowscale.py      # writes tests/golden/reference_rowscale_golden.npz

Calls the unmodified (stub-imported, see make_golden.py)
    LmResidualFns._scale_down_rows_from_r_J_pose_below_error          optimization_utils.py:288-329
    LmResidualFns._scale_down_rows_from_r_J_differencing_below_error  optimization_utils.py:352-397
    filter_rows_from_r_J_differencing                                  optimization_utils.py:736-768
on (a) the input of the reference's own unit test (tests/optimization_utils_test.py:126-155, Fetch: the joint-0 rows
are prismatic) and (b) seeded random residual / Jacobian blocks for Fetch and Panda whose entries straddle the
thresholds, each with and without the shift-to-threshold option.  These options are OFF in both live parameter sets
(lm_hyper_parameters.py:86-151), and get_r_and_J cannot even reach them today: it reads `pms.constraints`
(optimization_utils.py:515-520, :562-567), a field OptimizationParameters does not have (lm_hyper_parameters.py:14-63).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    MG.install_stubs()
    import cppflow  # noqa: F401
    from cppflow import optimization_utils as rou

    out = {}
    # (a) the reference's own test input, Fetch
    r_test = torch.tensor([0.5, 0.1, 1.6, 0.1, 0.1, 0.1, 0.1, 0.1, -0.4, 1.7, -1.7, 0.1, 0.1, 0.1, 0.1, 0.1,
                           0.2, 0.01, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1])[:, None]
    cases = {"fetch/unit": (MG.Fetch(), r_test, torch.ones((24, 32)))}
    # (b) seeded random blocks: n waypoints -> (n - 1) * ndof differencing rows, n * ndof columns
    g = torch.Generator().manual_seed(77)
    for rname, RC, n in (("fetch", MG.Fetch, 7), ("panda", MG.Panda, 6)):
        robot = RC()
        D = robot.ndof
        r = (torch.rand(((n - 1) * D, 1), generator=g) - 0.5) * 4.0
        r[::5] *= 0.05  # plenty of rows below the thresholds
        J = torch.randn(((n - 1) * D, n * D), generator=g)
        cases[f"{rname}/random"] = (robot, r, J)
    for key, (robot, r, J) in cases.items():
        out[f"{key}/robot"] = np.array(robot.name)
        out[f"{key}/diff/r_in"], out[f"{key}/diff/J_in"] = r.numpy(), J.numpy()
        for shift in (False, True):
            Jo, ro, inv = rou.LmResidualFns._scale_down_rows_from_r_J_differencing_below_error(
                robot, r.clone(), J.clone(), mjac_threshold_m=0.25, mjac_threshold_rad=1.5, scale=0.5,
                shift_invalid_to_threshold=shift)
            out[f"{key}/diff/scale_shift{int(shift)}/r"], out[f"{key}/diff/scale_shift{int(shift)}/J"] = ro.numpy(), Jo.numpy()
            out[f"{key}/diff/scale_shift{int(shift)}/invalid"] = inv.numpy()
            rf, Jf = rou.filter_rows_from_r_J_differencing(robot, r.clone(), J.clone(), threshold_rad=1.5, threshold_m=0.25,
                                                           shift_to_threshold=shift)
            out[f"{key}/diff/filter_shift{int(shift)}/r"], out[f"{key}/diff/filter_shift{int(shift)}/J"] = rf.numpy(), Jf.numpy()
    # pose rows: 6 per waypoint, [rot x3, pos x3]
    for name, n in (("pose/a", 5), ("pose/b", 9)):
        r = (torch.rand((6 * n, 1), generator=g) - 0.5) * 0.2
        r[::4] *= 0.01
        J = torch.randn((6 * n, 8 * n), generator=g)
        out[f"{name}/r_in"], out[f"{name}/J_in"] = r.numpy(), J.numpy()
        ro, Jo, inv = rou.LmResidualFns._scale_down_rows_from_r_J_pose_below_error(
            r.clone(), J.clone(), error_threshold_m=0.01, error_threshold_rad=0.03, scale=0.25)
        out[f"{name}/r"], out[f"{name}/J"], out[f"{name}/invalid"] = ro.numpy(), Jo.numpy(), inv.numpy()
    np.savez_compressed(os.path.join(HERE, "reference_rowscale_golden.npz"), **out)
    print("wrote reference_rowscale_golden.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
