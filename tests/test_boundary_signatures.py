"""The drop-in boundary: every hot-path name keeps the reference's signature.

The real `/root/reference/cppflow` is imported in a subprocess with jrl / klampt / ikflow / matplotlib stubbed (the
stubs of tests/golden/make_golden.py) and `inspect.signature` of each boundary function is compared with the repo's:
the reference's parameters must come first, in the same order, with the same names and defaults; anything the repo adds
(e.g. `mesh_validator=`, `native=`) must be optional.  Dataclasses of the boundary must have the reference's fields in
the reference's order.  Skipped where /root/reference does not exist (the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"

FUNCTIONS = [  # (module, qualified name)
    ("search", "dp_search"), ("search", "joint_limit_almost_violations_3d"),
    ("collision_detection", "qpaths_batched_env_collisions"), ("collision_detection", "qpaths_batched_self_collisions"),
    ("collision_detection", "get_only_non_colliding_qpaths"),
    ("optimization", "run_lm_optimization"), ("optimization", "run_lm_alternating_loss"),
    ("optimization", "levenberg_marquardt_only_pose"), ("optimization", "levenberg_marquardt_full"),
    ("optimization_utils", "LmResidualFns.get_r_and_J"), ("optimization_utils", "get_6d_pose_errors"),
    ("optimization_utils", "clamp_to_joint_limits"), ("optimization_utils", "x_is_valid"),
    ("evaluation_utils", "angular_changes"), ("evaluation_utils", "errors_are_below_threshold"),
    ("data_type_utils", "problem_from_filename"),
]
DATACLASSES = [
    ("optimization", "OptimizationProblem"), ("optimization", "OptimizationState"), ("optimization", "OptimizationResult"),
    ("lm_hyper_parameters", "OptimizationParameters"), ("data_types", "Constraints"), ("data_types", "Problem"),
    ("data_types", "PlannerSettings"), ("data_types", "TimingData"),
]

_SCRIPT = r'''
import dataclasses, importlib, inspect, json, sys
sys.path.insert(0, %(repo)r); sys.path.insert(0, %(repo)r + "/tests/golden")
import make_golden as MG
MG.install_stubs()
import cppflow  # the real reference package

def resolve(mod, qual):
    obj = importlib.import_module(mod)
    for part in qual.split("."):
        obj = getattr(obj, part)
    return obj

def params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        d = None if p.default is inspect._empty else repr(p.default)
        out.append([p.name, str(p.kind), p.default is not inspect._empty, d])
    return out

res = {"functions": {}, "dataclasses": {}}
for mod, qual in %(functions)r:
    res["functions"][mod + "." + qual] = [params(resolve("cppflow." + mod, qual)), params(resolve("cppflow_b200." + mod, qual))]
for mod, qual in %(dataclasses)r:
    ref, ours = resolve("cppflow." + mod, qual), resolve("cppflow_b200." + mod, qual)
    res["dataclasses"][mod + "." + qual] = [[f.name for f in dataclasses.fields(ref)], [f.name for f in dataclasses.fields(ours)]]
print("RESULT" + json.dumps(res))
'''


@pytest.fixture(scope="module")
def signatures():
    if not os.path.isdir(os.path.join(REFERENCE, "cppflow")):
        pytest.skip("/root/reference is not available here")
    code = _SCRIPT % dict(repo=REPO, functions=FUNCTIONS, dataclasses=DATACLASSES)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    return json.loads(line[len("RESULT"):])


def test_function_signatures_keep_the_reference_prefix(signatures):
    problems = []
    for name, (ref, ours) in signatures["functions"].items():
        if len(ours) < len(ref):
            problems.append(f"{name}: {len(ours)} parameters, the reference has {len(ref)}")
            continue
        for i, (r, o) in enumerate(zip(ref, ours)):
            if r[0] != o[0]:
                problems.append(f"{name}: parameter {i} is '{o[0]}', the reference calls it '{r[0]}'")
            elif r[2] != o[2] and r[2]:
                problems.append(f"{name}: '{r[0]}' has a default in the reference, none here")
            elif r[2] and o[2] and r[3] != o[3]:
                problems.append(f"{name}: default of '{r[0]}' is {o[3]}, the reference has {r[3]}")
        for o in ours[len(ref):]:
            if not o[2] and "VAR_" not in o[1]:
                problems.append(f"{name}: extra parameter '{o[0]}' has no default")
    assert not problems, "\n".join(problems)


def test_dataclass_fields_keep_the_reference_order(signatures):
    problems = []
    for name, (ref, ours) in signatures["dataclasses"].items():
        if ours[: len(ref)] != ref:
            problems.append(f"{name}: fields {ours}, the reference has {ref}")
    assert not problems, "\n".join(problems)
