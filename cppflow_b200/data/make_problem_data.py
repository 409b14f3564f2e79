"""Pack the reference's benchmark problem definitions and Cartesian target paths into data files the GPU box can
read (it has no /root/reference).  Run once in the build container:

    python cppflow_b200/data/make_problem_data.py

Inputs (read-only): /root/reference/cppflow/problems/*.yaml, /root/reference/tests/*.yaml,
                    /root/reference/cppflow/paths/*.csv   (format time,x,y,z,qw,qx,qy,qz - paths/circle.csv:1)
Outputs: cppflow_b200/data/problems.json   (robot, path name, offsets, obstacles per problem - verbatim values)
         cppflow_b200/data/target_paths.npz (one float64 [T,7] array per path, the csv columns 1..7)
"""
import csv
import glob
import json
import os

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    problems = {}
    files = sorted(glob.glob(os.path.join(REF, "cppflow/problems/*.yaml"))) + sorted(
        glob.glob(os.path.join(REF, "tests/*.yaml")))
    for f in files:
        with open(f) as fh:
            d = yaml.load(fh, Loader=yaml.FullLoader)
        name = os.path.splitext(os.path.basename(f))[0]
        obstacles = []
        for obs in d.get("obstacles", []) or []:
            parsed = {}
            for item in obs:
                parsed.update(item)
            obstacles.append({k: float(v) for k, v in parsed.items()})
        problems[name] = {
            "robot": d["robot"],
            "path_name": d["path_name"],
            "path_offset_frame": d["path_offset_frame"],
            "path_xyz_offset": [float(v) for v in d["path_xyz_offset"]],
            "path_R_offset": [[float(v) for v in row] for row in d["path_R_offset"]],
            "obstacle_xyz_offset": [float(v) for v in d.get("obstacle_xyz_offset", [0, 0, 0])],
            "obstacles": obstacles,
            "source": os.path.relpath(f, REF),
        }
    paths = {}
    for f in sorted(glob.glob(os.path.join(REF, "cppflow/paths/*.csv"))):
        with open(f) as fh:
            rows = [[float(x) for x in row] for i, row in enumerate(csv.reader(fh)) if i > 0]
        paths[os.path.splitext(os.path.basename(f))[0]] = np.array(rows, dtype=np.float64)[:, 1:]
    with open(os.path.join(HERE, "problems.json"), "w") as fh:
        json.dump(problems, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "target_paths.npz"), **paths)
    print(len(problems), "problems;", {k: v.shape for k, v in paths.items()})


if __name__ == "__main__":
    main()
