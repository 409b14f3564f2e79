"""Per-kernel SASS opcode histogram of the built library (no GPU needed: cuobjdump + nvdisasm on the in-tree .so).

    python profiles/sass_histogram.py [cppflow_b200/libcppflow_b200.so] > profiles/r02_sass_histogram.md

One row per kernel (Fetch instantiations; the other robots differ only in the unrolled chain): instruction count,
registers, and the opcodes that tell how the kernel maps to sm_100a - UBLKCP (TMA bulk copy), SYNCS (mbarrier),
UCGABAR / cluster barriers, REDUX (warp reduce), FFMA2 (packed FP32), LDGSTS (cp.async), MUFU, and the FP32 mix.
No UTCMMA / LDTM / UTMALDG is expected: the path has no contraction wider than 8."""
import collections
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "cppflow_b200", "libcppflow_b200.so")
SHOW = ["UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "REDUX", "SHFL", "FFMA2", "FFMA", "FADD", "FMUL", "FMNMX", "FMNMX3",
        "FSETP", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "ATOMS", "ATOMG", "MEMBAR", "BAR", "STL", "LDL", "UTCMMA",
        "LDTM", "UTMALDG"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    names = []
    for o in out:
        o = o.replace("void ", "").replace("cppflow::", "").replace("(bool)0", "false").replace("(bool)1", "true")
        o = o.replace("(int)", "")
        names.append(o.split(">(")[0] + ">" if ">(" in o else o.split("(")[0])
    return names


def main():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(LIB)], cwd=tmp, capture_output=True)
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", os.path.abspath(LIB)], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", ln)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    rows = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, ops = None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur = m.group(1)
                ops = collections.Counter()
                rows.append((cur, ops))
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if m and ops is not None:
                ops[m.group(1)] += 1
    rows = [(n, o) for n, o in rows if sum(o.values()) > 0 and ("FetchTILb0" in n or "cppflow" in n and "Fetch" not in n and "Panda" not in n)]
    names = demangle([n for n, _ in rows])
    print("# SASS opcode histogram per kernel (sm_100a cubins of libcppflow_b200.so; `python profiles/sass_histogram.py`)\n")
    print("Fetch instantiations (`FetchT<false>`) and the robot-independent kernels; counts are static instructions.\n")
    used = [k for k in SHOW if any(o.get(k, 0) for _, o in rows)]
    print("| kernel | instr | regs | stack B | " + " | ".join(used) + " |")
    print("|---|---|---|---|" + "---|" * len(used))
    tot = collections.Counter()
    for (mangled, ops), name in sorted(zip(rows, names), key=lambda t: t[1]):
        r = regs.get(mangled, ("?", "?", "?"))
        print(f"| `{name[:88]}` | {sum(ops.values())} | {r[0]} | {r[1]} | " + " | ".join(str(ops.get(k, 0)) for k in used) + " |")
        tot.update(ops)
    print("\nWhole library (all robots): " + ", ".join(f"{k} {v}" for k, v in tot.most_common(12)))
    absent = [k for k in ("UTCMMA", "LDTM", "UTMALDG") if not tot.get(k)]
    print("\nAbsent, as expected for a path without GEMM-shaped work: " + ", ".join(absent))


if __name__ == "__main__":
    main()
