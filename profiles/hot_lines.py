"""Per-source-line instruction counts and stall samples of one kernel of an `ncu --set full --import-source on` report.

    python profiles/hot_lines.py gpurun_out/prof_lm.ncu-rep lm_assemble cppflow_b200/libcppflow_b200.so [top_n]

The ncu CLI prints metrics only on the SASS view, so the SASS rows (in address order) are matched against
`nvdisasm -g` of the kernel's cubin, whose `//## File "...", line N` annotations give the source line (the innermost
inlined location) of each instruction.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    launches, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            launches.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    return launches[0]


def line_table(lib, kernel, n_rows=None):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel not in txt:
            continue
        # split into functions
        lines = txt.splitlines()
        funcs, out, cur_loc = [], None, ("?", 0)
        for ln in lines:
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                out = None
                if kernel in m.group(1):
                    out = []
                    funcs.append(out)
                continue
            if out is None:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur_loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((int(m.group(1), 16), m.group(2), cur_loc))
        # several template instantiations may match: take the one with as many instructions as ncu reports
        for f_ in funcs:
            if n_rows is None or len(f_) == n_rows:
                return f_
        if funcs:
            return funcs[0]
    raise SystemExit("kernel not found in any cubin")


def main():
    rep, kernel, lib = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    L = sass_rows(rep, kernel)
    hdr = L["hdr"]
    i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    rows = L["rows"]
    tbl = line_table(lib, kernel, len(rows))
    if len(tbl) != len(rows):
        print(f"warning: {len(rows)} ncu rows vs {len(tbl)} disassembled instructions; matching by order", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for k, r in enumerate(rows[: len(tbl)]):
        loc = tbl[k][2]
        n, s = int(r[i_inst] or 0), int(r[i_samp] or 0)
        a = agg[loc]
        a[0] += n
        a[1] += s
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                a[2][h[6:]] += v
        tot_i += n
        tot_s += s
    print(f"kernel {L['name'][:100]}\n total warp-instructions {tot_i}, samples {tot_s}")
    print(f"{'file:line':<28}{'inst %':>8}{'samples %':>10}  top stalls")
    for loc, (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        tops = ", ".join(f"{k}:{v}" for k, v in st.most_common(3))
        print(f"{loc[0] + ':' + str(loc[1]):<28}{100 * n / tot_i:8.2f}{100 * s / max(tot_s, 1):10.2f}  {tops}")
    by_file = collections.Counter()
    for loc, (n, s, st) in agg.items():
        by_file[loc[0]] += n
    print("by file (inst %):", {k: round(100 * v / tot_i, 1) for k, v in by_file.most_common()})


if __name__ == "__main__":
    main()
