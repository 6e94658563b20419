"""Where a kernel's instructions and stall samples go, from the source page of an .ncu-rep (captured with
`--set full --import-source on`; kernels built with -lineinfo).

    python tools/ncu_source_hot.py gpurun_out/prof_step.ncu-rep [kernel-name-regex] [--lines N]

Prints, per kernel in the report: warp instructions and stall samples by opcode, and the N hottest SASS lines.
"""
import collections
import csv
import re
import subprocess
import sys


def tables(rep, regex=None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if regex:
        cmd += ["--kernel-name", "regex:" + regex]
    rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
    out, i = [], 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "Kernel Name" and i + 1 < len(rows):
            out.append((r[1], rows[i + 1], []))
            i += 2
            continue
        if out:
            out[-1][2].append(r)
        i += 1
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_lines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 12
    if "--lines" in sys.argv:
        args = [a for a in args if a != sys.argv[sys.argv.index("--lines") + 1]]
    rep, regex = args[0], (args[1] if len(args) > 1 else None)
    for name, hdr, data in tables(rep, regex):
        si, st, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        d = [(r[si].strip(), int(r[st] or 0), int(r[ie] or 0)) for r in data if len(r) > max(si, st, ie)]
        tot_s, tot_i = max(1, sum(x[1] for x in d)), max(1, sum(x[2] for x in d))
        print(f"===== {name[:100]}\n      SASS lines {len(d)}, warp instructions {tot_i}, stall samples {tot_s}")
        by_i, by_s = collections.Counter(), collections.Counter()
        for s, sm, n in d:
            op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0] if s else "?"
            op = ".".join(op.split(".")[:2])
            by_i[op] += n
            by_s[op] += sm
        print("  by opcode (sorted by stall samples)")
        for op, c in by_s.most_common(18):
            print(f"    {op:24s} samples {c:7d} {100 * c / tot_s:5.1f} %   instructions {by_i[op]:10d} {100 * by_i[op] / tot_i:5.1f} %")
        print(f"  hottest {n_lines} SASS lines")
        for s, sm, n in sorted(d, key=lambda x: -x[1])[:n_lines]:
            print(f"    {100 * sm / tot_s:5.1f} %  x{n:<9d} {s[:90]}")


if __name__ == "__main__":
    main()
