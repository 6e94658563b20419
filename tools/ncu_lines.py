"""Instructions and stall samples per CUDA source line of one kernel in an .ncu-rep captured with --import-source on.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel-regex> [N]
"""
import collections
import csv
import subprocess
import sys


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + regex], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = fname = None
    agg = collections.OrderedDict()
    seen = set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            if fname in seen:              # the report lists a kernel once per captured launch: keep the first
                break
            seen.add(fname)
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0] != "":
            cur = (fname, r[0])
            agg.setdefault(cur, [0, 0, " ".join(r[1:4])[:100]])
            continue
        try:
            st, ie = int(r[4] or 0), int(r[7] or 0)
        except (ValueError, IndexError):
            continue
        if cur:
            agg[cur][0] += st
            agg[cur][1] += ie
    tot_i = max(1, sum(v[1] for v in agg.values()))
    tot_s = max(1, sum(v[0] for v in agg.values()))
    print(f"warp instructions {tot_i}, stall samples {tot_s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
        print(f"{k[0]:22s}:{k[1]:>5s} inst {100 * v[1] / tot_i:5.1f}%  stall {100 * v[0] / tot_s:5.1f}%  {v[2]}")


if __name__ == "__main__":
    main()
