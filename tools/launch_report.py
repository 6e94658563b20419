"""Summarise an ncu launch list (gpu__time_duration csv): per-kernel totals and one forward in order."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void sdpc::", "").replace("sdpc::", "") for r in rows]
    t = [float(r["Metric Value"].replace(",", "")) / 1e3 for r in rows]
    return names, t


def main():
    names, t = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in zip(names, t):
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v for _, v in agg.values())
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f}%  n={c:4d}  {k[:90]}")
    print(f"total {tot / 1e3:.3f} ms over {len(names)} launches")
    if len(sys.argv) > 2:
        starts = [i for i, n in enumerate(names) if n.startswith("begin_conv")]
        s, e = starts[0], starts[1]
        for i in range(s, e):
            print(f"{i - s:3d} {names[i][:44]:44s} {t[i]:8.1f} us")
        print("one forward+step:", sum(t[s:e]) / 1e3, "ms")


if __name__ == "__main__":
    main()
