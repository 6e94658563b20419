"""Summarise an `ncu --set full` capture of the tensor-core convolution launches into profiles/conv_traffic.json, the file
bench.py's roofline.traffic reads: dram__bytes_read.sum + dram__bytes_write.sum per launch, per layer class, together
with the digest of the kernel sources the capture was taken of (bench.py drops the figure when the library was built from
other sources).

    python tools/ncu_traffic.py bf16=gpurun_out/r2_prof_conv_cg2.ncu-rep [bf16x3=...] --source "profiles/r02_ncu_conv.txt" [--out file]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpc_b200  # noqa: E402,F401
from sdpc_b200 import build as b  # noqa: E402


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size")
           if k in hdr}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = []
    for r in rows[2:]:
        if "conv_umma" not in r[col["Kernel Name"]]:
            continue
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
        res.append(dict(kernel=r[col["Kernel Name"]][:120], dram_read=rd, dram_write=wr, dram_bytes=rd + wr,
                        us=float(r[col["gpu__time_duration.sum"]]),
                        tensor_pipe_pct=float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
                        grid=int(float(r[col["launch__grid_size"]]))))
    return res


def main():
    args = [a for a in sys.argv[1:] if "=" in a and not a.startswith("--")]
    source = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else "ncu --set full capture"
    rec = dict(kernel_source_digest=b.kernel_digest(), source=source, arms={})
    for a in args:
        arm, rep = a.split("=", 1)
        ls = launches(rep)
        rec["arms"][arm] = dict(launches=ls, n=len(ls),
                                dram_bytes_per_launch_mean=sum(l["dram_bytes"] for l in ls) / max(1, len(ls)),
                                tensor_pipe_pct_mean=sum(l["tensor_pipe_pct"] for l in ls) / max(1, len(ls)))
        print(arm, len(ls), "launches, mean dram bytes", rec["arms"][arm]["dram_bytes_per_launch_mean"],
              "tensor pipe %", rec["arms"][arm]["tensor_pipe_pct_mean"])
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(ROOT, "profiles", "conv_traffic.json")
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
