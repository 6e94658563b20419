"""Build libsdpc_b200.so (C ABI, include/sdpc_b200.h) with nvcc for sm_100a, in-tree.

No torch dependency in the library: plain CUDA runtime + driver entry point for TMA descriptors.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsdpc_b200.so")
OBJ = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
if os.environ.get("SDPC_DEV_HOOKS"):            # timing probes that corrupt results (tools/gpu_epi_probe.sh); never in a shipped build
    COMMON.append("-DSDPC_DEV_HOOKS")
# (source, extra flags)
UNITS = [
    ("common.cu", []),
    # reference evaluates every op with its own rounding: no FMA contraction (crossview_core.h)
    ("crossview.cu", ["-fmad=false"]),
    ("lidar_projection.cu", ["-fmad=false"]),
    ("output_stage.cu", ["-fmad=false"]),
    ("dataset_assembly.cu", ["-fmad=false"]),
    ("conv_umma.cu", []),
    ("scorenet.cu", []),
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "sdpc_b200.h"))
    return deps


def kernel_digest():
    """sha256 over the convolution kernel sources: profiles/conv_traffic.json records it, so bench.py can tell whether
    the committed ncu capture is of the kernels this library was built from."""
    h = hashlib.sha256()                     # file names + contents, not paths: the capture is taken on another box
    for f in ("conv_umma.cu", "conv_umma.h", "sm100_prims.cuh", "score_types.cuh"):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest(sources())
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    nvcc = _nvcc()
    objs = []
    for src, extra in UNITS:
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + COMMON + extra + ["-c", os.path.join(CSRC, src), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        objs.append(o)
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static"])
    with open(stamp, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
