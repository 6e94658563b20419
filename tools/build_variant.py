"""Build a VARIANT of libsdpc_b200.so from the tree plus one or more patches, without touching the shipped library:

    python tools/build_variant.py gpurun_ab/lib7.so tools/patches/cg2_7stages.patch [more.patch ...] [-D NAME[=VALUE] ...]

The package's csrc/ and include/ are copied to a scratch directory, the patches are applied there (`git apply`), every
unit is compiled with the flags of sdpc_b200/build.py and linked into the given path.  `gpurun_ab/` is git-ignored but
travels to the GPU box, so a variant built here can be timed against the shipped build on the same box with
`SDPC_LIB=$PWD/gpurun_ab/lib7.so python tools/quick_time.py 8 bf16` (tools/gpu_ab_lib.sh alternates the two) and checked
bit for bit with tools/ab_probe.py.  Prints the kernels' register / spill / shared-memory lines of the patched units."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpc_b200  # noqa: E402,F401
from sdpc_b200 import build as b  # noqa: E402


def main():
    argv = sys.argv[1:]
    defines = []
    while "-D" in argv:
        i = argv.index("-D")
        defines.append("-D" + argv[i + 1])
        del argv[i:i + 2]
    if not argv:
        raise SystemExit(__doc__)
    out, patches = os.path.abspath(argv[0]), [os.path.abspath(p) for p in argv[1:]]
    work = tempfile.mkdtemp(prefix="sdpc_variant_")
    pkg = os.path.basename(b.HERE)
    shutil.copytree(b.CSRC, os.path.join(work, pkg, "csrc"))
    shutil.copytree(os.path.join(ROOT, "include"), os.path.join(work, "include"))
    for p in patches:
        subprocess.check_call(["git", "apply", "--include", pkg + "/csrc/*", "--include", "include/*", p], cwd=work)
    changed = set()
    for dirpath, _, files in os.walk(work):
        for f in files:
            rel = os.path.relpath(os.path.join(dirpath, f), work)
            if open(os.path.join(work, rel), "rb").read() != open(os.path.join(ROOT, rel), "rb").read():
                changed.add(f)
    print("patched files:", sorted(changed) or "none")
    nvcc, objs = b._nvcc(), []
    for src, extra in b.UNITS:
        o = os.path.join(work, src.replace(".cu", ".o"))
        verbose = ["-Xptxas=-v"] if (src in changed or any(c.endswith((".h", ".cuh")) for c in changed)) else []
        cmd = [nvcc] + b.ARCH + b.COMMON + extra + defines + verbose + ["-c", os.path.join(work, pkg, "csrc", src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stderr)
            raise SystemExit(f"{src}: nvcc failed")
        if verbose:
            lines = r.stderr.splitlines()
            for i, l in enumerate(lines):
                if "Compiling entry function" in l and ("conv_umma" in l or src != "conv_umma.cu"):
                    name = l.split("'")[1]
                    used = next((x for x in lines[i + 1:i + 4] if "Used" in x), "")
                    spill = next((x for x in lines[i + 1:i + 4] if "spill" in x), "")
                    print(f"  {name[:70]:70s} {used.split('ptxas info    :')[-1].strip()} |{spill.split(':')[-1].strip()}")
        objs.append(o)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call([nvcc] + b.ARCH + ["-shared", "-o", out] + objs + ["-cudart", "static"])
    shutil.rmtree(work)
    print("built", out)


if __name__ == "__main__":
    main()
