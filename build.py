#!/usr/bin/env python
"""Builds como_b200/libcomo_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "como_b200", "csrc")
OUT = os.path.join(ROOT, "como_b200", "libcomo_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


# depthcov.cu holds the bit-exact fp32 anchor-selection path: no FMA contraction there
PER_FILE = {"depthcov.cu": ["-fmad=false"]}


def needs_build(srcs):
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_v=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if not force and not needs_build(srcs):
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(ROOT, "build", os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if (not force) and os.path.exists(o) and os.path.getmtime(o) > max(
            [os.path.getmtime(s)] + [os.path.getmtime(d) for d in glob.glob(os.path.join(CSRC, "*.cuh"))]
            + [os.path.getmtime(d) for d in glob.glob(os.path.join(ROOT, "include", "*.h"))]):
            continue
        cmd = [NVCC] + FLAGS + PER_FILE.get(os.path.basename(s), []) + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    fail = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            fail = True
        if out.strip() and (verbose or ptxas_v or p.returncode != 0):
            print(out)
    if fail:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_v="--ptxas" in sys.argv)
    print("built", OUT)
