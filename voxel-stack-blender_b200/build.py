"""Builds libvsb_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build().

    python voxel-stack-blender_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvsb_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["vsb_api.cu", "vsb_mask.cu", "vsb_zblend.cu", "vsb_dt.cu", "vsb_edt.cu", "vsb_ccl.cu", "vsb_xy.cu", "vsb_layer.cu", "vsb_pack_tma.cu", "vsb_unbounded.cu", "vsb_geom.cu", "vsb_roi.cu", "vsb_synth.cu",
           "vsb_stack.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: the float32 fade arithmetic must round after every operation exactly like NumPy does
# (SURVEY.md 7.3 item 2); the one place that needs an FMA (bilateral) asks for it explicitly.
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-cudart", "static", "-ccbin", "g++"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "vsb.h"))
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ccbin", "g++", "-o", OUT] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
