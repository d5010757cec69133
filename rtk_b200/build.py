"""Compile librtk_b200.so in-tree: CUDA kernels + thin C-ABI layer with nvcc for sm_100a, the C
host layer with gcc, linked into one shared library (static cudart).  Cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librtk_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
SOURCES = ["k_sah.cuh", "rtk_device.cu", "rtk_host.c", "rtk_place.c", "rtk_common.cuh", "rtk_math.cuh", "k_build.cuh", "k_trace.cuh", "k_wavefront.cuh", "rtk_device.h"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: experiment variants, e.g. build(defines=["-DRTK_TRACE_MINB=4"], out="librtk_b200_m4.so")"""
    global OUT
    saved = OUT
    if out:
        OUT = os.path.join(HERE, out)
        force = True
    try:
        return _build(force, verbose, list(defines))
    finally:
        OUT = saved


def _build(force, verbose, defines):
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "..", "include", h) for h in ("rtk.h", "rtk_cuda.h")]
    if not force and not _newer(OUT, deps):
        return OUT
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    tag = os.path.basename(OUT).replace(".so", "")
    dev_o = os.path.join(bdir, tag + "_device.o")
    host_o = os.path.join(bdir, tag + "_host.o")
    place_o = os.path.join(bdir, tag + "_place.o")
    cmds = [
        [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v", *defines,
         "-c", os.path.join(CSRC, "rtk_device.cu"), "-o", dev_o],
        ["gcc", "-O2", "-fPIC", "-Wall", "-std=gnu11", "-c", os.path.join(CSRC, "rtk_host.c"), "-o", host_o],
        ["gcc", "-O2", "-fPIC", "-Wall", "-std=gnu11", "-c", os.path.join(CSRC, "rtk_place.c"), "-o", place_o],
        [NVCC, *ARCH, "-shared", "-o", OUT, dev_o, host_o, place_o, "-lpthread", "-ldl"],
    ]
    for c in cmds:
        r = subprocess.run(c, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(c) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("rtk_b200 build failed: " + " ".join(c))
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
