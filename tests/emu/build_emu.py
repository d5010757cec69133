"""Compile the product's CUDA sources against the SIMT emulator (tests/emu/simt.h) into
tests/emu/librtk_emu.so so that the CPU-only test tier can execute the very same kernels and host
logic through the very same C ABI.  TEST INFRASTRUCTURE ONLY -- the product never loads this."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "rtk_b200", "csrc")
OUT = os.path.join(HERE, "librtk_emu.so")


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "simt.h")]
    deps += [os.path.join(ROOT, "include", h) for h in ("rtk.h", "rtk_cuda.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    dev_o, host_o = os.path.join(bdir, "rtk_device_emu.o"), os.path.join(bdir, "rtk_host_emu.o")
    place_o = os.path.join(bdir, "rtk_place_emu.o")
    cmds = [
        ["g++", "-x", "c++", "-std=c++17", "-O2", "-g", "-fPIC", "-ffp-contract=off", "-mfma", "-w",
         *os.environ.get("RTK_EMU_DEFINES", "").split(),
         "-DRTK_SIMT_EMU=1", "-DSIMT_IMPL=1", "-include", os.path.join(HERE, "simt.h"),
         "-c", os.path.join(CSRC, "rtk_device.cu"), "-o", dev_o],
        ["gcc", "-O2", "-g", "-fPIC", "-std=gnu11", "-c", os.path.join(CSRC, "rtk_host.c"), "-o", host_o],
        ["gcc", "-O2", "-g", "-fPIC", "-std=gnu11", "-c", os.path.join(CSRC, "rtk_place.c"), "-o", place_o],
        ["g++", "-shared", "-o", OUT, dev_o, host_o, place_o, "-lpthread", "-lm"],
    ]
    for c in cmds:
        r = subprocess.run(c, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(c) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("emulator build failed")
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
