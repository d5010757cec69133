"""rtk_trace_rays / rtk_trace_rays_compact on page-locked host arrays, one workload, a few repetitions: the command
behind the host-path experiments (RTK_B200_HOST_DIRECT=0/1, RTK_B200_HOST_CHUNK_LOG2, device lists).
usage: python tools/prof_e2e.py [workload] [rays] [devices] [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rtk_b200 import api, scenes  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "C3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
lib = api.load() if not os.environ.get("RTK_LIB") else api.Library(os.path.abspath(os.environ["RTK_LIB"]))
devs = (C.c_int * ndev)(*range(ndev))
assert lib.rtk_cuda_init_devices(devs, ndev) == 0, lib.last_error()
s = scenes.config_scene(workload)
one = scenes.bounce_rays(s, min(n, 1 << 22)) if workload == "C3" else scenes.mixed_rays(s, min(n, 1 << 22), threads=8)


def pinned(count, dtype):
    dt = np.dtype(dtype)
    p = lib.rtk_cuda_host_alloc_batch(dt.itemsize, count)
    assert p, lib.last_error()
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(count * dt.itemsize,)).view(dt)


rays = pinned(n, api.RAY_DTYPE)
for lo in range(0, n, len(one)):
    k = min(len(one), n - lo)
    rays[lo:lo + k] = one[:k]
hits, mask, h16 = pinned(n, api.HIT_DTYPE), pinned(n, np.uint8), pinned(n, api.HIT16_DTYPE)
t0 = time.perf_counter()
sc = lib.build_scene(s["meshes"])
print("build + replicate over %d device(s): %.1f ms" % (ndev, (time.perf_counter() - t0) * 1e3))
for name, fn in (("rows", lambda: lib.rtk_trace_rays(sc.ptr, rays.ctypes.data, hits.ctypes.data, mask.ctypes.data, n)),
                 ("rows, no mask", lambda: lib.rtk_trace_rays(sc.ptr, rays.ctypes.data, hits.ctypes.data, None, n)),
                 ("compact", lambda: lib.rtk_trace_rays_compact(sc.ptr, rays.ctypes.data, h16.ctypes.data, n))):
    ms = []
    for i in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ms.append((time.perf_counter() - t0) * 1e3)
        assert r != C.c_size_t(-1).value and r >= 0, lib.last_error()
    print("%-14s %d rays on %d device(s), direct=%s mix=%s numa=%s: ms %s | best %.1f Mrays/s" %
          (name, n, ndev, os.environ.get("RTK_B200_HOST_DIRECT", "1"), os.environ.get("RTK_B200_HOST_MIX", "auto"), os.environ.get("RTK_B200_NUMA", "1"),
           " ".join("%.2f" % m for m in ms), n / min(ms) / 1e3))
for d in (1, 2, 3):
    for k in sorted(set([1, ndev])):
        g = C.c_double(0)
        lib.rtk_cuda_measure_host_link(k, 256 << 20, d, 4, C.byref(g))
        print("host link, %d device(s), %s: %.1f GB/s" % (k, {1: "h2d", 2: "d2h", 3: "both"}[d], g.value))
sc.free()
