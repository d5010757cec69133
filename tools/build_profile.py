"""Build-only driver for profiling: C3 terrain (or --scale), N rebuilds in the given mode."""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rtk_b200 import api, scenes
ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="sah")
ap.add_argument("--config", default="C3")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--rebuilds", type=int, default=3)
a = ap.parse_args()
lib = api.load()
assert lib.rtk_cuda_init(0) == 0, lib.last_error()
lib.rtk_cuda_set_build_mode(1 if a.mode == "sah" else 0)
s = scenes.config_scene(a.config, a.scale)
t = time.perf_counter(); sc = lib.build_scene(s["meshes"]); print("first build wall ms", (time.perf_counter() - t) * 1e3, "device ms", sc.info().build_device_ms)
for i in range(a.rebuilds):
    t = time.perf_counter(); assert lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0; w = (time.perf_counter() - t) * 1e3
    i_ = sc.info(); print("rebuild wall ms %.3f device ms %.3f nodes %d leaves %d depth %d sah %.2f" % (w, i_.build_device_ms, i_.num_wide_nodes, i_.num_leaves, i_.wide_depth, i_.sah_cost))
