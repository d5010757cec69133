"""Build + two rebuilds of one workload's scene: the command behind the build launch list."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rtk_b200 import api, scenes  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "C3"
mode = api.RTK_CUDA_BUILD_SAH if (len(sys.argv) < 3 or sys.argv[2] == "sah") else api.RTK_CUDA_BUILD_LBVH
lib = api.load() if not os.environ.get("RTK_LIB") else api.Library(os.path.abspath(os.environ["RTK_LIB"]))
assert lib.rtk_cuda_init(0) == 0, lib.last_error()
lib.rtk_cuda_set_build_mode(mode)
s = scenes.config_scene(workload)
sc = lib.build_scene(s["meshes"])
ms = []
for i in range(3):
    assert lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0, lib.last_error()
    ms.append(sc.info().build_device_ms)
print(workload, "rebuild device ms:", " ".join("%.3f" % m for m in ms))
sc.free()
