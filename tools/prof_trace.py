"""One workload, N identical k_trace launches: the command profiled with ncu (tools/experiments/exp*.sh).
usage: python tools/prof_trace.py [workload] [launches] [rays]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rtk_b200 import api, scenes  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "C3"
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 24
lib = api.load() if not os.environ.get("RTK_LIB") else api.Library(os.path.abspath(os.environ["RTK_LIB"]))
assert lib.rtk_cuda_init(0) == 0, lib.last_error()
lib.rtk_cuda_set_build_mode(api.RTK_CUDA_BUILD_SAH)
s = scenes.config_scene(workload)
if workload == "C2":
    rays = scenes.config_rays("C2", s)
elif workload == "C3":
    rays = scenes.bounce_rays(s, n)
else:
    rays = scenes.mixed_rays(s, n)
sc = lib.build_scene(s["meshes"])
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32)).cuda()
d_hit = torch.zeros((len(rays), 16), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(launches + 1)]
ev[0].record()
for i in range(launches):
    assert lib.rtk_trace_rays_compact_device(sc.ptr, d_rays.data_ptr(), d_hit.data_ptr(), len(rays), st) == 0, lib.last_error()
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(launches)]
print(workload, len(rays), "rays; k_trace ms per launch:", " ".join("%.2f" % m for m in ms),
      "| Mrays/s (last): %.1f" % (len(rays) / ms[-1] / 1e3))
sc.free()
