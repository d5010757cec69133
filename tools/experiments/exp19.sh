# Round 2, first call (1 GPU, ~3 min of box time): everything written after round 1's GPU budget ran out.
#   gpurun --timeout 420 -- 'sh tools/experiments/exp19.sh'
# 1. the whole GPU tier (new cases: baked instancing, concurrent host threads, pathological scenes,
#    compact host batch, small-batch path);  2. the default bench line (look at e2e_compact and
#    e2e);  3. latency of rtk_trace_ray on the small-batch path.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/exp19_pytest.log 2>&1; tail -3 gpurun_out/exp19_pytest.log
python bench.py --steps 30 --warmup 3 > gpurun_out/exp19_bench.json 2> gpurun_out/exp19_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/exp19_bench.json"))
print("value %.1f Mrays/s | k_trace %.3f ms | e2e %.1f (%.2f ms) | e2e_compact %s" % (
    d["value"], d["kernels_ms"]["k_trace"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("e2e_compact")))
PY
python - <<'PY'
import time, ctypes as C, numpy as np
from rtk_b200 import api, scenes
lib = api.load(); assert lib.rtk_cuda_init(0) == 0
s = scenes.config_scene("C3", 0.1)
sc = lib.build_scene(s["meshes"])
rays = scenes.bounce_rays(s, 2000)
h = np.zeros(1, dtype=api.HIT_DTYPE)
for n in (1, 64, 2048, 4096):
    sub = np.ascontiguousarray(rays[:n]) if n <= len(rays) else np.ascontiguousarray(np.resize(rays, n))
    sc.trace_rays(sub)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        sc.trace_rays(sub)
    print("rtk_trace_rays n=%d: %.1f us per call" % (n, (time.perf_counter() - t0) / reps * 1e6))
PY
