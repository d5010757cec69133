#!/usr/bin/env python
"""bench.py -- closest-hit Mrays/s (incoherent) on the BASELINE.json workload, plus BVH build Mtris/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rays R] [--scale S]

A step is one pass of the hot path over one batch of synthetic rays: the traversal kernel
(k_trace) followed by the hit-expansion kernel (k_resolve), inputs already resident in HBM.
N=1 workload: BASELINE.json configs[2] -- 1M-triangle procedural terrain, 16 777 216 incoherent
diffuse-bounce rays (the configuration the headline "incoherent Mrays/s" metric is quoted on).
N>1 (torchrun): scene replicated per GPU, every rank traces its own 16 777 216 rays (weak
scaling), compact hit records are gathered on rank 0 inside the timed region.

Beside the headline the line carries
  e2e          the same batch(es) through the host API rtk_trace_rays (host rtk_ray[] in, rtk_hit[] out).  For
               N > 1 this is ONE process (rank 0) driving all N GPUs through rtk_cuda_init_devices -- the
               library's own multi-GPU path -- on all N x 16 777 216 rays, with the measured PCIe ceiling;
  c4           BASELINE configs[3]: 10M-triangle terrain in two meshes, GPU build + 67 108 864 mixed rays, the
               FIXED batch split over the N ranks (strong scaling), with its own parity and roofline;
  c5           (N = 1) BASELINE configs[4]: one band of the 4K x 16 spp x 4 bounce wavefront, rays generated
               on the device;
  roofline, build.roofline, cpu_baseline, parity (>= 10 240 oracle rays), clocks (NVML, 5 ms).

`--impl reference` times the reference's own CPU path instead: the patched rtk.c traversal
(oracle/_ref, built from /root/reference/rtk.c) over a reference-format blob packed by the
oracle's binned-SAH restatement, all host threads, on a bounded sample of the same rays.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rtk_b200 import scenes  # noqa: E402

METRIC = "closest-hit Mrays/s (incoherent)"
UNIT = "Mrays/s"
FULL_RAYS = 16_777_216
C4_RAYS = 67_108_864


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profile_summary(workload, wide_nodes):
    """ncu counters of the dominant kernel for this workload (profiles/r2_summary.json, written from the
    committed captures).  Only returned when the capture was taken on the very tree this run built
    (same number of wide nodes): a stale capture is worse than none."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_summary.json")) as f:
            e = json.load(f).get(workload)
        if e and int(e.get("wide_nodes", -1)) == int(wide_nodes):
            return e
    except Exception:
        pass
    return None


WORKLOADS = {
    # name: (scene config, default rays per GPU, description)
    "C3": ("C3", FULL_RAYS, "C3: 1M-triangle procedural terrain (1001x501 value-noise heightfield), "
                            "16777216 incoherent diffuse-bounce rays per GPU"),
    "C2": ("C2", 2_073_600, "C2: 1M-triangle random soup, 1920x1080 coherent primary rays per GPU"),
    "C4": ("C4", C4_RAYS, "C4: 10M-triangle procedural terrain in 2 meshes (2501x2001 vertices), 67108864 mixed rays "
                          "per GPU (thirds of coherent primary / diffuse bounce / short segments, 64Ki blocks)"),
    # wavefront path tracing: rays are generated on the device and never leave it (run_wavefront)
    "C5": ("C4", 66_355_200, "C5: 4K frame (3840x2160) x 16 spp x 4 bounces on the 10M-triangle terrain, pixel rows "
                             "split in 8 bands; one band (3840x270 pixels, 66355200 rays per step) per GPU"),
}

# declared bytes per triangle of the SAH build ("Build"): what the kernels read and
# write by construction, not a counter
BUILD_BYTES_PER_TRI = {
    "decode (36 + 12 idx in, 48 out)": 96, "scene bounds": 48, "morton (48 in, 12 out)": 60,
    "radix sort, 3 passes x (hist 8 + scatter 20 in + 12 out)": 120, "prim bounds (48 in, 36 out)": 84,
    "binned SAH, large levels (~13 x 52: binning 4 + 32 in, 4 out; partition 8 in, 4 out; on the shrinking large set, "
    "measured average 5.5 full passes)": 286,
    "small subtrees (36 in, 4 out)": 40, "collapse records (0.4 binary nodes per triangle x (40 in, 17 out))": 23,
    "collapse + leaf emission (~64 + 52 in, ~80 + 51 out)": 247,
}


def gen_rays(scene, n, rank=0, workload="C3", threads=8):
    if workload == "C2":
        r = scenes.soup_primary_rays()
        return np.ascontiguousarray(np.resize(r, n))
    if workload == "C4":
        return scenes.mixed_rays(scene, n, seed=0xD4 + 0x1000 * rank, threads=threads)
    out = np.empty(n, dtype=scenes.RAY_DTYPE)
    step = 1 << 20

    def one(lo):
        hi = min(n, lo + step)
        out[lo:hi] = scenes.bounce_rays(scene, hi - lo, seed=0xD3, first=rank * n + lo)
    if threads > 1 and n > step:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(0, n, step)))
    else:
        for lo in range(0, n, step):
            one(lo)
    return out


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML every few
    milliseconds (a 0.2 s region gets dozens of samples); nvidia-smi as the fallback."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period=0.005):
        self.index, self.period = index, period
        self.rows, self.first = [], 0
        self.nv, self.handle, self.proc, self.thread, self.stopf = None, None, None, None, False
        self.mx = None

    def _nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        self.nv, self.handle = pynvml, h

    def mark(self):
        """samples from here on belong to the timed region"""
        self.first = len(self.rows)

    def start(self):
        try:
            self._nvml()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv, self.handle
        names = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
        while not self.stopf:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((sm, [n for b, n in names if bits & b]))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8])
                           if v.lower().startswith("active")]
                self.mx = float(r[1])
                self.rows.append((float(r[0]), reasons))
            except Exception:
                pass

    def stop(self):
        self.stopf = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=1)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"], "samples": 0}
        rows = self.rows[self.first:] if len(self.rows) - self.first >= 3 else self.rows
        reasons = set()
        for _, rs in rows:
            reasons.update(rs)
        return {"sm_mhz": float(np.median([r[0] for r in rows])), "sm_max_mhz": self.mx, "reasons": sorted(reasons),
                "samples": len(rows), "source": "nvml, in-process, %.0f ms period" % (self.period * 1e3) if self.nv else "nvidia-smi -lms 20"}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline (the one place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------

def cpu_reference_run(scene, rays, sample, steps, warmup):
    """Patched rtk.c traversal (oracle/_ref) over a blob from the oracle's binned-SAH packer."""
    from oracle import orc
    cores = orc.num_cores()
    blob = orc.ReferenceBlob(scene["tris"])
    build_s = blob.stats.build_seconds
    kind = "reference" if orc.have_reference(patched=True) else None
    if kind is None:
        raise RuntimeError("oracle/_ref/librtk_ref_patched.so is missing (build it where /root/reference exists)")
    sub = rays[:sample]
    times = []
    for i in range(warmup + steps):
        _, sec = blob.trace(sub, patched=True, threads=cores, want_hits=False)
        if i >= warmup:
            times.append(sec)
    blob.close()
    sec = float(np.mean(times))
    return {"value": sample / sec / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"first {sample} of the {len(rays)} rays of the workload, {cores} threads, "
                      f"rtk.c traversal (6-line stack fix) over a binned-SAH BVH4 blob from the oracle packer",
            "ms_per_step": sec * 1e3,
            "build": {"value": len(scene["tris"]) / build_s / 1e6, "unit": "Mtris/s", "cores": 1,
                      "kind": "port", "what": "rtk.c binned-SAH algorithm restated (oracle), blob packing included"}}


def run_reference(args, workload, scene, rays):
    sample = min(len(rays), args.ref_sample)
    cb = cpu_reference_run(scene, rays, sample, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "cpu_build_baseline": cb["build"],
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# parity beside the numbers (outside every timed region)
# ------------------------------------------------------------------------------------------------

def oracle_parity(api, scene, rays_np, got16, k, flat_rays=1024):
    """`got16`: the device's compact records of rays_np[:len(got16)].  k sampled rays (evenly spread) against
    the CPU oracle, bit for bit; where oracle/_ref exists, a few of them also against the UNMODIFIED rtk.c
    leaf code (flat blobs): index exact, t within the north star's 1e-5."""
    from oracle import orc
    n = len(got16)
    k = min(k, n)
    idx = np.unique(np.linspace(0, n - 1, k).astype(np.int64))
    t0 = time.perf_counter()
    want = orc.trace_brute(scene["tris"], np.ascontiguousarray(rays_np[idx]))
    sec = time.perf_counter() - t0
    got = got16[idx]
    out = {"rays_checked": int(len(idx)), "index_mismatches": int((got["prim"] != want["prim"]).sum()),
           "bit_exact": bool(got.tobytes() == want.tobytes()), "against": "oracle (CPU brute force over every triangle)",
           "oracle_seconds": round(sec, 2)}
    if flat_rays > 0 and orc.have_reference():
        try:
            j = idx[:: max(1, len(idx) // flat_rays)][:flat_rays]
            lit = orc.trace_flat_reference(scene["tris"], np.ascontiguousarray(rays_np[j]))
            g = got16[j]
            hit = lit["prim"] != orc.MISS
            rel = np.abs(g["t"][hit] - lit["t"][hit]) / np.maximum(np.abs(lit["t"][hit]), 1e-30)
            out["unmodified_rtk_c_leaf_code"] = {"rays": int(len(j)), "index_mismatches": int((g["prim"] != lit["prim"]).sum()),
                                                 "max_rel_t_error": float(rel.max()) if hit.any() else 0.0,
                                                 "tolerance": 1e-5}
        except Exception as ex:
            out["unmodified_rtk_c_leaf_code"] = {"error": str(ex)}
    return out


def gpu_bruteforce_parity(lib, api, torch, sc, d_rays, d_h16, kb, sh):
    d_b = torch.zeros((kb, 16), dtype=torch.uint8, device="cuda")
    if lib.rtk_trace_rays_bruteforce_device(sc.ptr, d_rays.data_ptr(), d_b.data_ptr(), kb, sh) != 0:
        raise RuntimeError(lib.last_error())
    torch.cuda.synchronize()
    a = d_h16[:kb].cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
    b = d_b.cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
    return {"gpu_bruteforce_rays": int(kb), "gpu_bruteforce_index_mismatches": int((a["prim"] != b["prim"]).sum()),
            "gpu_bruteforce_bit_exact": bool(a.tobytes() == b.tobytes())}


# ------------------------------------------------------------------------------------------------
# C5: wavefront path tracing, rays generated on the device
# ------------------------------------------------------------------------------------------------

def terrain_camera(api, scene, width, height):
    """the camera of scenes.terrain_primary_rays as an rtk_cuda_camera"""
    lo, hi = scenes.scene_bounds(scene)
    c = (lo + hi) / 2
    eye = np.array([c[0], hi[1] + 0.8 * (hi[2] - lo[2]), lo[2] - 0.6 * (hi[2] - lo[2])], dtype=np.float32)
    fwd = c - eye
    fwd = fwd / np.linalg.norm(fwd)
    right = np.cross([0.0, 1.0, 0.0], fwd)
    right /= np.linalg.norm(right)
    upv = np.cross(fwd, right)
    cam = api.rtk_cuda_camera()
    cam.eye[:], cam.forward[:], cam.right[:], cam.up[:] = [float(x) for x in eye], [float(x) for x in fwd], \
        [float(x) for x in right], [float(x) for x in upv]
    cam.tan_half_fov, cam.width, cam.height = float(np.tan(np.radians(50.0) / 2)), width, height
    return cam


def wavefront_leg(args, lib, api, scene, sc, rank, world, steps, warmup, gather=True):
    """A step = one band of the 4K frame: 16 jittered primary rays per pixel generated on the device,
    then 4 x (k_trace, k_gen_bounce with relaunch of the paths that left the scene).  Rays and hits
    never leave HBM; for N > 1 the last bounce's compact hit records are gathered on rank 0.
    Returns (rays per step, ms per step [max over ranks], counts, parity)."""
    import torch
    import torch.distributed as dist
    W, H, SPP, BOUNCES, BANDS = 3840, 2160, 16, 4, 8
    if args.wf_frame:                                       # dry runs only: the workload is the 4K frame
        W, H, SPP = (int(v) for v in args.wf_frame.split("x"))
    band = (rank % BANDS) if world > 1 else 0
    rows = H // BANDS
    npx = W * rows
    n = npx * SPP
    cam = terrain_camera(api, scene, W, H)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    bufs = [torch.empty((n, 32), dtype=torch.uint8, device="cuda") for _ in range(2)]
    d_h16 = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    d_alive = torch.empty((n,), dtype=torch.uint8, device="cuda")
    gather_list = [torch.empty_like(d_h16) for _ in range(world)] if (world > 1 and rank == 0 and gather) else None
    seed = 0xD5
    counts = {}

    def step():
        for s_ in range(SPP):
            rc = lib.rtk_cuda_generate_primary_rays(C.byref(cam), seed, s_, band * npx, npx,
                                                    bufs[0].data_ptr() + 32 * npx * s_, sh)
            if rc:
                raise RuntimeError(lib.last_error())
        cur = 0
        for b in range(BOUNCES):
            rc = lib.rtk_trace_rays_compact_device(sc.ptr, bufs[cur].data_ptr(), d_h16.data_ptr(), n, sh)
            if b + 1 < BOUNCES:
                rc |= lib.rtk_cuda_generate_bounce_rays(sc.ptr, bufs[cur].data_ptr(), d_h16.data_ptr(), bufs[cur ^ 1].data_ptr(),
                                                        d_alive.data_ptr(), n, seed, b, band * n, api.RTK_CUDA_BOUNCE_RELAUNCH, sh)
                cur ^= 1
            if rc:
                raise RuntimeError(lib.last_error())
        if world > 1 and gather:
            dist.gather(d_h16, gather_list, dst=0)
    step()
    torch.cuda.synchronize()
    counts["last_bounce_hit_fraction"] = float((d_h16.view(torch.int32)[:, 3] != -1).float().mean().item())
    counts["relaunched_fraction_bounce3"] = float((d_alive == 2).float().mean().item())

    # parity: the last bounce's rays (as the device generated them) against the exhaustive GPU kernel
    # and, for a few, against the CPU oracle
    parity = None
    if rank == 0 and args.parity_rays > 0:
        last = bufs[(BOUNCES - 1) & 1]
        kb = min(n, 65536)
        parity = gpu_bruteforce_parity(lib, api, torch, sc, last, d_h16, kb, sh)
        k = min(args.parity_rays, 1024, n)
        rays_np = last[:kb].cpu().numpy().view(api.RAY_DTYPE).reshape(-1)
        got = d_h16[:kb].cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
        parity.update(oracle_parity(api, scene, rays_np, got, k, flat_rays=0))
        parity["against"] += " on device-generated bounce-3 rays"

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(stream)
    for i in range(steps):
        step()
    e_end.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = e_start.elapsed_time(e_end)
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    return n * BOUNCES, total_ms / steps, counts, parity, (SPP + 2 * BOUNCES - 1)


def run_wavefront(args, workload, lib, api, scene, rank, world, local_rank):
    """--workload C5 as the headline of a run"""
    sc = lib.build_scene(scene["meshes"])
    assert lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0, lib.last_error()     # device time of a warm build
    info = sc.info()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark()
    rays_per_step, ms_per_step, counts, parity, launches = wavefront_leg(args, lib, api, scene, sc, rank, world, args.steps, args.warmup)
    clocks = sampler.stop()
    if rank != 0:
        return
    workload = dict(workload)
    workload.update({"triangles": int(len(scene["tris"])), "rays_per_gpu": rays_per_step,
                     "l2_policy": "ray, hit and alive buffers (%.1f GB per bounce) stream through the 126 MB L2; the 10M-triangle "
                                  "scene (%.2f GB) does not fit it" % (rays_per_step / 4 * 81 / 1e9, info.device_bytes / 1e9)})
    line = {"metric": "closest-hit Mrays/s (wavefront, device-generated rays)", "value": world * rays_per_step / (ms_per_step * 1e-3) / 1e6,
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload, "gpu_launches": launches * args.steps,
            "e2e": None, "wavefront": counts, "parity": parity, "clocks": clocks,
            "build": {"metric": "BVH build Mtris/s", "value": len(scene["tris"]) / (info.build_device_ms * 1e-3) / 1e6,
                      "unit": "Mtris/s", "device_ms": info.build_device_ms}}
    emit(line)
    sc.free()


# ------------------------------------------------------------------------------------------------
# build numbers: cold and warm end to end (host mesh buffers in), device-only rebuilds, declared roofline
# ------------------------------------------------------------------------------------------------

def build_numbers(lib, scene, mode, peak):
    ntris = int(len(scene["tris"]))
    t0 = time.perf_counter()
    sc = lib.build_scene(scene["meshes"])
    cold_s = time.perf_counter() - t0
    warm = []
    for _ in range(2):
        t0 = time.perf_counter()
        s2 = lib.build_scene(scene["meshes"])
        warm.append(time.perf_counter() - t0)
        s2.free()
    build_ms = []
    for i in range(5):
        assert lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0, lib.last_error()
        build_ms.append(sc.info().build_device_ms)
    info = sc.info()
    dev_ms = float(np.median(build_ms[1:]))
    bpt = float(sum(BUILD_BYTES_PER_TRI.values()))
    tris_s = ntris / (dev_ms * 1e-3)
    out = {"metric": "BVH build Mtris/s", "value": tris_s / 1e6, "unit": "Mtris/s", "device_ms": dev_ms, "mode": mode,
           "e2e": {"value": ntris / min(warm) / 1e6, "unit": "Mtris/s", "ms": min(warm) * 1e3,
                   "api": "rtk_build_scene (host mesh buffers in: upload + decode + build), warm: best of 2 after the first call",
                   "first_call_ms": cold_s * 1e3,
                   "first_call_note": "includes CUDA context and module load"},
           "roofline": {"bound": "hbm", "achieved": bpt * tris_s / 1e9, "peak": peak, "unit": "GB/s", "frac": bpt * tris_s / 1e9 / peak,
                        "bytes_per_triangle": bpt, "declared_passes": BUILD_BYTES_PER_TRI,
                        "note": "declared bytes (what the kernels read and write by construction) x triangles/s; the build is "
                                "bound by kernel launches, shared-memory atomics of the binning and device-side latency, "
                                "not by HBM bytes"},
           "wide_nodes": int(info.num_wide_nodes), "leaves": int(info.num_leaves), "depth": int(info.wide_depth),
           "sah_cost": info.sah_cost, "scene_bytes": int(info.device_bytes)}
    return sc, info, out


def trace_stats(lib, api, sc, d_rays, d_h16, n, sh):
    st = api.rtk_cuda_trace_stats()
    ns = min(n, 1 << 20)
    assert lib.rtk_trace_stats_device(sc.ptr, d_rays.data_ptr(), d_h16.data_ptr(), ns, C.byref(st), sh) == 0, lib.last_error()
    per = {"wide_node_visits": st.node_visits / ns, "leaf_visits": st.leaf_visits / ns,
           "triangle_tests": st.tri_tests / ns, "hit_fraction": st.hits / ns}
    return per, 32 + 16 + 256 * per["wide_node_visits"] + 48 * per["triangle_tests"]


def roofline_block(lib, workload_key, info, per_ray, bytes_per_ray, n, trace_ms, hbm_peak, peak_src, probe_bytes):
    """The dominant kernel against the roof that binds it.  The traversal fetches 256-byte nodes and 128-byte
    leaf lines at scattered addresses; for a scene that lives in the L2 the roof is the L2's random-gather
    bandwidth (measured here with the library's gather probe over a buffer of the scene's size), for a scene
    that does not fit it the same probe over 1 GiB is the HBM gather roof.  DRAM traffic comes from the
    committed ncu capture of this very tree, or null."""
    g256, g128, stream_l2, hbm_read = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0)
    lib.rtk_cuda_measure_gather_bandwidth(probe_bytes, 256, 3, C.byref(g256))
    lib.rtk_cuda_measure_gather_bandwidth(probe_bytes, 128, 3, C.byref(g128))
    lib.rtk_cuda_measure_read_bandwidth(64 << 20, 40, C.byref(stream_l2))
    lib.rtk_cuda_measure_read_bandwidth(4 << 30, 3, C.byref(hbm_read))
    in_l2 = info.device_bytes < 120e6
    achieved = bytes_per_ray * n / (trace_ms * 1e-3) / 1e9
    prof = profile_summary(workload_key, info.num_wide_nodes)
    compulsory = 32.0 * n + 16.0 * n + float(info.device_bytes) * (1.0 if in_l2 else 0.0)
    # the capture may have been taken with another ray count: DRAM traffic of this kernel is per ray (rays in, hits out,
    # scene misses), so it is scaled to the launch this run timed
    dram_bytes = prof["dram_bytes"] * (n / prof["rays"]) if prof else None
    blk = {"bound": "l2" if in_l2 else "hbm", "kernel": "k_trace", "achieved": achieved, "peak": g256.value, "unit": "GB/s",
           "frac": achieved / g256.value if g256.value else None,
           "traffic": dram_bytes,
           "traffic_source": (prof.get("source") if prof else "no ncu capture of this tree (%d wide nodes) under profiles/" % info.num_wide_nodes),
           "peak_source": "measured in this run: warps gathering 256-byte records at hashed offsets of a %d MiB buffer "
                          "(rtk_cuda_measure_gather_bandwidth; %s); 128-byte records reach %.0f GB/s, a coalesced stream over "
                          "64 MiB %.0f GB/s, over 4 GiB (HBM) %.0f GB/s" % (probe_bytes >> 20, "inside the L2" if in_l2 else "far beyond the L2",
                                                                          g128.value, stream_l2.value, hbm_read.value),
           "limiter": ("instruction issue: %.0f %% of the issue slots busy, %.0f warp instructions per ray at %.1f of 32 threads each, ALU pipe %.0f %%; "
                       "L2 hit rate %.0f %%, L2 throughput %.0f %%, DRAM %.0f %% of the measured HBM peak (ncu, %s)"
                       % (prof["issue_active_pct"], prof["warp_instructions_per_ray"], prof["threads_per_instruction"], prof["pipe_alu_pct"],
                          prof["l2_hit_rate_pct"], prof.get("l2_throughput_pct") or 0.0,
                          100.0 * prof["dram_bytes"] / (prof["duration_ms"] * 1e-3) / 1e9 / hbm_peak, prof["source"])) if prof else
                      ("instruction issue (no ncu capture of this tree at hand)" if in_l2 else "instruction issue and memory latency (no ncu capture of this tree at hand)"),
           "issue_active_pct": prof.get("issue_active_pct") if prof else None,
           "warp_instructions_per_ray": prof.get("warp_instructions_per_ray") if prof else None,
           "dram": {"bytes_per_launch": dram_bytes, "compulsory_bytes_per_launch": compulsory,
                    "gbs": (dram_bytes / (trace_ms * 1e-3) / 1e9) if dram_bytes else None, "peak": hbm_peak,
                    "frac": (dram_bytes / (trace_ms * 1e-3) / 1e9 / hbm_peak) if dram_bytes else None, "peak_source": peak_src},
           "hbm_equivalent": {"achieved": achieved, "peak": hbm_peak, "frac": achieved / hbm_peak,
                              "note": "algorithmic bytes against the HBM copy peak, for comparison with round 1 only: "
                                      "these bytes are served by the L2, not by HBM" if in_l2 else "algorithmic bytes against the HBM copy peak"},
           "bytes_per_ray": bytes_per_ray, "per_ray": per_ray,
           "note": "algorithmic bytes = 32 (ray) + 16 (compact hit) + 256 per wide-node visit + 48 per triangle tested, "
                   "counted by the instrumented kernel on the first 2^20 rays; scene %.0f MB, L2 126 MB" % (info.device_bytes / 1e6)}
    return blk


# ------------------------------------------------------------------------------------------------
# the sharded device path: every rank traces its rays, compact records are gathered on rank 0
# ------------------------------------------------------------------------------------------------

class ShardedTrace:
    """k_trace + k_resolve per step on this rank's rays; for N > 1 the compact records of step k are
    gathered on rank 0 (NCCL, or copy-engine pushes into a peer window) while step k+1 traces."""

    def __init__(self, args, lib, torch, dist, sc, rays_np, rank, world, gather):
        self.lib, self.torch, self.dist, self.sc, self.rank, self.world = lib, torch, dist, sc, rank, world
        n = self.n = len(rays_np)
        self.stream = torch.cuda.current_stream()
        self.sh = self.stream.cuda_stream
        self.d_rays = torch.from_numpy(rays_np.view(np.uint8).reshape(-1, 32)).cuda()
        # compact hits are double buffered so that, for N > 1, the gather of step k runs on the
        # communication stream while step k+1 traces into the other buffer
        self.d_h16s = [torch.zeros((n, 16), dtype=torch.uint8, device="cuda") for _ in range(2 if world > 1 else 1)]
        self.d_hits = torch.zeros((n, 68), dtype=torch.uint8, device="cuda")
        self.d_mask = torch.zeros((n,), dtype=torch.uint8, device="cuda")
        self.gather_lists = [None, None]
        self.p2p = world > 1 and gather == "p2p"
        self.peer, self.copy_stream = None, None
        if self.p2p:
            from rtk_b200 import shard

            def exchange(handle):
                box = [handle]
                dist.broadcast_object_list(box, src=0)
                return box[0]
            self.peer = shard.PeerGather(lib, rank, world, 16 * n, 2, exchange)
            self.copy_stream = torch.cuda.Stream()
            self.traced_ev = [torch.cuda.Event() for _ in range(2)]
            self.pushed_ev = [None, None]
        elif world > 1 and rank == 0:
            self.gather_lists = [[torch.empty_like(self.d_h16s[0]) for _ in range(world)] for _ in range(2)]
        self.pending = [None, None]
        self.counter = 0

    def step(self, ev=None, exchange=True):
        lib, torch, dist = self.lib, self.torch, self.dist
        world, rank, n, stream, sh = self.world, self.rank, self.n, self.stream, self.sh
        i = self.counter & 1 if (world > 1 and exchange) else 0
        if self.pending[i] is not None:
            self.pending[i].wait()                # the buffer's previous gather must have drained
            self.pending[i] = None
        h16_ptr = self.d_h16s[i].data_ptr()
        if self.p2p and exchange:
            if rank == 0:
                h16_ptr = self.peer.slot(i)       # the gathering rank traces straight into its slot of the window
            elif self.pushed_ev[i] is not None:
                stream.wait_event(self.pushed_ev[i])      # the push of step k-2 read this buffer
        if ev:
            ev[0].record(stream)
        rc = lib.rtk_trace_rays_compact_device(self.sc.ptr, self.d_rays.data_ptr(), h16_ptr, n, sh)
        if ev:
            ev[1].record(stream)
        if self.p2p and exchange and rank != 0:
            # the records cross NVLink on the copy engines while k_resolve and the next k_trace run
            self.traced_ev[i].record(stream)
            self.copy_stream.wait_event(self.traced_ev[i])
            self.peer.push(i, h16_ptr, 16 * n, self.copy_stream.cuda_stream)
            self.pushed_ev[i] = torch.cuda.Event()
            self.pushed_ev[i].record(self.copy_stream)
        rc |= lib.rtk_resolve_hits_device(self.sc.ptr, h16_ptr, self.d_hits.data_ptr(), self.d_mask.data_ptr(), n, sh)
        if ev:
            ev[2].record(stream)
        if rc:
            raise RuntimeError(lib.last_error())
        if world > 1 and exchange:
            if not self.p2p:
                self.pending[i] = dist.gather(self.d_h16s[i], self.gather_lists[i], dst=0, async_op=True)
            self.counter += 1

    def drain(self):
        for i in range(2):
            if self.pending[i] is not None:
                self.pending[i].wait()
                self.pending[i] = None
        if self.copy_stream is not None:
            self.copy_stream.synchronize()        # this rank's pushes have landed in rank 0's HBM

    def timed(self, steps, warmup, sampler=None):
        """W warm-up steps, then exactly K steps between barrier + synchronize on both sides; device time,
        max over ranks.  Returns (ms per step, k_trace ms, k_resolve ms, per-rank [k_trace, k_resolve] ms)."""
        torch, dist, world, stream = self.torch, self.dist, self.world, self.stream
        for _ in range(warmup):
            self.step()
        self.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record(stream)
        for i in range(steps):
            self.step(evs[i])
        self.drain()                                  # the last gathers are inside the timed region
        e_end.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = e_start.elapsed_time(e_end)
        trace_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
        resolve_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
        per_rank = None
        if world > 1:
            t = torch.tensor([total_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
            mine = torch.tensor([trace_ms, resolve_ms], device="cuda")
            alls = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(alls, mine)
            per_rank = [[round(float(x), 4) for x in a.tolist()] for a in alls]
        return total_ms / steps, trace_ms, resolve_ms, per_rank

    def last_gathered(self, from_rank, count):
        """what rank 0 holds from `from_rank` after the last step (first `count` records)"""
        last = (self.counter - 1) & 1
        if self.p2p:
            got = self.torch.empty((count, 16), dtype=self.torch.uint8, device="cuda")
            if self.lib.rtk_cuda_peer_push(got.data_ptr(), self.peer.slot(last, from_rank), 16 * count, self.sh) != 0:
                raise RuntimeError(self.lib.last_error())
            self.torch.cuda.synchronize()
            return got
        return self.gather_lists[last][from_rank][:count]

    def close(self):
        if self.peer is not None:
            self.dist.barrier()                       # nobody unmaps the window while rank 0 still reads it
            self.peer.close()
            self.peer = None


# ------------------------------------------------------------------------------------------------
# end to end through the host API
# ------------------------------------------------------------------------------------------------

class Pinned:
    """page-locked host arrays from the library's own allocator (rtk_cuda_host_alloc)"""

    def __init__(self, lib):
        self.lib, self.ptrs = lib, []

    def array(self, n, dtype):
        """an array for batches of n rays: with several devices in use each device's share sits on its NUMA node"""
        dt = np.dtype(dtype)
        nbytes = max(int(n) * dt.itemsize, 16)
        p = self.lib.rtk_cuda_host_alloc_batch(dt.itemsize, int(n))
        if not p:
            raise RuntimeError("rtk_cuda_host_alloc_batch: " + self.lib.last_error())
        self.ptrs.append(p)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))[:int(n) * dt.itemsize].view(dt)

    def free(self):
        for p in self.ptrs:
            self.lib.rtk_cuda_host_free(p)
        self.ptrs = []


def link_ceiling(lib, ndev, mb=256):
    """aggregate GB/s of concurrent pinned-memory copies to / from the first ndev devices of the library's list"""
    out = {"devices": ndev, "bytes_per_device_per_pass": mb << 20}
    for name, d in (("h2d_gbs", 1), ("d2h_gbs", 2), ("both_gbs", 3)):
        g = C.c_double(0)
        out[name] = round(g.value, 2) if lib.rtk_cuda_measure_host_link(ndev, mb << 20, d, 4, C.byref(g)) == 0 else None
    return out


def e2e_legs(lib, api, sc, scene, rays_all, steps, ndev, check16=None):
    """rtk_trace_rays / rtk_trace_rays_compact on host arrays over the library's `ndev` devices (one process).
    Timed with the wall clock around K calls -- the call is synchronous: rays leave host memory and rows are
    back in host memory inside it.  check16: compact records the device path produced for the first rays."""
    n = len(rays_all)
    pin = Pinned(lib)
    out = {}
    try:
        h_rays = pin.array(n, api.RAY_DTYPE)
        h_rays[:] = rays_all
        h_hits = pin.array(n, api.HIT_DTYPE)
        h_mask = pin.array(n, np.uint8)
        h_hits.view(np.uint8)[:] = 0
        h_mask[:] = 0

        def rows():
            got = lib.rtk_trace_rays(sc.ptr, h_rays.ctypes.data, h_hits.ctypes.data, h_mask.ctypes.data, n)
            if got == C.c_size_t(-1).value:
                raise RuntimeError(lib.last_error())
            return got
        rows()
        rows()
        t0 = time.perf_counter()
        for _ in range(steps):
            nh = rows()
        sec = (time.perf_counter() - t0) / steps
        d2h = 68 * int(nh) + n
        k = min(n, 1 << 21)
        same = None
        if check16 is not None:
            kk = min(k, len(check16))
            got16 = api.hits_to_hit16(h_hits[:kk], h_mask[:kk], scene["mesh_first"])
            same = bool(got16.tobytes() == check16[:kk].tobytes())
            k = kk
        ceiling = link_ceiling(lib, ndev)
        moved = (32 * n + d2h) / sec / 1e9
        out["e2e"] = {"value": n / sec / 1e6, "unit": UNIT, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": d2h,
                      "ms_per_step": sec * 1e3, "rays_per_step": n, "devices": ndev, "processes": 1,
                      "api": "rtk_trace_rays (host rtk_ray[] in, rtk_hit[] + mask out; page-locked arrays: the resolve kernel "
                             "writes the rows of the rays that hit straight into the caller's array)",
                      "hits_per_step": int(nh), "rows_equal_device_path": same, "rows_compared": k,
                      "roofline": dict(ceiling, bound="pcie", achieved_gbs=moved,
                                       frac_of_bidirectional=moved / ceiling["both_gbs"] if ceiling.get("both_gbs") else None,
                                       upload_floor_ms=32 * n / (ceiling["h2d_gbs"] * 1e9) * 1e3 if ceiling.get("h2d_gbs") else None,
                                       note="ceiling = concurrent cudaMemcpyAsync between pinned host memory and every device in use, "
                                            "measured in this run; the batch needs 32 B/ray up and ~%.0f B/ray down" % (d2h / n))}
        # compact records
        h_h16 = pin.array(n, api.HIT16_DTYPE)

        def compact():
            if lib.rtk_trace_rays_compact(sc.ptr, h_rays.ctypes.data, h_h16.ctypes.data, n) != 0:
                raise RuntimeError(lib.last_error())
        compact()
        t0 = time.perf_counter()
        for _ in range(steps):
            compact()
        c_s = (time.perf_counter() - t0) / steps
        out["e2e_compact"] = {"value": n / c_s / 1e6, "unit": UNIT, "ms_per_step": c_s * 1e3, "devices": ndev,
                              "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 16 * n,
                              "api": "rtk_trace_rays_compact (host rtk_ray[] in, 16-byte (t,u,v,triangle) record per ray out)",
                              "records_equal_device_path": bool(h_h16[:k].tobytes() == check16[:k].tobytes()) if check16 is not None else None}
        # pageable arrays: staged rows + host placement threads
        m = min(n, FULL_RAYS)
        p_hits = np.zeros(m, dtype=api.HIT_DTYPE)
        p_mask = np.zeros(m, dtype=np.uint8)
        p_rays = np.array(rays_all[:m])
        lib.rtk_trace_rays(sc.ptr, p_rays.ctypes.data, p_hits.ctypes.data, p_mask.ctypes.data, m)
        t0 = time.perf_counter()
        got = lib.rtk_trace_rays(sc.ptr, p_rays.ctypes.data, p_hits.ctypes.data, p_mask.ctypes.data, m)
        p_s = time.perf_counter() - t0
        mm = p_mask.astype(bool)
        out["e2e_pageable"] = {"value": m / p_s / 1e6, "unit": UNIT, "ms_per_step": p_s * 1e3, "rays_per_step": m, "devices": ndev,
                               "api": "rtk_trace_rays with pageable (malloc) arrays: rays through pinned bounce buffers, rows through pinned staging + host placement threads",
                               "rows_equal_pinned_path": bool(got == int(h_mask[:m].sum()) and p_hits[mm].tobytes() == h_hits[:m][mm].tobytes())}
    finally:
        pin.free()
    return out


# ------------------------------------------------------------------------------------------------

def emit(line):
    """the one JSON line goes to the real stdout; everything else a library prints (NCCL banner
    ...) was redirected to stderr at start-up"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def shm_path(rank):
    return "/dev/shm/rtk_b200_bench_%s_%d.npy" % (os.environ.get("MASTER_PORT", "0"), rank)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU (default: the workload's own count)")
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS), help="BASELINE.json config; C3 is the headline")
    ap.add_argument("--scale", type=float, default=1.0, help="triangle-count scale of the terrain (1.0 = 1M)")
    ap.add_argument("--ref-sample", type=int, default=1 << 22)
    ap.add_argument("--cpu-sample", type=int, default=1 << 22)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--build-mode", default="sah", choices=["lbvh", "sah"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-rays", type=int, default=10240, help="rays checked against the CPU oracle per config (0: none)")
    ap.add_argument("--cull", type=int, default=1, help="1 provable dominant-axis culling (default), 0 full-box culling")
    ap.add_argument("--reserve-sms", type=int, default=0, help="experiment: SMs kept out of the traversal grid for the NCCL kernels of "
                                                               "the overlapped gather (measured at 2 GPUs: 0 -> 3131, 4 -> 3066, 8 -> 3003 Mrays/s)")
    ap.add_argument("--gather", default="p2p", choices=["nccl", "p2p"],
                    help="N > 1: how the compact hit records reach rank 0.  p2p (default): copy-engine pushes into a peer-memory window "
                         "on rank 0 (rtk_cuda_peer_*, CUDA IPC over NVLink) -- no kernel beside the persistent traversal grid; measured "
                         "at 8 GPUs: 12702 Mrays/s against 11855 with nccl (torch.distributed gather, whose send/receive kernels cost "
                         "every rank 6 %% of its k_trace and rank 0 10 %%)")
    ap.add_argument("--legs", default="c4,c5,e2e", help="extra legs beside the headline: any of c4, c5, e2e (comma separated) or none")
    ap.add_argument("--c4-rays", type=int, default=C4_RAYS, help="total rays of the C4 leg (split over the ranks)")
    ap.add_argument("--c4-scale", type=float, default=1.0)
    ap.add_argument("--c4-steps", type=int, default=10)
    ap.add_argument("--lib", default=None, help="experiment: alternative build of librtk_b200 (same ABI)")
    ap.add_argument("--wf-frame", default=None, help="C5 dry runs on the emulator only: WIDTHxHEIGHTxSPP instead of 3840x2160x16")
    ap.add_argument("--gen-threads", type=int, default=0, help="host threads generating the synthetic rays (default: CPUs / ranks)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    legs = set(x for x in args.legs.split(",") if x and x != "none")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    gen_threads = args.gen_threads or max(1, min(16, (os.cpu_count() or 8) // max(1, world)))

    cfg_name, default_rays, cfg_text = WORKLOADS[args.workload]
    if args.rays <= 0:
        args.rays = default_rays
    workload = {"workload": cfg_text,
                "triangles": None, "rays_per_gpu": args.rays, "ray_bytes": 32, "hit_bytes": 68,
                "l2_policy": "ray and hit buffers (%.2f GB in, %.2f GB out per step) are far larger than the "
                             "126 MB L2 and stream through it every step; the scene (BVH + triangles) is the "
                             "reused working set" % (32 * args.rays / 1e9, 85 * args.rays / 1e9),
                "parallelism": f"rays sharded over {world} GPU(s), scene replicated"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        scene = scenes.config_scene(cfg_name, args.scale)
        workload["triangles"] = int(len(scene["tris"]))
        rays = gen_rays(scene, min(args.rays, args.ref_sample), 0, args.workload, gen_threads)
        workload["rays_per_gpu"] = args.rays
        workload["gather"] = (args.gather if world > 1 else None)      # the same `config` object as the repo's arm prints
        run_reference(args, workload, scene, rays)
        return 0

    import torch
    import torch.distributed as dist
    from rtk_b200 import api

    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")        # host-side barriers: ranks that wait must not spin on their GPU
    lib = api.load() if not args.lib else api.Library(os.path.abspath(args.lib))
    r = lib.rtk_cuda_init(local_rank)
    if r != 0:
        raise RuntimeError("rtk_cuda_init failed (no CPU fallback): " + lib.last_error())
    lib.rtk_cuda_set_cull_mode(args.cull)
    reserve = max(0, args.reserve_sms)
    if lib.rtk_cuda_reserve_sms(reserve) != 0:
        raise RuntimeError(lib.last_error())
    lib.rtk_cuda_set_build_mode(api.RTK_CUDA_BUILD_SAH if args.build_mode == "sah" else api.RTK_CUDA_BUILD_LBVH)
    hbm_peak, peak_src = measured_peaks()

    scene = scenes.config_scene(cfg_name, args.scale)
    ntris = int(len(scene["tris"]))
    workload["triangles"] = ntris
    if args.workload == "C5":
        run_wavefront(args, workload, lib, api, scene, rank, world, local_rank)
        if world > 1:
            dist.destroy_process_group()
        return 0
    n = args.rays
    rays_np = gen_rays(scene, n, rank, args.workload, gen_threads)
    if world > 1 and "e2e" in legs:
        np.save(shm_path(rank), rays_np.view(np.uint8).reshape(-1, 32))      # rank 0 drives every GPU in the e2e leg

    # ---- build: end to end from host buffers (cold, warm), then device-only rebuilds ------------
    sc, info, build = build_numbers(lib, scene, args.build_mode, hbm_peak)

    workload["gather"] = (args.gather if world > 1 else None)
    T = ShardedTrace(args, lib, torch, dist, sc, rays_np, rank, world, args.gather)
    sh = T.sh

    # ---- parity self-check against the CPU oracle (outside the timed region) -------------------
    parity = None
    if rank == 0 and args.parity_rays > 0:
        T.step(exchange=False)                 # rank-local: no collective outside the common path
        torch.cuda.synchronize()
        kb = min(n, 65536)
        got16 = T.d_h16s[0].cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
        parity = oracle_parity(api, scene, rays_np, got16, args.parity_rays)
        parity.update(gpu_bruteforce_parity(lib, api, torch, sc, T.d_rays, T.d_h16s[0], kb, sh))

    # ---- algorithmic bytes per ray from the counter-instrumented kernel ------------------------
    per_ray, bytes_per_ray = trace_stats(lib, api, sc, T.d_rays, T.d_h16s[0], n, sh)

    # ---- timed region --------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_per_step, trace_ms, resolve_ms, per_rank = T.timed(args.steps, args.warmup, sampler)
    clocks = sampler.stop()
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---- what rank 0 gathered from the LAST rank in the last step is what that rank's rays give here
    # (outside the timed region; the scene is replicated, so rank 0 can trace them itself) ----------
    gather_check = None
    if world > 1 and args.workload in ("C3", "C2"):
        kv = min(n, 1 << 20)
        if rank == 0:
            try:
                theirs = gen_rays(scene, kv, 0, args.workload) if args.workload == "C2" else \
                    scenes.bounce_rays(scene, kv, seed=0xD3, first=(world - 1) * n)
                d_r = torch.from_numpy(theirs.view(np.uint8).reshape(-1, 32)).cuda()
                d_o = torch.zeros((kv, 16), dtype=torch.uint8, device="cuda")
                if lib.rtk_trace_rays_compact_device(sc.ptr, d_r.data_ptr(), d_o.data_ptr(), kv, sh) != 0:
                    raise RuntimeError(lib.last_error())
                torch.cuda.synchronize()
                gather_check = {"records_compared": kv, "from_rank": world - 1, "equal": bool(torch.equal(T.last_gathered(world - 1, kv), d_o))}
            except Exception as ex:          # a failed self-check must not cost the measurement
                gather_check = {"error": str(ex)}
    T.close()

    # ---- occlusion (any-hit) query on the same rays: SURVEY 8(f) N3, reported beside the headline --
    occ = None
    if world == 1:
        d_occ = torch.zeros((n,), dtype=torch.uint8, device="cuda")
        eo = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(7):
            if i == 2:
                eo[0].record(T.stream)
            assert lib.rtk_occluded_rays_device(sc.ptr, T.d_rays.data_ptr(), d_occ.data_ptr(), n, sh) == 0, lib.last_error()
        eo[1].record(T.stream)
        torch.cuda.synchronize()
        occ_ms = eo[0].elapsed_time(eo[1]) / 5
        occ = {"value": n / (occ_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": occ_ms,
               "agrees_with_closest_hit_mask": bool(torch.equal(d_occ != 0, T.d_mask != 0))}
        del d_occ

    roofline = None
    if rank == 0:
        roofline = roofline_block(lib, args.workload, info, per_ray, bytes_per_ray, n, trace_ms, hbm_peak, peak_src,
                                  probe_bytes=(64 << 20) if info.device_bytes < 120e6 else (1 << 30))

    # ---- end to end through the host API ----------------------------------------------------------
    # N = 1: this process, its one GPU.  N > 1: ONE process (rank 0) drives all N GPUs through the library's own
    # multi-device path (rtk_cuda_init_devices): the other ranks free their GPUs and wait on a host-side barrier.
    e2e_out = {}
    check16 = None
    if rank == 0:
        T.step(exchange=False)               # this rank's own records in buffer 0, whatever the gather mode left there
        torch.cuda.synchronize()
        check16 = T.d_h16s[0][:min(n, 1 << 21)].cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
    del T
    if "e2e" in legs:
        if world == 1:
            e2e_out = e2e_legs(lib, api, sc, scene, rays_np, args.e2e_steps, 1, check16)
        else:
            sc.free()
            sc = None
            torch.cuda.empty_cache()
            dist.barrier(group=cpu_group)
            if rank == 0:
                try:
                    lib.rtk_cuda_shutdown()
                    devs = (C.c_int * world)(*range(world))
                    if lib.rtk_cuda_init_devices(devs, world) != 0:
                        raise RuntimeError(lib.last_error())
                    lib.rtk_cuda_set_build_mode(api.RTK_CUDA_BUILD_SAH if args.build_mode == "sah" else api.RTK_CUDA_BUILD_LBVH)
                    parts = [rays_np] + [np.load(shm_path(r)).view(api.RAY_DTYPE).reshape(-1) for r in range(1, world)]
                    rays_all = np.concatenate(parts)
                    del parts
                    lib.build_scene(scene["meshes"]).free()                    # contexts, modules and peer mappings of all devices exist after this
                    t0 = time.perf_counter()
                    sc_all = lib.build_scene(scene["meshes"])                  # builds on device 0, replicates to the others
                    rep_ms = (time.perf_counter() - t0) * 1e3
                    e2e_out = e2e_legs(lib, api, sc_all, scene, rays_all, args.e2e_steps, world, check16)
                    e2e_out["e2e"]["build_and_replicate_ms"] = rep_ms
                    e2e_out["e2e"]["single_link"] = link_ceiling(lib, 1)
                    sc_all.free()
                    del rays_all
                    lib.rtk_cuda_shutdown()
                    if lib.rtk_cuda_init(local_rank) != 0:
                        raise RuntimeError(lib.last_error())
                except Exception as ex:
                    e2e_out = {"e2e": {"error": str(ex)}}
            dist.barrier(group=cpu_group)
            try:
                os.remove(shm_path(rank))
            except OSError:
                pass
    if sc is not None:
        sc.free()
    del rays_np

    # ---- C4: the 10M-triangle scene, one FIXED batch of mixed rays split over the ranks (strong scaling) ----
    c4 = None
    c4_scene, c4_sc = None, None
    if "c4" in legs or ("c5" in legs and world == 1):
        c4_scene = scenes.config_scene("C4", args.c4_scale)
    if "c4" in legs:
        total = args.c4_rays
        blk = 65536
        per = (total + world - 1) // world
        if per % blk == 0 and per * world == total:
            # block-cyclic: rank r traces the 64Ki-ray blocks b = r (mod N).  Contiguous ranges give the ranks different
            # parts of the camera image (measured at 8 GPUs: k_trace 3.95 .. 5.52 ms per rank, the step waits for the slowest)
            mine = list(range(rank, total // blk, world))
            rays4 = np.empty(per, dtype=api.RAY_DTYPE)

            def one(i):
                rays4[i * blk:(i + 1) * blk] = scenes.mixed_rays(c4_scene, total, seed=0xD4, first=mine[i] * blk, count=blk)
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(gen_threads) as ex:
                list(ex.map(one, range(len(mine))))
        else:
            # odd totals (scaled-down runs): every rank makes the whole set and keeps its part, padded to equal length
            allr = scenes.mixed_rays(c4_scene, total, seed=0xD4, threads=gen_threads)
            rays4 = np.ascontiguousarray(np.resize(allr[rank * per:min(total, (rank + 1) * per)] if rank * per < total else allr[:1], per))
            del allr
        c4_sc, info4, build4 = build_numbers(lib, c4_scene, args.build_mode, hbm_peak)
        T4 = ShardedTrace(args, lib, torch, dist, c4_sc, rays4, rank, world, args.gather)
        parity4 = None
        if rank == 0 and args.parity_rays > 0:
            T4.step(exchange=False)
            torch.cuda.synchronize()
            got16 = T4.d_h16s[0].cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
            parity4 = oracle_parity(api, c4_scene, rays4, got16, args.parity_rays, flat_rays=256)
            parity4.update(gpu_bruteforce_parity(lib, api, torch, c4_sc, T4.d_rays, T4.d_h16s[0], min(per, 65536), T4.sh))
        per_ray4, bpr4 = trace_stats(lib, api, c4_sc, T4.d_rays, T4.d_h16s[0], per, T4.sh)
        ms4, tr4, rs4, per_rank4 = T4.timed(args.c4_steps, 3)
        T4.close()
        if rank == 0:
            c4 = {"metric": "closest-hit Mrays/s (mixed rays, 10M triangles)", "value": total / (ms4 * 1e-3) / 1e6, "unit": UNIT,
                  "scaling": "strong", "n_gpus": world, "rays_total": total, "rays_per_gpu": per, "steps": args.c4_steps, "warmup": 3,
                  "ms_per_step": ms4, "kernels_ms": {"k_trace": tr4, "k_resolve": rs4}, "kernels_ms_per_rank": per_rank4,
                  "workload": WORKLOADS["C4"][2].replace("per GPU", "in total, the 64Ki-ray blocks dealt out to the ranks round-robin"),
                  "triangles": int(len(c4_scene["tris"])), "gather": workload["gather"],
                  "parity": parity4, "build": build4,
                  "roofline": roofline_block(lib, "C4", info4, per_ray4, bpr4, per, tr4, hbm_peak, peak_src, probe_bytes=1 << 30)}
        del T4, rays4

    # ---- C5 (N = 1): one band of the wavefront, rays generated on the device ---------------------
    c5 = None
    if "c5" in legs and world == 1:
        if c4_sc is None:
            c4_sc = lib.build_scene(c4_scene["meshes"])
        rps, ms5, counts5, parity5, launches5 = wavefront_leg(args, lib, api, c4_scene, c4_sc, rank, world, 3, 1)
        c5 = {"metric": "closest-hit Mrays/s (wavefront, device-generated rays)", "value": rps / (ms5 * 1e-3) / 1e6, "unit": UNIT,
              "ms_per_step": ms5, "rays_per_step": rps, "steps": 3, "warmup": 1, "workload": WORKLOADS["C5"][2],
              "includes": "generation of the primary and bounce rays (16 + 3 launches) beside the 4 traversals of a step",
              "wavefront": counts5, "parity": parity5}
    if c4_sc is not None:
        c4_sc.free()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload,
        "e2e": e2e_out.get("e2e"), "e2e_compact": e2e_out.get("e2e_compact"), "e2e_pageable": e2e_out.get("e2e_pageable"),
        "gpu_launches": 2 * args.steps,
        "kernels_ms": {"k_trace": trace_ms, "k_resolve": resolve_ms},
        "roofline": roofline, "build": build,
        "occlusion": occ, "parity": parity, "clocks": clocks,
    }
    if per_rank is not None:
        line["kernels_ms_per_rank"] = per_rank
    if gather_check is not None:
        line["gather_check"] = gather_check
    if c4 is not None:
        line["c4"] = c4
    if c5 is not None:
        line["c5"] = c5
    if world == 1 and not args.no_cpu_baseline:
        try:
            scene3 = scene
            rays3 = gen_rays(scene3, min(n, args.cpu_sample), 0, args.workload, gen_threads)
            cb = cpu_reference_run(scene3, rays3, min(n, args.cpu_sample), 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_build_baseline"] = cb["build"]
        except Exception as ex:  # the checker is missing: say so instead of inventing a number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
