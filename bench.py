#!/usr/bin/env python
"""bench.py -- closest-hit Mrays/s (incoherent) on the BASELINE.json workload, plus BVH build Mtris/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rays R] [--scale S]

A step is one pass of the hot path over one batch of synthetic rays: the traversal kernel
(k_trace) followed by the hit-expansion kernel (k_resolve), inputs already resident in HBM.
N=1 workload: BASELINE.json configs[2] -- 1M-triangle procedural terrain, 16 777 216 incoherent
diffuse-bounce rays (the configuration the headline "incoherent Mrays/s" metric is quoted on).
N>1 (torchrun): scene replicated per GPU, every rank traces its own 16 777 216 rays (weak
scaling), compact hit records are gathered on rank 0 inside the timed region -- with NCCL
(default) or, with --gather p2p, by copy-engine pushes into a peer-memory window on rank 0.

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


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


WORKLOADS = {
    # name: (scene config, default rays per GPU, description)
    "C3": ("C3", FULL_RAYS, "C3: 1M-triangle procedural terrain (1001x501 value-noise heightfield), "
                            "16777216 incoherent diffuse-bounce rays per GPU"),
    "C2": ("C2", 2_073_600, "C2: 1M-triangle random soup, 1920x1080 coherent primary rays per GPU"),
    "C4": ("C4", 67_108_864, "C4: 10M-triangle procedural terrain in 2 meshes (2501x2001 vertices), 67108864 mixed rays "
                             "per GPU (thirds of coherent primary / diffuse bounce / short segments, 64Ki blocks)"),
    # wavefront path tracing: rays are generated on the device and never leave it (run_wavefront)
    "C5": ("C4", 66_355_200, "C5: 4K frame (3840x2160) x 16 spp x 4 bounces on the 10M-triangle terrain, pixel rows "
                             "split in 8 bands; one band (3840x270 pixels, 66355200 rays per step) per GPU"),
}


def gen_rays(scene, n, rank=0, workload="C3"):
    if workload == "C2":
        r = scenes.soup_primary_rays()
        return np.ascontiguousarray(np.resize(r, n))
    if workload == "C4":
        return scenes.mixed_rays(scene, n, seed=0xD4 + 0x1000 * rank)
    out = np.empty(n, dtype=scenes.RAY_DTYPE)
    step = 1 << 21
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        out[lo:hi] = scenes.bounce_rays(scene, hi - lo, seed=0xD3, first=rank * n + lo)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def mark(self):
        """samples from here on belong to the timed region"""
        self.first = len(self.rows)

    def start(self):
        self.first = 0
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        rows = self.rows[self.first:] if len(self.rows) - self.first >= 3 else self.rows
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
# C5: wavefront path tracing, rays generated on the device
# ------------------------------------------------------------------------------------------------

def terrain_camera(api, scene, width, height):
    """the camera of scenes.terrain_primary_rays as an rtk_cuda_camera"""
    tris = scene["tris"].reshape(-1, 3)
    lo, hi = tris.min(0), tris.max(0)
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


def run_wavefront(args, workload, lib, api, scene, rank, world, local_rank):
    """A step = one band of the 4K frame: 16 jittered primary rays per pixel generated on the device,
    then 4 x (k_trace, k_gen_bounce with relaunch of the paths that left the scene).  Rays and hits
    never leave HBM; for N > 1 the last bounce's compact hit records are gathered on rank 0."""
    import torch
    import torch.distributed as dist
    W, H, SPP, BOUNCES, BANDS = 3840, 2160, 16, 4, 8
    if args.wf_frame:                                       # dry runs only: the workload is the 4K frame
        W, H, SPP = (int(v) for v in args.wf_frame.split("x"))
    band = (rank % BANDS) if world > 1 else 0
    rows = H // BANDS
    npx = W * rows
    n = npx * SPP
    sc = lib.build_scene(scene["meshes"])
    assert lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0, lib.last_error()     # device time of a warm build
    info = sc.info()
    cam = terrain_camera(api, scene, W, H)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    bufs = [torch.empty((n, 32), dtype=torch.uint8, device="cuda") for _ in range(2)]
    d_h16 = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    d_alive = torch.empty((n,), dtype=torch.uint8, device="cuda")
    gather_list = [torch.empty_like(d_h16) for _ in range(world)] if (world > 1 and rank == 0) else None
    seed = 0xD5
    counts = {}

    def step(ev=None):
        if ev:
            ev[0].record(stream)
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
        if ev:
            ev[1].record(stream)
        if world > 1:
            dist.gather(d_h16, gather_list, dst=0)
    step()
    torch.cuda.synchronize()
    counts["last_bounce_hit_fraction"] = float((d_h16.view(torch.int32)[:, 3] != -1).float().mean().item())
    counts["relaunched_fraction_bounce3"] = float((d_alive == 2).float().mean().item())

    # parity: the last bounce's rays (as the device generated them) against the exhaustive GPU kernel
    # and, for a few, against the CPU oracle
    parity = None
    if rank == 0 and args.parity_rays > 0:
        from oracle import orc
        last = bufs[(BOUNCES - 1) & 1]
        kb = min(n, 65536)
        d_b = torch.zeros((kb, 16), dtype=torch.uint8, device="cuda")
        assert lib.rtk_trace_rays_bruteforce_device(sc.ptr, last.data_ptr(), d_b.data_ptr(), kb, sh) == 0
        torch.cuda.synchronize()
        a = d_h16[:kb].cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
        bb = d_b.cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
        k = min(args.parity_rays, 512)
        rays_np = last[:k].cpu().numpy().view(api.RAY_DTYPE).reshape(-1)
        want = orc.trace_brute(scene["tris"], rays_np)
        parity = {"rays_checked": k, "index_mismatches": int((a[:k]["prim"] != want["prim"]).sum()),
                  "bit_exact": bool(a[:k].tobytes() == want.tobytes()), "against": "oracle (CPU brute force) on device-generated bounce-3 rays",
                  "gpu_bruteforce_rays": kb, "gpu_bruteforce_bit_exact": bool(a.tobytes() == bb.tobytes())}

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(stream)
    for i in range(args.steps):
        step(evs[i])
    e_end.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    total_ms = e_start.elapsed_time(e_end)
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    rays_per_step = n * BOUNCES
    if rank != 0:
        return
    workload = dict(workload)
    workload.update({"triangles": int(len(scene["tris"])), "rays_per_gpu": rays_per_step,
                     "l2_policy": "ray, hit and alive buffers (%.1f GB per bounce) stream through the 126 MB L2; the 10M-triangle "
                                  "scene (%.2f GB) does not fit it" % (n * 81 / 1e9, info.device_bytes / 1e9)})
    line = {"metric": "closest-hit Mrays/s (wavefront, device-generated rays)", "value": world * rays_per_step / (ms_per_step * 1e-3) / 1e6,
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload, "gpu_launches": (SPP + 2 * BOUNCES - 1) * args.steps,
            "e2e": None, "wavefront": counts, "parity": parity, "clocks": clocks,
            "build": {"metric": "BVH build Mtris/s", "value": len(scene["tris"]) / (info.build_device_ms * 1e-3) / 1e6,
                      "unit": "Mtris/s", "device_ms": info.build_device_ms}}
    emit(line)
    sc.free()

# ------------------------------------------------------------------------------------------------

def emit(line):
    """the one JSON line goes to the real stdout; everything else a library prints (NCCL banner
    ...) was redirected to stderr at start-up"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


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
    ap.add_argument("--parity-rays", type=int, default=1024)
    ap.add_argument("--cull", type=int, default=1, help="1 provable dominant-axis culling (default), 0 full-box culling")
    ap.add_argument("--presort", action="store_true", help="experiment: order the rays by origin cell + direction octant on the host first")
    ap.add_argument("--reserve-sms", type=int, default=0, help="experiment: SMs kept out of the traversal grid for the NCCL kernels of "
                                                               "the overlapped gather (measured at 2 GPUs: 0 -> 3131, 4 -> 3066, 8 -> 3003 Mrays/s)")
    ap.add_argument("--gather", default="nccl", choices=["nccl", "p2p"],
                    help="N > 1: how the compact hit records reach rank 0.  nccl (default, measured): torch.distributed gather, "
                         "overlapped with the next step.  p2p: copy-engine pushes into a peer-memory window on rank 0 "
                         "(rtk_cuda_peer_*, CUDA IPC over NVLink; no NCCL kernels beside the persistent traversal grid)")
    ap.add_argument("--lib", default=None, help="experiment: alternative build of librtk_b200 (same ABI)")
    ap.add_argument("--wf-frame", default=None, help="C5 dry runs on the emulator only: WIDTHxHEIGHTxSPP instead of 3840x2160x16")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

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
        rays = gen_rays(scene, min(args.rays, args.ref_sample), 0, args.workload)
        workload["rays_per_gpu"] = args.rays
        run_reference(args, workload, scene, rays)
        return 0

    import torch
    import torch.distributed as dist
    from rtk_b200 import api

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = api.load() if not args.lib else api.Library(os.path.abspath(args.lib))
    r = lib.rtk_cuda_init(local_rank)
    if r != 0:
        raise RuntimeError("rtk_cuda_init failed (no CPU fallback): " + lib.last_error())
    lib.rtk_cuda_set_cull_mode(args.cull)
    reserve = max(0, args.reserve_sms)
    if lib.rtk_cuda_reserve_sms(reserve) != 0:
        raise RuntimeError(lib.last_error())
    workload["reserved_sms"] = reserve
    lib.rtk_cuda_set_build_mode(api.RTK_CUDA_BUILD_SAH if args.build_mode == "sah" else api.RTK_CUDA_BUILD_LBVH)

    scene = scenes.config_scene(cfg_name, args.scale)
    ntris = int(len(scene["tris"]))
    workload["triangles"] = ntris
    if args.workload == "C5":
        run_wavefront(args, workload, lib, api, scene, rank, world, local_rank)
        if world > 1:
            dist.destroy_process_group()
        return 0
    n = args.rays
    rays_np = gen_rays(scene, n, rank, args.workload)
    if args.presort:
        lo = scene["tris"].reshape(-1, 3).min(0)
        ext = scene["tris"].reshape(-1, 3).max(0) - lo
        q = np.clip(((rays_np["o"] - lo) / ext * 1023).astype(np.int64), 0, 1023)

        def spread(v):
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            v = (v | (v << 2)) & 0x09249249
            return v
        key = (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])
        octant = ((rays_np["d"][:, 0] < 0).astype(np.int64) | ((rays_np["d"][:, 1] < 0).astype(np.int64) << 1) |
                  ((rays_np["d"][:, 2] < 0).astype(np.int64) << 2))
        rays_np = np.ascontiguousarray(rays_np[np.argsort((key << 3) | octant, kind="stable")])

    # ---- build: end to end from host buffers, then device-only rebuilds ------------------------
    t0 = time.perf_counter()
    sc = lib.build_scene(scene["meshes"])
    build_e2e_s = time.perf_counter() - t0
    build_ms = []
    for i in range(5):
        assert lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0, lib.last_error()
        build_ms.append(sc.info().build_device_ms)
    info = sc.info()
    build_dev_ms = float(np.median(build_ms[1:]))

    # ---- device buffers ------------------------------------------------------------------------
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    d_rays = torch.from_numpy(rays_np.view(np.uint8).reshape(-1, 32)).cuda()
    # compact hits are double buffered so that, for N > 1, the NCCL gather of step k runs on the
    # communication stream while step k+1 traces into the other buffer
    d_h16s = [torch.zeros((n, 16), dtype=torch.uint8, device="cuda") for _ in range(2 if world > 1 else 1)]
    d_h16 = d_h16s[0]
    d_hits = torch.zeros((n, 68), dtype=torch.uint8, device="cuda")
    d_mask = torch.zeros((n,), dtype=torch.uint8, device="cuda")
    gather_lists = [None, None]
    p2p = world > 1 and args.gather == "p2p"
    workload["gather"] = (args.gather if world > 1 else None)
    peer, copy_stream = None, None
    if p2p:
        from rtk_b200 import shard

        def exchange(handle):
            box = [handle]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        peer = shard.PeerGather(lib, rank, world, 16 * n, 2, exchange)
        copy_stream = torch.cuda.Stream()
        traced_ev = [torch.cuda.Event() for _ in range(2)]
        pushed_ev = [None, None]
    elif world > 1 and rank == 0:
        gather_lists = [[torch.empty_like(d_h16) for _ in range(world)] for _ in range(2)]
    pending = [None, None]
    counter = [0]

    def step(ev=None, exchange=True):
        i = counter[0] & 1 if (world > 1 and exchange) else 0
        if pending[i] is not None:
            pending[i].wait()                # the buffer's previous gather must have drained
            pending[i] = None
        h16_ptr = d_h16s[i].data_ptr()
        if p2p and exchange:
            if rank == 0:
                h16_ptr = peer.slot(i)       # the gathering rank traces straight into its slot of the window
            elif pushed_ev[i] is not None:
                stream.wait_event(pushed_ev[i])      # the push of step k-2 read this buffer
        if ev:
            ev[0].record(stream)
        rc = lib.rtk_trace_rays_compact_device(sc.ptr, d_rays.data_ptr(), h16_ptr, n, sh)
        if ev:
            ev[1].record(stream)
        if p2p and exchange and rank != 0:
            # the records cross NVLink on the copy engines while k_resolve and the next k_trace run
            traced_ev[i].record(stream)
            copy_stream.wait_event(traced_ev[i])
            peer.push(i, h16_ptr, 16 * n, copy_stream.cuda_stream)
            pushed_ev[i] = torch.cuda.Event()
            pushed_ev[i].record(copy_stream)
        rc |= lib.rtk_resolve_hits_device(sc.ptr, h16_ptr, d_hits.data_ptr(), d_mask.data_ptr(), n, sh)
        if ev:
            ev[2].record(stream)
        if rc:
            raise RuntimeError(lib.last_error())
        if world > 1 and exchange:
            if not p2p:
                pending[i] = dist.gather(d_h16s[i], gather_lists[i], dst=0, async_op=True)
            counter[0] += 1

    def drain():
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None
        if copy_stream is not None:
            copy_stream.synchronize()        # this rank's pushes have landed in rank 0's HBM

    # ---- parity self-check against the CPU oracle (outside the timed region) -------------------
    parity = None
    if rank == 0 and args.parity_rays > 0:
        from oracle import orc
        step(exchange=False)                 # rank-local: no collective outside the common path
        torch.cuda.synchronize()
        k = min(args.parity_rays, n, 2048)       # CPU brute force: 1M triangles per ray
        got = d_h16[:k].cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
        want = orc.trace_brute(scene["tris"], rays_np[:k])
        parity = {"rays_checked": k, "index_mismatches": int((got["prim"] != want["prim"]).sum()),
                  "bit_exact": bool(got.tobytes() == want.tobytes()), "against": "oracle (CPU brute force)"}
        if args.parity_rays >= 65536:
            # wide check against the exhaustive GPU kernel (same arithmetic, no BVH)
            kb = min(n, args.parity_rays)
            d_b = torch.zeros((kb, 16), dtype=torch.uint8, device="cuda")
            assert lib.rtk_trace_rays_bruteforce_device(sc.ptr, d_rays.data_ptr(), d_b.data_ptr(), kb, sh) == 0
            torch.cuda.synchronize()
            a = d_h16[:kb].cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
            b = d_b.cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
            parity["gpu_bruteforce_rays"] = kb
            parity["gpu_bruteforce_index_mismatches"] = int((a["prim"] != b["prim"]).sum())
            parity["gpu_bruteforce_bit_exact"] = bool(a.tobytes() == b.tobytes())

    # ---- algorithmic bytes per ray from the counter-instrumented kernel ------------------------
    st = api.rtk_cuda_trace_stats()
    ns = min(n, 1 << 20)
    assert lib.rtk_trace_stats_device(sc.ptr, d_rays.data_ptr(), d_h16.data_ptr(), ns, C.byref(st), sh) == 0, lib.last_error()
    nodes_per_ray = st.node_visits / ns
    tris_per_ray = st.tri_tests / ns
    leaves_per_ray = st.leaf_visits / ns
    hit_frac = st.hits / ns
    bytes_per_ray = 32 + 16 + 256 * nodes_per_ray + 48 * tris_per_ray

    # ---- timed region --------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(stream)
    for i in range(args.steps):
        step(evs[i])
    drain()                                  # the last gathers are inside the timed region
    e_end.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    total_ms = e_start.elapsed_time(e_end)
    trace_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    resolve_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---- what rank 0 gathered from the LAST rank in the last step is what that rank's rays give here
    # (outside the timed region; the scene is replicated, so rank 0 can trace them itself) ----------
    gather_check = None
    if world > 1 and args.workload in ("C3", "C2") and not args.presort:
        kv = min(n, 1 << 20)
        last = (counter[0] - 1) & 1
        if rank == 0:
            try:
                theirs = gen_rays(scene, kv, 0, args.workload) if args.workload == "C2" else \
                    scenes.bounce_rays(scene, kv, seed=0xD3, first=(world - 1) * n)
                d_r = torch.from_numpy(theirs.view(np.uint8).reshape(-1, 32)).cuda()
                d_o = torch.zeros((kv, 16), dtype=torch.uint8, device="cuda")
                if lib.rtk_trace_rays_compact_device(sc.ptr, d_r.data_ptr(), d_o.data_ptr(), kv, sh) != 0:
                    raise RuntimeError(lib.last_error())
                torch.cuda.synchronize()
                if p2p:
                    got = torch.empty((kv, 16), dtype=torch.uint8, device="cuda")
                    if lib.rtk_cuda_peer_push(got.data_ptr(), peer.slot(last, world - 1), 16 * kv, sh) != 0:
                        raise RuntimeError(lib.last_error())
                    torch.cuda.synchronize()
                else:
                    got = gather_lists[last][world - 1][:kv]
                gather_check = {"records_compared": kv, "from_rank": world - 1, "equal": bool(torch.equal(got, d_o))}
            except Exception as ex:          # a failed self-check must not cost the measurement
                gather_check = {"error": str(ex)}
    if peer is not None:
        dist.barrier()                       # nobody unmaps the window while rank 0 still reads it
        peer.close()

    # ---- occlusion (any-hit) query on the same rays: SURVEY 8(f) N3, reported beside the headline --
    occ = None
    if world == 1:
        d_occ = torch.zeros((n,), dtype=torch.uint8, device="cuda")
        eo = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(7):
            if i == 2:
                eo[0].record(stream)
            assert lib.rtk_occluded_rays_device(sc.ptr, d_rays.data_ptr(), d_occ.data_ptr(), n, sh) == 0, lib.last_error()
        eo[1].record(stream)
        torch.cuda.synchronize()
        occ_ms = eo[0].elapsed_time(eo[1]) / 5
        occ = {"value": n / (occ_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": occ_ms,
               "agrees_with_closest_hit_mask": bool(torch.equal(d_occ != 0, d_mask != 0))}

    # ---- end to end through the host API (pinned host buffers, H2D + D2H inside) ---------------
    e2e = None
    h_rays = torch.from_numpy(rays_np.view(np.uint8).reshape(-1, 32)).pin_memory()
    h_hits = torch.empty((n, 68), dtype=torch.uint8).pin_memory()
    h_mask = torch.empty((n,), dtype=torch.uint8).pin_memory()

    def e2e_step():
        got = lib.rtk_trace_rays(sc.ptr, h_rays.data_ptr(), h_hits.data_ptr(), h_mask.data_ptr(), n)
        if got == C.c_size_t(-1).value:
            raise RuntimeError(lib.last_error())
        return got
    e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        nh = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    # what crosses PCIe per step: rays up; per ray one mask byte, per 128 rays a 4-byte block base,
    # and the 68-byte rows of the rays that hit (rows of misses stay untouched, rtk.c:571-576)
    d2h = 68 * int(nh) + n + 4 * ((n + 127) // 128) + 16 * ((n + (1 << 20) - 1) >> 20)
    k = min(n, 1 << 21)
    hm = h_mask[:k].numpy().astype(bool)
    same_rows = bool(np.array_equal(h_hits[:k].numpy()[hm], d_hits[:k].cpu().numpy()[hm]) and
                     np.array_equal(hm, d_mask[:k].cpu().numpy().astype(bool)))
    e2e = {"value": world * n / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_s * 1e3, "api": "rtk_trace_rays (host rtk_ray[] in, rtk_hit[] + mask out, pinned)",
           "hits_per_step": int(nh), "rows_equal_device_path": same_rows, "rows_compared": k}

    # ---- the same batch with compact results (rtk_trace_rays_compact: 16-byte records, no host-side
    # row placement), reported beside the headline e2e; single GPU only, and never at the measurement's cost
    e2e_compact = None
    if world == 1:
        try:
            h_h16 = torch.empty((n, 16), dtype=torch.uint8).pin_memory()

            def compact_step():
                if lib.rtk_trace_rays_compact(sc.ptr, h_rays.data_ptr(), h_h16.data_ptr(), n) != 0:
                    raise RuntimeError(lib.last_error())
            compact_step()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                compact_step()
            torch.cuda.synchronize()
            c_s = (time.perf_counter() - t0) / args.e2e_steps
            e2e_compact = {"value": n / c_s / 1e6, "unit": UNIT, "ms_per_step": c_s * 1e3,
                           "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 16 * n,
                           "api": "rtk_trace_rays_compact (host rtk_ray[] in, 16-byte (t,u,v,triangle) record per ray out, pinned)",
                           "records_equal_device_path": bool(torch.equal(h_h16[:k], d_h16[:k].cpu()))}
        except Exception as ex:
            e2e_compact = {"error": str(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    achieved = bytes_per_ray * n / (trace_ms * 1e-3) / 1e9
    # the scene of the 1M-triangle workloads lives in L2: measure the L2 (and HBM) read bandwidth of
    # this very GPU with the library's probe kernel
    l2_gbs, hbm_read_gbs = C.c_double(0), C.c_double(0)
    lib.rtk_cuda_measure_read_bandwidth(64 << 20, 40, C.byref(l2_gbs))
    lib.rtk_cuda_measure_read_bandwidth(4 << 30, 3, C.byref(hbm_read_gbs))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload,
        "e2e": e2e, "e2e_compact": e2e_compact, "gpu_launches": 2 * args.steps,
        "kernels_ms": {"k_trace": trace_ms, "k_resolve": resolve_ms},
        "roofline": {"bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one k_trace launch on this very
                     # workload (profiles/r1e_k_trace_raw.csv); unknown for any other workload
                     "traffic": 1.728e9 if (args.workload == "C3" and n == FULL_RAYS and args.scale == 1.0 and args.build_mode == "sah") else None,
                     "traffic_unit": "bytes per launch (ncu, profiles/r1e_k_trace_raw.csv; captured before the collapse rule "
                                     "that absorbs small subtrees, i.e. on a tree with a third more wide nodes)",
                     "peak_source": peak_src,
                     "l2": {"achieved": achieved, "peak": l2_gbs.value, "unit": "GB/s",
                            "frac": achieved / l2_gbs.value if l2_gbs.value else None,
                            "peak_source": "measured in this run: 16-byte loads over a 64 MiB buffer, 40 passes "
                                           "(rtk_cuda_measure_read_bandwidth); the same probe over 4 GiB reads "
                                           "%.0f GB/s from HBM" % hbm_read_gbs.value},
                     "bytes_per_ray": bytes_per_ray,
                     "per_ray": {"wide_node_visits": nodes_per_ray, "leaf_visits": leaves_per_ray,
                                 "triangle_tests": tris_per_ray, "hit_fraction": hit_frac},
                     "note": "algorithmic bytes = 32 (ray) + 16 (compact hit) + 256 per wide-node visit + 48 per "
                             "triangle tested, counted by the instrumented kernel on the first 2^20 rays; the scene "
                             "is %.0f MB against a 126 MB L2, so DRAM traffic is %s this"
                             % (info.device_bytes / 1e6, "far below" if info.device_bytes < 120e6 else "a fraction of")},
        "build": {"metric": "BVH build Mtris/s", "value": ntris / (build_dev_ms * 1e-3) / 1e6, "unit": "Mtris/s",
                  "device_ms": build_dev_ms, "mode": args.build_mode,
                  "e2e": {"value": ntris / build_e2e_s / 1e6, "unit": "Mtris/s", "ms": build_e2e_s * 1e3,
                          "api": "rtk_build_scene (host mesh buffers in, first call, includes CUDA context warm-up)"},
                  "wide_nodes": int(info.num_wide_nodes), "leaves": int(info.num_leaves), "depth": int(info.wide_depth),
                  "sah_cost": info.sah_cost, "scene_bytes": int(info.device_bytes)},
        "occlusion": occ, "parity": parity, "clocks": clocks,
    }
    if gather_check is not None:
        line["gather_check"] = gather_check
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_reference_run(scene, rays_np, min(n, args.cpu_sample), 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_build_baseline"] = cb["build"]
        except Exception as ex:  # the checker is missing: say so instead of inventing a number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
    emit(line)
    sc.free()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
