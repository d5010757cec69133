#!/usr/bin/env python
"""Dry run of bench.py's CONTROL FLOW on a machine without a GPU.

bench.py can only be measured on a B200, but its branches (parity self-check, statistics, timed
loop, overlapped gather, gather check, occlusion leg, host-API legs, JSON assembly; N = 1 and N = 2)
can be executed here: the library is the emulator build (tests/emu, device memory = host memory),
torch's CUDA entry points are replaced by host stand-ins, and NCCL by gloo.  Nothing printed by a dry
run is a measurement -- the line is checked for shape only.  TEST INFRASTRUCTURE.

    python tools/bench_emu.py [world] [bench.py arguments...]
    python tools/bench_emu.py 2                   # peer-memory gather (the default): emulated device memory is
                                                  # POSIX shared memory, the window really is mapped by the other process
    python tools/bench_emu.py 2 --gather nccl     # torch.distributed gather (gloo here)
"""
import json
import os
import socket
import subprocess
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def install_shims():
    import torch
    import torch.distributed as dist

    def strip(fn):
        def wrapped(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k.pop("device")
            return fn(*a, **k)
        return wrapped
    for name in ("zeros", "empty", "full", "tensor", "ones"):
        setattr(torch, name, strip(getattr(torch, name)))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.pin_memory = lambda self, *a, **k: self

    class Event:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return max((other.t - self.t) * 1e3, 1e-6)

        def synchronize(self):
            pass

    class Stream:
        cuda_stream = 0

        def wait_event(self, ev):
            pass

        def synchronize(self):
            pass
    torch.cuda.Event, torch.cuda.Stream = Event, Stream
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: Stream()
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **k: real_init("gloo", **{kk: v for kk, v in k.items() if kk != "device_id"})

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    from rtk_b200 import api
    real_lib_init = api.Library.__init__

    def lib_init(self, path):
        real_lib_init(self, path)
        probe, gprobe, lprobe = self.rtk_cuda_measure_read_bandwidth, self.rtk_cuda_measure_gather_bandwidth, self.rtk_cuda_measure_host_link
        # the emulator streams a buffer with fibers: keep the probes tiny
        self.rtk_cuda_measure_read_bandwidth = lambda nbytes, passes, out: probe(1 << 16, 1, out)
        self.rtk_cuda_measure_gather_bandwidth = lambda nbytes, rec, passes, out: gprobe(1 << 16, rec, 1, out)
        self.rtk_cuda_measure_host_link = lambda ndev, nbytes, d, passes, out: lprobe(ndev, 1 << 16, d, 1, out)
    api.Library.__init__ = lib_init


def child():
    install_shims()
    import runpy
    sys.argv = ["bench.py"] + sys.argv[3:]
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child()
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    extra = sys.argv[2:]
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    lib = build_emu.build()
    args = ["--gpus", str(world), "--steps", "2", "--warmup", "1", "--rays", "4096", "--scale", "0.004",
            "--lib", lib, "--e2e-steps", "1", "--parity-rays", "256", "--cpu-sample", "4096",
            "--legs", "c4,e2e", "--c4-rays", "6000", "--c4-scale", "0.0004", "--c4-steps", "1"] + extra
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(world):
        env = dict(os.environ, SIMT_SHM_MALLOC="0" if "nccl" in extra else "1", SIMT_DEVICES=str(world), RTK_B200_HOST_MIN_SHARE_LOG2="10", RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE=str(world), LOCAL_WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", "x"] + args, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    ok = True
    lines = []
    for rank, p in enumerate(procs):
        out, err = p.communicate(timeout=1500)
        if p.returncode != 0:
            ok = False
            sys.stderr.write(f"--- rank {rank} exited {p.returncode}\n{err[-4000:]}\n")
        lines += [ln for ln in out.splitlines() if ln.startswith("{")]
    if not ok:
        return 1
    assert len(lines) == 1, f"exactly one JSON line expected, got {len(lines)}"
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "gpu_launches", "clocks"):
        assert key in line, f"missing key {key}"
    assert line["n_gpus"] == world
    assert line["parity"]["bit_exact"], line["parity"]
    if "C5" in extra:
        assert line["config"]["workload"].startswith("C5") and line["parity"]["gpu_bruteforce_bit_exact"]
        assert 0.0 < line["wavefront"]["last_bounce_hit_fraction"] <= 1.0
        print(json.dumps(line)[:1500])
        print(f"dry run ok (world {world}, wavefront); nothing above is a measurement")
        return 0
    assert line["config"]["workload"].startswith("C3")
    for key in ("e2e", "roofline"):
        assert key in line, f"missing key {key}"
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    assert line["e2e"]["rows_equal_device_path"], line["e2e"]
    assert line["e2e"]["devices"] == world and line["e2e"]["rays_per_step"] == 4096 * world, line["e2e"]
    assert line["e2e_compact"].get("records_equal_device_path") is True, line["e2e_compact"]
    assert line["e2e_pageable"]["rows_equal_pinned_path"], line["e2e_pageable"]
    assert line["c4"]["parity"]["bit_exact"] and line["c4"]["parity"]["gpu_bruteforce_bit_exact"], line["c4"]["parity"]
    assert line["c4"]["scaling"] == "strong" and line["c4"]["rays_total"] == 6000
    assert line["build"]["roofline"]["frac"] > 0
    assert line["parity"]["gpu_bruteforce_bit_exact"]
    if world == 1:
        assert line["cpu_baseline"]["kind"] in ("reference", "unavailable"), line["cpu_baseline"]
        assert line["occlusion"]["agrees_with_closest_hit_mask"]
    else:
        assert line["gather_check"].get("equal") is True, line["gather_check"]
        assert line["config"]["gather"] == ("nccl" if "nccl" in extra else "p2p")
    print(json.dumps(line)[:3000])
    print(f"dry run ok (world {world}); nothing above is a measurement")
    return 0


if __name__ == "__main__":
    sys.exit(main())
