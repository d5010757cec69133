"""Multi-GPU host logic on the CPU: world_size 2 over gloo.  Each rank owns a replica of the scene
(emulated kernels), traces its contiguous ray range and rank 0 gathers -- the result must be
byte-identical to the single-process result."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import torch.distributed as dist
    import build_emu
    from rtk_b200 import api, scenes, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = api.Library(build_emu.build())
    assert lib.rtk_cuda_init(0) == 0
    s = scenes.config_scene("C3", 0.004)
    rays = scenes.bounce_rays(s, 1001)                     # odd count: ranges differ in length
    sc = lib.build_scene(s["meshes"], mode=api.RTK_CUDA_BUILD_SAH)
    hits, mask = shard.trace_sharded(sc, rays, dist, rank, world, device="cpu")
    if rank == 0:
        full_hits, full_mask, _ = sc.trace_rays(rays)
        ok = hits.tobytes() == full_hits.tobytes() and mask.tobytes() == full_mask.tobytes()
        with open(out_path, "w") as f:
            f.write("ok %d" % int(mask.sum()) if ok else "mismatch")
    sc.free()
    dist.barrier()
    dist.destroy_process_group()


def test_ray_ranges_tile():
    from rtk_b200 import shard
    for n in (0, 1, 7, 1000, 16_777_216):
        for w in (1, 2, 3, 8):
            r = [shard.ray_range(k, w, n) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_ranks_gloo(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    build_emu.build()
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = open(out).read()
    assert res.startswith("ok"), res
    assert int(res.split()[1]) > 100
