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


def test_peer_gather_window_emulated(emu_lib, orc):
    """The copy-engine gather (shard.PeerGather over rtk_cuda_peer_*): three 'ranks' played by one
    process on the emulator (whose IPC handles are plain pointers).  Rank 0 traces straight into its
    slot of the window, the others push theirs; the window then holds exactly the records of the
    unsharded trace, in both buffers of the double-buffered scheme."""
    import ctypes as C
    from rtk_b200 import api, scenes, shard
    lib = emu_lib
    s = scenes.config_scene("C3", 0.004)
    world, per = 3, 400
    rays = scenes.bounce_rays(s, world * per)
    sc = lib.build_scene(s["meshes"])
    want = orc.trace_brute(s["tris"], rays)
    box = {}

    def exchange(h):
        if h is not None:
            box["h"] = h
        return box["h"]
    gathers = [shard.PeerGather(lib, r, world, 16 * per, 2, exchange) for r in range(world)]
    assert gathers[0].owner and not gathers[1].owner
    for buf in (0, 1):
        for r in range(world):
            lo, hi = shard.ray_range(r, world, len(rays))
            sub = np.ascontiguousarray(rays[lo:hi])
            if r == 0:
                dst = gathers[0].slot(buf)                     # the gathering rank needs no copy at all
                assert lib.rtk_trace_rays_compact_device(sc.ptr, sub.ctypes.data, dst, len(sub), None) == 0, lib.last_error()
            else:
                local = np.zeros(len(sub), dtype=api.HIT16_DTYPE)
                assert lib.rtk_trace_rays_compact_device(sc.ptr, sub.ctypes.data, local.ctypes.data, len(sub), None) == 0, lib.last_error()
                gathers[r].push(buf, local.ctypes.data, local.nbytes, None)
        got = np.ctypeslib.as_array(C.cast(gathers[0].slot(buf, 0), C.POINTER(C.c_ubyte)), shape=(world * per * 16,)).copy().view(api.HIT16_DTYPE)
        assert got.tobytes() == want.tobytes(), f"buffer {buf}"
    # a push that does not fit its slot is a caller bug
    import pytest
    with pytest.raises(AssertionError):
        gathers[1].push(0, 0, 16 * per + 16, None)
    for g in reversed(gathers):
        g.close()
    assert lib.rtk_cuda_peer_push(None, None, 16, None) != 0
    sc.free()
