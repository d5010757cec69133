"""Ray sharding across the GPUs of one box: the scene is replicated, rays are split
into contiguous ranges, there is no reduction, and the only exchange is the gather of the hit
records on rank 0.  One process per GPU, torch.distributed for the plumbing (NCCL on GPUs; the
CPU test tier drives the same code over gloo)."""
import ctypes as C

import numpy as np


def ray_range(rank, world, n):
    """Contiguous range [lo, hi) of rank `rank` out of `world` over n rays; ranges tile [0, n)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_records(local, n_total, dist, rank, world, device):
    """Gather per-rank record arrays (numpy structured, contiguous ranges in rank order) on rank 0.
    Ranges may differ in length by one, so records are padded to the longest range."""
    import torch
    itemsize = local.dtype.itemsize
    longest = max(ray_range(r, world, n_total)[1] - ray_range(r, world, n_total)[0] for r in range(world))
    buf = np.zeros((longest, itemsize), dtype=np.uint8)
    buf[:len(local)] = local.view(np.uint8).reshape(len(local), itemsize)
    t = torch.from_numpy(buf).to(device)
    out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, out, dst=0)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        lo, hi = ray_range(r, world, n_total)
        parts.append(out[r][:hi - lo].cpu().numpy().reshape(-1).view(local.dtype))
    return np.concatenate(parts)


def trace_sharded(scene, rays, dist, rank, world, device="cpu"):
    """Every rank traces its own range of `rays` against its replica of `scene` (api.Scene) and
    rank 0 receives all expanded hits and masks, in ray order.  Returns (hits, mask) on rank 0,
    (None, None) elsewhere."""
    lo, hi = ray_range(rank, world, len(rays))
    hits, mask, _ = scene.trace_rays(np.ascontiguousarray(rays[lo:hi]))
    all_hits = gather_records(hits, len(rays), dist, rank, world, device)
    all_mask = gather_records(mask.view(np.dtype([("m", "u1")])), len(rays), dist, rank, world, device)
    if rank != 0:
        return None, None
    return all_hits, all_mask.view(np.uint8)


class PeerGather:
    """The gather of SURVEY 8(e) as copy-engine pushes into a peer-memory window (rtk_cuda.h,
    rtk_cuda_peer_*) instead of an NCCL gather: rank `dst` owns a window of `nbuf` x `world` slots of
    `slot_bytes` in its HBM, every other rank maps it over NVLink and pushes its records into slot
    (buffer, rank).  No send/receive kernel runs beside the persistent traversal grid.

    `exchange(handle_or_None)` ships the 64-byte window handle from `dst` to everybody (the bench
    uses torch.distributed.broadcast_object_list) and returns it on every rank.  Streams are raw
    cudaStream_t values.  Ordering is the caller's: make the push stream wait for the kernel that
    produced the records, and synchronise it (plus a barrier) before `dst` reads the window."""

    def __init__(self, lib, rank, world, slot_bytes, nbuf, exchange, dst=0):
        self.lib, self.rank, self.world, self.dst = lib, rank, world, dst
        self.slot_bytes, self.nbuf = int(slot_bytes), int(nbuf)
        self.window, self.owner = None, rank == dst
        win = C.c_void_p()
        if self.owner:
            handle = (C.c_ubyte * 64)()
            if lib.rtk_cuda_peer_window_create(self.nbuf * world * self.slot_bytes, C.byref(win), handle) != 0:
                exchange(b"")                                   # do not leave the others waiting
                raise RuntimeError("rtk_cuda_peer_window_create: " + lib.last_error())
            exchange(bytes(handle))
        else:
            raw = exchange(None)
            if len(raw) != 64:
                raise RuntimeError("the gathering rank could not create its peer window")
            handle = (C.c_ubyte * 64).from_buffer_copy(raw)
            if lib.rtk_cuda_peer_window_open(handle, C.byref(win)) != 0:
                raise RuntimeError("rtk_cuda_peer_window_open: " + lib.last_error())
        self.window = win.value

    def slot(self, buf, rank=None):
        """device address (valid in THIS process) of slot (buf, rank) of the window"""
        rank = self.rank if rank is None else rank
        assert 0 <= buf < self.nbuf and 0 <= rank < self.world
        return self.window + (buf * self.world + rank) * self.slot_bytes

    def push(self, buf, d_src, nbytes, stream):
        """this rank's records -> its slot of buffer `buf`, asynchronous on `stream`"""
        assert nbytes <= self.slot_bytes
        if self.lib.rtk_cuda_peer_push(self.slot(buf), d_src, nbytes, stream) != 0:
            raise RuntimeError("rtk_cuda_peer_push: " + self.lib.last_error())

    def close(self):
        if self.window is None:
            return
        fn = self.lib.rtk_cuda_peer_window_destroy if self.owner else self.lib.rtk_cuda_peer_window_close
        fn(self.window)
        self.window = None
