#!/usr/bin/env python
"""Randomised parity hunt on the CPU: the product's kernels (compiled against the SIMT emulator,
tests/emu) against the oracle on small adversarial scenes -- mixed scales, slivers, degenerate and
duplicated triangles, axis-aligned geometry, rays along axes / through vertices / starting on
surfaces, tight [min_t, max_t] windows -- under both builders, with and without a triangle filter,
closest-hit and occlusion.  TEST INFRASTRUCTURE: only the emulator library is loaded.

    python tools/fuzz_emu.py [seconds] [first_seed]

Prints one line per failing seed (and stops at the first by default); exit code 1 on a mismatch."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

import build_emu  # noqa: E402
import parity_cases as pc  # noqa: E402
from oracle import orc  # noqa: E402
from rtk_b200 import api  # noqa: E402


def make_scene(rng):
    kind = rng.integers(0, 7)
    n = int(rng.choice([1, 2, 3, 7, 8, 9, 16, 17, 60, 130, 400, 900]))
    scale = np.float32(rng.choice([1e-3, 1.0, 1.0, 1e3]))
    if kind == 0:                                   # soup
        c = rng.random((n, 1, 3))
        tris = c + 0.2 * (rng.random((n, 3, 3)) - 0.5)
    elif kind == 1:                                 # slivers and needles
        c = rng.random((n, 1, 3))
        e = (rng.random((n, 3, 3)) - 0.5) * np.array([1.0, 1e-4, 1e-2])
        tris = c + e
    elif kind == 2:                                 # axis-aligned quads on a lattice (exact zeros everywhere)
        q = rng.integers(0, 5, size=(n, 1, 3)).astype(np.float64) / 4
        ax = rng.integers(0, 3, size=n)
        tris = np.repeat(q, 3, axis=1)
        for i in range(n):
            a, b = (ax[i] + 1) % 3, (ax[i] + 2) % 3
            tris[i, 1, a] += 0.25
            tris[i, 2, b] += 0.25
    elif kind == 3:                                 # many duplicates and shared edges
        base = rng.random((max(1, n // 4), 3, 3))
        tris = base[rng.integers(0, len(base), n)]
    elif kind == 4:                                 # mixed sizes: a few huge triangles over small ones
        c = rng.random((n, 1, 3))
        s = rng.choice([1e-3, 1e-2, 0.1, 3.0], size=(n, 1, 1))
        tris = c + s * (rng.random((n, 3, 3)) - 0.5)
    elif kind == 5:                                 # degenerate: zero-area, repeated vertices, collinear
        c = rng.random((n, 1, 3))
        tris = c + 0.2 * (rng.random((n, 3, 3)) - 0.5)
        k = rng.integers(0, 3, size=n)
        for i in range(0, n, 2):
            if k[i] == 0:
                tris[i, 1] = tris[i, 0]
            elif k[i] == 1:
                tris[i, 2] = 0.5 * (tris[i, 0] + tris[i, 1])
            else:
                tris[i, :] = tris[i, 0]
    else:                                           # a flat heightfield patch (coplanar neighbours)
        g = int(np.ceil(np.sqrt(n / 2))) + 1
        xs, ys = np.meshgrid(np.arange(g) / (g - 1), np.arange(g) / (g - 1))
        h = np.where(rng.random((g, g)) < 0.5, 0.0, rng.random((g, g)) * 0.1)
        p = np.stack([xs, h, ys], -1)
        t = []
        for j in range(g - 1):
            for i in range(g - 1):
                t.append([p[j, i], p[j, i + 1], p[j + 1, i]])
                t.append([p[j, i + 1], p[j + 1, i + 1], p[j + 1, i]])
        tris = np.array(t)
    off = np.float32(rng.choice([0.0, 0.0, 100.0, -1e4]))
    return np.ascontiguousarray((tris * scale + off).astype(np.float32))


def make_rays(rng, tris, n):
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    ext = np.maximum(hi - lo, 1e-6)
    rays = np.zeros(n, dtype=api.RAY_DTYPE)
    o = lo + (rng.random((n, 3)) * 1.6 - 0.3) * ext
    tgt = tris[rng.integers(0, len(tris), n)]
    w = rng.random((n, 3))
    w /= w.sum(1, keepdims=True)
    mode = rng.integers(0, 6, size=n)
    p = (tgt * w[:, :, None]).sum(1)
    p = np.where((mode == 1)[:, None], tgt[:, 0], p)                       # through a vertex
    p = np.where((mode == 2)[:, None], 0.5 * (tgt[:, 0] + tgt[:, 1]), p)   # through an edge midpoint
    d = p - o
    axis = rng.integers(0, 3, size=n)
    ax_d = np.zeros((n, 3))
    ax_d[np.arange(n), axis] = rng.choice([-1.0, 1.0], size=n)
    d = np.where((mode == 3)[:, None], ax_d, d)                            # along an axis
    o = np.where((mode == 4)[:, None], p, o)                               # starting on a surface
    d = np.where((mode == 4)[:, None], rng.normal(size=(n, 3)), d)
    d = d * rng.choice([1e-3, 1.0, 1.0, 50.0], size=(n, 1))
    dead = np.abs(d).max(1) == 0
    d[dead] = (0, 0, 1)
    rays["o"], rays["d"] = o.astype(np.float32), d.astype(np.float32)
    bad = np.abs(rays["d"]).max(1) == 0
    rays["d"][bad] = (0, 1, 0)
    rays["min_t"] = rng.choice([0.0, 0.0, 0.0, 1e-4, 0.5, -1.0], size=n).astype(np.float32)
    rays["max_t"] = rng.choice([api.RTK_INF, api.RTK_INF, 1.0, 1.0, 0.999, 2.0, 1e-3], size=n).astype(np.float32)
    return rays


def one(lib, seed):
    rng = np.random.default_rng(seed)
    tris = make_scene(rng)
    rays = make_rays(rng, tris, int(rng.choice([1, 31, 33, 200, 700])))
    mode = int(rng.integers(0, 2))
    lib.rtk_cuda_set_build_mode(mode)
    lib.rtk_cuda_set_cull_mode(1)
    sc = lib.build_scene(pc.soup_mesh(tris))
    try:
        keep = None
        if rng.random() < 0.3:
            keep = rng.random(len(tris)) < 0.6
            sc.set_triangle_filter(keep)
        want = pc.filtered_oracle(orc, tris, rays, keep) if keep is not None else orc.trace_brute(tris, rays)
        hits, mask, nh = sc.trace_rays(rays)
        got = api.hits_to_hit16(hits, mask, [0, len(tris)])
        pc.assert_same(got, want, f"seed {seed}: {len(tris)} triangles, {len(rays)} rays, mode {mode}, filter {keep is not None}")
        occ = np.full(len(rays), 7, dtype=np.uint8)
        r = np.ascontiguousarray(rays)
        assert lib.rtk_occluded_rays_device(sc.ptr, r.ctypes.data, occ.ctypes.data, len(rays), None) == 0, lib.last_error()
        assert np.array_equal(occ.astype(bool), want["prim"] != api.RTK_CUDA_MISS), f"seed {seed}: occlusion query"
    finally:
        sc.free()
    return len(tris), len(rays), int((want["prim"] != api.RTK_CUDA_MISS).sum())


def one_update(lib, seed):
    """device-mesh build (indexed, shared vertices), then vertex perturbations with refit / rebuild,
    optionally under a triangle filter -- against the oracle on the moved triangles"""
    rng = np.random.default_rng(seed)
    g = int(rng.integers(3, 14))
    xs, ys = np.meshgrid(np.arange(g) / (g - 1), np.arange(g) / (g - 1))
    pos = np.stack([xs.ravel(), 0.2 * rng.random(g * g), ys.ravel()], -1).astype(np.float32)
    quads = [(j * g + i, j * g + i + 1, (j + 1) * g + i, (j + 1) * g + i + 1) for j in range(g - 1) for i in range(g - 1)]
    idx = np.array([t for a, b, c, d in quads for t in ((a, b, c), (b, d, c))], dtype=np.uint32)
    meshes = (api.rtk_cuda_mesh * 1)()
    ix = np.ascontiguousarray(idx)
    meshes[0].d_indices, meshes[0].num_vertices, meshes[0].num_triangles = ix.ctypes.data, len(pos), len(idx)
    p0 = np.ascontiguousarray(pos)
    meshes[0].d_positions = p0.ctypes.data
    lib.rtk_cuda_set_build_mode(int(rng.integers(0, 2)))
    ptr = lib.rtk_cuda_build_scene(meshes, 1, None)
    assert ptr, lib.last_error()
    sc = api.Scene(lib, ptr)
    hits_total = 0
    try:
        keep = None
        if rng.random() < 0.4:
            keep = rng.random(len(idx)) < 0.7
            sc.set_triangle_filter(keep)
        for step in range(3):
            cur = (pos + (rng.random(pos.shape) - 0.5) * np.float32(rng.choice([1e-3, 0.05, 0.5]))).astype(np.float32)
            if rng.random() < 0.3:
                cur *= np.float32(rng.choice([0.01, 7.0]))
            cur = np.ascontiguousarray(cur)
            meshes[0].d_positions = cur.ctypes.data
            mode = int(rng.integers(0, 2))
            assert lib.rtk_cuda_update_scene(sc.ptr, meshes, 1, mode, None) == 0, lib.last_error()
            tris = np.ascontiguousarray(cur[idx.astype(np.int64)])
            rays = make_rays(rng, tris, int(rng.choice([33, 150])))
            want = pc.filtered_oracle(orc, tris, rays, keep) if keep is not None else orc.trace_brute(tris, rays)
            got = sc.trace_rays_compact(rays)
            pc.assert_same(got, want, f"update seed {seed} step {step} mode {mode} filter {keep is not None}")
            hits_total += int((want["prim"] != api.RTK_CUDA_MISS).sum())
    finally:
        sc.free()
    return len(idx), 0, hits_total


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = api.Library(build_emu.build())
    assert lib.rtk_cuda_init(0) == 0
    t0, n, hits, rays = time.time(), 0, 0, 0
    while time.time() - t0 < budget:
        try:
            _, r, h = one_update(lib, seed) if seed % 4 == 3 else one(lib, seed)
        except AssertionError as ex:
            print("MISMATCH", ex)
            return 1
        n += 1
        rays += r
        hits += h
        seed += 1
    print(f"fuzz: {n} scenes, {rays} rays, {hits} hits, all bit-exact; next seed {seed}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
