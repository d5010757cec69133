"""Parity cases shared by the CPU tier (emulated kernels, tests/emu) and the GPU tier (-m gpu).
Every case goes through the C ABI of `lib` and is checked against the CPU oracle bit for bit
(hit index, t, u, v) -- the north star allows 1e-5 relative on t/u/v, the tests demand equality."""
import ctypes as C
import json
import os
import struct

import numpy as np

from rtk_b200 import api, scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kat.json")


def unhex(h):
    return struct.unpack("<f", struct.pack("<I", int(h, 16)))[0]


def load_kats():
    with open(GOLDEN) as f:
        g = json.load(f)
    kats = []
    for k in g["kats"]:
        tris = np.array([[unhex(c) for c in t] for t in k["tris"]], dtype=np.float32).reshape(-1, 3, 3)
        ray = np.zeros(1, dtype=api.RAY_DTYPE)
        ray["o"] = [unhex(x) for x in k["ray"]["o"]]
        ray["d"] = [unhex(x) for x in k["ray"]["d"]]
        ray["min_t"] = unhex(k["ray"]["min_t"])
        ray["max_t"] = unhex(k["ray"]["max_t"])
        kats.append((k["name"], tris, ray, k["expect"], k["literal"], k["num_real_tris"]))
    r = g["random"]
    rnd = (np.array(r["tris"], dtype=np.uint32).view(np.float32).reshape(-1, 3, 3),
           np.array(r["rays"], dtype=np.uint32).view(api.RAY_DTYPE),
           np.array(r["hits"], dtype=np.uint32).view(api.HIT16_DTYPE))
    return kats, rnd


def expect_hit16(expect):
    h = np.zeros(1, dtype=api.HIT16_DTYPE)
    h["prim"] = api.RTK_CUDA_MISS
    if expect["hit"]:
        h["t"], h["u"], h["v"] = unhex(expect["t"]), unhex(expect["u"]), unhex(expect["v"])
        h["prim"] = expect["prim"]
    return h


def soup_mesh(tris):
    return [{"positions": np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 3), "indices": None}]


def trace_hit16(lib, scene_dict_or_meshes, rays, mesh_first=None, mode=None):
    """build through rtk_build_scene, trace through rtk_trace_rays, return compact records."""
    meshes = scene_dict_or_meshes["meshes"] if isinstance(scene_dict_or_meshes, dict) else scene_dict_or_meshes
    if mesh_first is None:
        mesh_first = scene_dict_or_meshes["mesh_first"] if isinstance(scene_dict_or_meshes, dict) else \
            np.cumsum([0] + [len(m["indices"]) if m.get("indices") is not None else len(m["positions"]) // 3 for m in meshes])
    sc = lib.build_scene(meshes, mode=mode)
    try:
        hits, mask, nh = sc.trace_rays(rays)
        assert nh == int(mask.sum())
        return api.hits_to_hit16(hits, mask, mesh_first), hits, mask
    finally:
        sc.free()


def assert_same(got, want, what=""):
    if got.tobytes() == want.tobytes():
        return
    bad = np.nonzero((got["prim"] != want["prim"]) | (got["t"].view(np.uint32) != want["t"].view(np.uint32)) |
                     (got["u"].view(np.uint32) != want["u"].view(np.uint32)) |
                     (got["v"].view(np.uint32) != want["v"].view(np.uint32)))[0]
    i = bad[0]
    raise AssertionError(f"{what}: {len(bad)} of {len(got)} rays differ; first ray {i}: got {got[i]} want {want[i]}")


# ---------------------------------------------------------------------------------------------

def case_kats(lib, orc):
    kats, (rt, rr, rh) = load_kats()
    for name, tris, ray, expect, literal, nreal in kats:
        got, hits, mask = trace_hit16(lib, soup_mesh(tris), ray)
        assert_same(got, expect_hit16(expect), name)
        # the bare scene (what rtk.c literally sees) stays within the north star's tolerance
        got2, _, _ = trace_hit16(lib, soup_mesh(tris[:nreal]), ray)
        lit = expect_hit16(literal)
        assert got2["prim"][0] == lit["prim"][0], name
        if literal["hit"]:
            for f in ("t", "u", "v"):
                assert abs(float(got2[f][0]) - float(lit[f][0])) <= 1e-5 * max(abs(float(lit[f][0])), 1e-30) + 1e-7, (name, f)
    got, _, _ = trace_hit16(lib, soup_mesh(rt), rr)
    assert_same(got, rh, "golden random soup")


def case_config(lib, orc, name, scale, nrays, mode=None):
    s = scenes.config_scene(name, scale)
    if name == "C1":
        rays = scenes.config_rays("C1", s)
        if nrays:
            rays = rays[:: max(1, len(rays) // nrays)][:nrays]
    elif name == "C2":
        side = max(4, int(np.sqrt(nrays)))
        rays = scenes.soup_primary_rays(side * 16 // 9, side)
    elif name == "C3":
        rays = scenes.bounce_rays(s, nrays)
    else:
        rays = scenes.mixed_rays(s, nrays, block=max(64, nrays // 12))
    got, hits, mask = trace_hit16(lib, s, rays, mode=mode)
    want = orc.trace_brute(s["tris"], rays)
    assert_same(got, want, f"{name} x{scale}")
    m = mask.astype(bool)
    prim = want["prim"][m].astype(np.int64)
    # payload: the three rtk_vertex of the hit triangle, mesh and per-mesh triangle number
    assert np.array_equal(hits["vertex"]["position"][m], s["tris"][prim])
    mesh = np.searchsorted(s["mesh_first"], prim, side="right") - 1
    assert np.array_equal(hits["mesh_index"][m], mesh)
    assert np.array_equal(hits["triangle_index"][m], prim - s["mesh_first"][mesh])
    vidx = []
    for k, me in enumerate(s["meshes"]):
        nt = int(s["mesh_first"][k + 1] - s["mesh_first"][k])
        vidx.append(me["indices"].astype(np.uint32) if me["indices"] is not None else np.arange(3 * nt, dtype=np.uint32).reshape(-1, 3))
    vidx = np.concatenate(vidx)
    assert np.array_equal(hits["vertex"]["index"][m], vidx[prim])
    return int(m.sum())


# (scene, build mode) -> wide nodes, leaves, depth, SAH cost.  The builders are deterministic functions of the triangle
# SET (bins, counts and boxes do not depend on the order the triangles arrive in), so these numbers only move when
# the algorithm does: a guard for the restructurings of the build kernels, which must leave the trees alone.
TREE_STATS = {
    ("terrain 200x100", 1): (1353, 5217, 6, 9.980880375670187),
    ("terrain 200x100", 0): (1638, 7828, 7, 24.284567995333656),
    ("soup 30000", 1): (771, 3853, 5, 30.131097424911744),
    ("soup 30000", 0): (1473, 5396, 6, 32.72317522549396),
    ("terrain 40x30", 1): (74, 310, 4, 6.014886789363808),
    ("terrain 40x30", 0): (104, 474, 5, 11.919080584366526),
}


def case_tree_stats(lib):
    made = {"terrain 200x100": lambda: scenes.terrain(200, 100), "soup 30000": lambda: scenes.soup(30000),
            "terrain 40x30": lambda: scenes.terrain(40, 30)}
    try:
        for (name, mode), (nodes, leaves, depth, cost) in TREE_STATS.items():
            sc = lib.build_scene(made[name]()["meshes"], mode=mode)
            try:
                i = sc.info()
                got = (int(i.num_wide_nodes), int(i.num_leaves), int(i.wide_depth))
                assert got == (nodes, leaves, depth), f"{name}, mode {mode}: tree {got}, expected {(nodes, leaves, depth)}"
                # the areas are x*y + y*z + z*x: nvcc contracts them into FMAs, the emulator build does not
                assert abs(i.sah_cost - cost) <= 1e-6 * cost, f"{name}, mode {mode}: SAH cost {i.sah_cost!r}, expected {cost!r}"
            finally:
                sc.free()
    finally:
        lib.rtk_cuda_set_build_mode(api.RTK_CUDA_BUILD_SAH)     # the library's default


def case_edge_scenes(lib, orc):
    ray = np.zeros(4, dtype=api.RAY_DTYPE)
    ray["o"] = [(0.25, 0.25, 0), (0.25, 0.25, 0), (5, 5, 0), (0.25, 0.25, 2)]
    ray["d"] = [(0, 0, 1), (0, 0, -1), (0, 0, 1), (0, 0, -1)]
    ray["max_t"] = api.RTK_INF
    # empty scene
    sc = lib.build_scene([])
    hits, mask, nh = sc.trace_rays(ray)
    assert nh == 0 and not mask.any()
    sc.free()
    sc = lib.build_scene([{"positions": np.zeros((0, 3), np.float32), "indices": None}])
    assert sc.trace_rays(ray)[2] == 0
    sc.free()
    # one triangle, then 2..20 triangles (leaf / node boundary cases)
    rng = np.random.default_rng(7)
    for n in list(range(1, 21)) + [63, 64, 65]:
        tris = rng.random((n, 3, 3)).astype(np.float32)
        tris[0] = [(0, 0, 1), (1, 0, 1), (0, 1, 1)]
        got, _, _ = trace_hit16(lib, soup_mesh(tris), ray)
        assert_same(got, orc.trace_brute(tris, ray), f"{n} triangles")
    # zero rays
    sc = lib.build_scene(soup_mesh(tris))
    assert sc.trace_rays(ray[:0])[2] == 0
    sc.free()


def case_ties(lib, orc):
    """duplicates and shared edges: exact ties go to the lowest triangle number, whatever the
    order the BVH presents them in"""
    rng = np.random.default_rng(11)
    base = rng.random((40, 3, 3)).astype(np.float32)
    tris = np.concatenate([base, base[::-1], base])            # every triangle three times
    rays = np.zeros(600, dtype=api.RAY_DTYPE)
    rays["o"] = (rng.random((600, 3)) * 2 - 0.5).astype(np.float32)
    tgt = base[rng.integers(0, 40, 600)].mean(axis=1)
    rays["d"] = tgt - rays["o"]
    rays["max_t"] = api.RTK_INF
    got, _, _ = trace_hit16(lib, soup_mesh(tris), rays)
    want = orc.trace_brute(tris, rays)
    assert_same(got, want, "duplicated triangles")
    assert (want["prim"][want["prim"] != api.RTK_CUDA_MISS] < 80).all()
    # coplanar grid, rays through grid vertices and edges (exact zeros -> fp64 path)
    g = scenes._quad_grid((0, 0, 1), (1, 0, 0), (0, 1, 0), 8, 8)
    grid = g[0][g[1].astype(np.int64)]
    gx, gy = np.meshgrid(np.arange(0, 17) / 16.0, np.arange(0, 17) / 16.0)
    r2 = np.zeros(gx.size, dtype=api.RAY_DTYPE)
    r2["o"] = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], -1)
    r2["d"] = (0, 0, 1)
    r2["max_t"] = api.RTK_INF
    got, _, _ = trace_hit16(lib, soup_mesh(grid), r2)
    want = orc.trace_brute(grid, r2)
    assert_same(got, want, "coplanar grid")
    assert (want["prim"] != api.RTK_CUDA_MISS).all()          # watertight: no ray slips through an edge


def case_ray_limits(lib, orc):
    """min_t / max_t strictness, unnormalised directions, zero direction components, -0.0"""
    rng = np.random.default_rng(13)
    tris = (rng.random((300, 1, 3)) + 0.1 * (rng.random((300, 3, 3)) * 2 - 1)).astype(np.float32)
    n = 1500
    rays = np.zeros(n, dtype=api.RAY_DTYPE)
    rays["o"] = (rng.random((n, 3)) * 1.4 - 0.2).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32) * rng.choice([1e-3, 1.0, 1e3], size=(n, 1)).astype(np.float32)
    d[::7, 0] = 0.0
    d[::11, 1] = -0.0
    d[::13, 2] = 0.0
    d[(np.abs(d).max(axis=1) == 0)] = (0, 0, 1)
    rays["d"] = d
    rays["min_t"] = rng.choice([0.0, 0.0, 1e-3, 0.3], size=n).astype(np.float32)
    rays["max_t"] = rng.choice([api.RTK_INF, api.RTK_INF, 0.5, 2.0], size=n).astype(np.float32)
    got, _, _ = trace_hit16(lib, soup_mesh(tris), rays)
    want = orc.trace_brute(tris, rays)
    assert_same(got, want, "ray limits")
    # a second pass whose limits sit exactly on the found t: both ends are strict (rtk.c:354)
    hit = want["prim"] != api.RTK_CUDA_MISS
    r2 = rays[hit].copy()
    r2["max_t"] = want["t"][hit]
    r3 = rays[hit].copy()
    r3["min_t"] = want["t"][hit]
    for rr in (r2, r3):
        g2, _, _ = trace_hit16(lib, soup_mesh(tris), rr)
        w2 = orc.trace_brute(tris, rr)
        assert_same(g2, w2, "limits on t")
        assert (g2["prim"][:] != want["prim"][hit]).all()


def case_invariances(lib, orc):
    """triangle order permutes ids but not t; ray order permutes results only"""
    s = scenes.config_scene("C3", 0.004)
    rays = scenes.bounce_rays(s, 1500)
    a, _, _ = trace_hit16(lib, s, rays)
    perm = np.random.default_rng(5).permutation(len(s["tris"]))
    tp = s["tris"][perm]
    b, _, _ = trace_hit16(lib, soup_mesh(tp), rays)
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    hit = a["prim"] != api.RTK_CUDA_MISS
    # ids map back unless an exact tie picked another of the tied triangles
    same = perm[b["prim"][hit].astype(np.int64)] == a["prim"][hit]
    assert same.mean() > 0.98
    rp = np.random.default_rng(6).permutation(len(rays))
    c, _, _ = trace_hit16(lib, s, rays[rp])
    assert_same(c, a[rp], "ray order")


def case_mesh_formats(lib, orc):
    """U16 / U32 / implicit indices, F32 / F64 positions, strides, callbacks, DEFAULT types
    (reference rtk.c:1028-1114)"""
    s = scenes.config_scene("C1")
    m = s["meshes"][0]
    rays = scenes.config_rays("C1", s)[::131]
    want = orc.trace_brute(s["tris"], rays)
    pos, idx = m["positions"], m["indices"]
    variants = []
    variants.append(("u16", {"positions": pos, "indices": idx}))
    variants.append(("u32", {"positions": pos, "indices": idx.astype(np.uint32)}))
    variants.append(("default types", {"positions": pos, "indices": idx.astype(np.uint32),
                                       "position_type": api.RTK_TYPE_DEFAULT, "index_type": api.RTK_TYPE_DEFAULT}))
    variants.append(("real", {"positions": pos, "indices": idx, "position_type": api.RTK_TYPE_REAL}))
    variants.append(("f64", {"positions": pos.astype(np.float64), "indices": idx}))
    pad = np.zeros((len(pos), 5), dtype=np.float32)
    pad[:, :3] = pos
    pad[:, 3:] = 777.0
    variants.append(("position stride 20", {"positions": pad, "indices": idx, "position_stride": 20}))
    ipad = np.full((len(idx), 4), 65535, dtype=np.uint16)
    ipad[:, :3] = idx
    variants.append(("index stride 8", {"positions": pos, "indices": ipad, "index_stride": 8}))
    variants.append(("implicit", {"positions": np.ascontiguousarray(s["tris"].reshape(-1, 3)), "indices": None}))
    for name, mesh in variants:
        if name == "index stride 8":
            desc, keep = lib.make_desc([{"positions": pos, "indices": idx}])
            desc.meshes[0].index.data = ipad.ctypes.data
            desc.meshes[0].index.stride = 8
            ptr = lib.rtk_build_scene(C.byref(desc))
            assert ptr, lib.last_error()
            sc = api.Scene(lib, ptr)
            hits, mask, _ = sc.trace_rays(rays)
            got = api.hits_to_hit16(hits, mask, s["mesh_first"])
            sc.free()
        else:
            got, hits, mask = trace_hit16(lib, [mesh], rays, mesh_first=s["mesh_first"])
        assert_same(got, want, name)
        if name in ("u16", "u32", "f64", "position stride 20"):
            mm = mask.astype(bool)
            assert np.array_equal(hits["vertex"]["index"][mm], idx[want["prim"][mm].astype(np.int64)].astype(np.uint32)), name

    # callbacks (rtk.h:61-62): positions and indices pulled through user functions
    calls = {"pos": 0, "idx": 0, "max": 0}
    idx32 = idx.astype(np.uint32)

    def pos_cb(user, mesh, dst, indices, count):
        calls["pos"] += 1
        calls["max"] = max(calls["max"], count)
        ii = np.ctypeslib.as_array(indices, shape=(3 * count,))
        out = np.ctypeslib.as_array(C.cast(dst, C.POINTER(C.c_float)), shape=(3 * count, 3))
        out[:] = pos[ii]

    def idx_cb(user, mesh, dst, offset, count):
        calls["idx"] += 1
        out = np.ctypeslib.as_array(dst, shape=(3 * count,))
        out[:] = idx32[offset:offset + count].reshape(-1)
    pcb, icb = api.rtk_position_callback_fn(pos_cb), api.rtk_index_callback_fn(idx_cb)
    mesh = api.rtk_mesh()
    mesh.num_triangles = len(idx)
    mesh.position_cb = pcb
    mesh.index_cb = icb
    desc = api.rtk_scene_desc()
    desc.meshes = C.pointer(mesh)
    desc.num_meshes = 1
    ptr = lib.rtk_build_scene(C.byref(desc))
    assert ptr, lib.last_error()
    sc = api.Scene(lib, ptr)
    hits, mask, _ = sc.trace_rays(rays)
    assert_same(api.hits_to_hit16(hits, mask, s["mesh_first"]), want, "callbacks")
    assert calls["pos"] > 0 and calls["idx"] > 0 and calls["max"] <= 128       # rtk.c:1143
    sc.free()


def case_api_semantics(lib, orc):
    """rtk_trace_ray miss rule, filter, split-phase build, relocatable blob, log callback"""
    s = scenes.config_scene("C1")
    rays = scenes.config_rays("C1", s)[::997]
    want = orc.trace_brute(s["tris"], rays)
    logs = []
    log = api.rtk_log_fn(lambda user, build, msg: logs.append(msg.decode()))
    desc, keep = lib.make_desc(s["meshes"], log_fn=log)
    ptr = lib.rtk_build_scene(C.byref(desc))
    assert ptr
    assert any("Starting build" in m for m in logs)
    sc = api.Scene(lib, ptr)
    hdr = sc.header()
    assert hdr.magic == b"\x00RTK\r\n\x1a\n"[:len(hdr.magic)] or bytes(hdr.magic) == b""  # c_char stops at NUL
    assert hdr.endian == 0xAABB and hdr.sizeof_real == 4 and hdr.size_in_bytes >= 128
    # single-ray entry point: *hit untouched on a miss (rtk.c:571-576)
    nh = nm = 0
    for i in range(0, len(rays), 5):
        r = rays[i:i + 1]
        h = np.zeros(1, dtype=api.HIT_DTYPE)
        h.view(np.uint8)[:] = 0xAB
        ok = lib.rtk_trace_ray(sc.ptr, C.cast(r.ctypes.data, C.POINTER(api.rtk_ray)), C.cast(h.ctypes.data, C.POINTER(api.rtk_hit)))
        if want["prim"][i] == api.RTK_CUDA_MISS:
            assert not ok and (h.view(np.uint8) == 0xAB).all()
            nm += 1
        else:
            assert ok and h["t"][0] == want["t"][i] and h["triangle_index"][0] == want["prim"][i]
            nh += 1
    assert nh > 0
    # filter: rejecting everything reports misses, accepting everything equals rtk_trace_ray
    i = int(np.nonzero(want["prim"] != api.RTK_CUDA_MISS)[0][0])
    r = rays[i:i + 1]
    h = np.zeros(1, dtype=api.HIT_DTYPE)
    rej = api.rtk_filter_fn(lambda u, ray, hit: False)
    acc = api.rtk_filter_fn(lambda u, ray, hit: True)
    rp, hp = C.cast(r.ctypes.data, C.POINTER(api.rtk_ray)), C.cast(h.ctypes.data, C.POINTER(api.rtk_hit))
    assert not lib.rtk_trace_ray_filter(sc.ptr, rp, hp, rej, None)
    assert lib.rtk_trace_ray_filter(sc.ptr, rp, hp, acc, None) and h["t"][0] == want["t"][i]
    sc.free()

    # split-phase build into a caller buffer; too-small buffer -> NULL and the build survives
    desc, keep = lib.make_desc(s["meshes"])
    first = api.rtk_task()
    b = lib.rtk_start_build(C.byref(desc), C.byref(first))
    assert b and first.fn
    queue = (api.rtk_task * 4)()
    assert lib.rtk_run_task(C.byref(first), queue, 4) == 0
    size = lib.rtk_get_build_size(b)
    assert size > 128 and size % 128 == 0
    small = np.zeros(size - 128, dtype=np.uint8)
    assert not lib.rtk_finish_build_to(b, small.ctypes.data, small.nbytes)
    buf = np.zeros(size + 128, dtype=np.uint8)
    off = (-buf.ctypes.data) % 128
    p = lib.rtk_finish_build_to(b, buf.ctypes.data + off, size)
    assert p == buf.ctypes.data + off
    sc = api.Scene(lib, p, owner=buf)
    assert sc.header().size_in_bytes == size
    hits, mask, _ = sc.trace_rays(rays)
    assert_same(api.hits_to_hit16(hits, mask, s["mesh_first"]), want, "finish_build_to")
    # relocate: copy the blob elsewhere, drop the original, trace the copy
    blob = buf[off:off + size].copy()
    moved = np.zeros(size + 128, dtype=np.uint8)
    off2 = (-moved.ctypes.data) % 128
    moved[off2:off2 + size] = blob
    sc.free()
    buf[:] = 0
    sc2 = api.Scene(lib, moved.ctypes.data + off2, owner=moved)
    hits, mask, _ = sc2.trace_rays(rays)
    assert_same(api.hits_to_hit16(hits, mask, s["mesh_first"]), want, "relocated blob")
    assert lib.rtk_cuda_detach_scene(sc2.ptr) == 0
    sc2.ptr = None
    # a foreign blob is refused, loudly
    junk = np.zeros(256, dtype=np.uint8)
    assert lib.rtk_trace_rays(junk.ctypes.data, rays.ctypes.data, hits.ctypes.data, mask.ctypes.data, 1) == C.c_size_t(-1).value
    assert "magic" in lib.last_error()
    # rtk_finish_build (library-allocated blob)
    desc, keep = lib.make_desc(s["meshes"])
    b = lib.rtk_start_build(C.byref(desc), None)
    p = lib.rtk_finish_build(b)
    assert p
    sc = api.Scene(lib, p)
    assert sc.header().size_in_bytes == size
    hits, mask, _ = sc.trace_rays(rays)
    assert_same(api.hits_to_hit16(hits, mask, s["mesh_first"]), want, "finish_build")
    sc.free()
    # rtk_cuda_shutdown releases the library's own device memory (staging of the host batches);
    # the next call initialises again by itself
    lib.rtk_cuda_shutdown()
    got, _, _ = trace_hit16(lib, s, rays)
    assert_same(got, want, "after shutdown and implicit re-initialisation")
    lib.rtk_cuda_shutdown()
    assert lib.rtk_cuda_init(0) == 0
    got, _, _ = trace_hit16(lib, s, rays)
    assert_same(got, want, "after shutdown and rtk_cuda_init")


def case_threads(lib, orc, nthreads=4):
    """rtk_trace_ray is re-entrant and read-only on the scene in the reference (rtk.h:129, called from
    the user's own worker threads): several host threads query ONE scene at once -- single rays and
    batches -- and every answer must be the oracle's; one thread builds and frees other scenes
    meanwhile."""
    import threading
    s = scenes.config_scene("C1")
    rays = scenes.config_rays("C1", s)[::211]
    want = orc.trace_brute(s["tris"], rays)
    sc = lib.build_scene(s["meshes"])
    errors = []

    def single(tid):
        try:
            for i in range(tid, len(rays), nthreads):
                h = sc.trace_ray(rays[i])
                if want["prim"][i] == api.RTK_CUDA_MISS:
                    assert h is None, f"ray {i}: hit instead of a miss"
                else:
                    assert h is not None and h["t"] == want["t"][i] and h["u"] == want["u"][i] and h["triangle_index"] == want["prim"][i], f"ray {i}"
        except Exception as ex:                                   # noqa: BLE001 -- reported by the main thread
            errors.append(ex)

    def batch(tid):
        try:
            for rep in range(3):
                sub = np.ascontiguousarray(rays[tid::2])
                hits, mask, _ = sc.trace_rays(sub)
                assert_same(api.hits_to_hit16(hits, mask, s["mesh_first"]), want[tid::2], f"batch thread {tid} rep {rep}")
        except Exception as ex:                                   # noqa: BLE001
            errors.append(ex)

    def builder():
        try:
            s2 = scenes.config_scene("C3", 0.002)
            r2 = scenes.bounce_rays(s2, 300)
            w2 = orc.trace_brute(s2["tris"], r2)
            for rep in range(2):
                got, _, _ = trace_hit16(lib, s2, r2)
                assert_same(got, w2, f"builder thread rep {rep}")
        except Exception as ex:                                   # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=single, args=(t,)) for t in range(nthreads)]
    threads += [threading.Thread(target=batch, args=(t,)) for t in range(2)]
    threads.append(threading.Thread(target=builder))
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    sc.free()
    if errors:
        raise errors[0]


def case_occlusion(lib, orc, alloc):
    """any-hit query == hit mask of the closest-hit query (device buffers via `alloc`, which maps a
    numpy array to a device pointer holder and back)"""
    s = scenes.config_scene("C3", 0.004)
    rays = scenes.bounce_rays(s, 3000)
    rays["max_t"][::3] = 0.05                      # short shadow-style segments too
    want = orc.trace_brute(s["tris"], rays)["prim"] != api.RTK_CUDA_MISS
    sc = lib.build_scene(s["meshes"])
    d_rays, d_occ, fetch = alloc(rays, len(rays))
    assert lib.rtk_occluded_rays_device(sc.ptr, d_rays, d_occ, len(rays), None) == 0, lib.last_error()
    got = fetch().astype(bool)
    assert np.array_equal(got, want), int((got != want).sum())
    assert 0 < want.sum() < len(want)
    sc.free()


class HostDevice:
    """'Device memory' of the emulated tier is host memory.  put/empty return (handle, pointer);
    get copies a handle back as a numpy array of `dtype`."""
    stream = None

    def put(self, arr):
        a = np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()
        return a, a.ctypes.data

    def empty(self, nbytes, fill=0):
        a = np.full(max(nbytes, 16), fill, dtype=np.uint8)
        return a, a.ctypes.data

    def get(self, handle, dtype, count):
        return handle[:count * np.dtype(dtype).itemsize].copy().view(dtype)

    def sync(self):
        pass


class TorchDevice:
    """device buffers for the shared cases on a real GPU: torch owns the memory, the library sees raw pointers"""
    stream = None

    def put(self, arr):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).cuda()
        return t, t.data_ptr()

    def empty(self, nbytes, fill=0):
        import torch
        t = torch.full((max(nbytes, 16),), fill, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        return t, t.data_ptr()

    def get(self, handle, dtype, count):
        import torch
        torch.cuda.synchronize(handle.device)
        return handle[:count * np.dtype(dtype).itemsize].cpu().numpy().view(dtype)

    def sync(self):
        import torch
        torch.cuda.synchronize()

    def on_device(self, k):
        import torch
        return torch.cuda.device(k)


def case_wavefront(lib, orc, dev):
    """primary rays and bounce rays generated on the device, traced without leaving
    it.  Generators against the numpy restatement (oracle/wavefront_ref.py): primary rays bit for
    bit, bounce rays within 1e-5; every traced bounce against the brute-force oracle on the very
    rays the device produced (bit-exact)."""
    from oracle import wavefront_ref as wf
    s = scenes.config_scene("C1")
    sc = lib.build_scene(s["meshes"])
    info = sc.info()
    W, H = 48, 40
    cam = api.rtk_cuda_camera()
    eye, fwd, right, up, tan = (278.0, 273.0, -800.0), (0.0, 0.0, 1.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), 0.36
    cam.eye[:], cam.forward[:], cam.right[:], cam.up[:] = eye, fwd, right, up
    cam.tan_half_fov, cam.width, cam.height = tan, W, H
    seed, n = 0xD5, W * H - 7
    first_pixel = 5
    h_rays, d_rays = dev.empty(32 * n)
    h_next, d_next = dev.empty(32 * n)
    h_hit, d_hit = dev.empty(16 * n)
    h_alive, d_alive = dev.empty(n, fill=9)
    assert lib.rtk_cuda_generate_primary_rays(C.byref(cam), seed, 3, first_pixel, n, d_rays, dev.stream) == 0, lib.last_error()
    rays = dev.get(h_rays, api.RAY_DTYPE, n)
    want = wf.primary_rays(eye, fwd, right, up, tan, W, H, seed, 3, first_pixel, n)
    assert rays.tobytes() == want.tobytes(), "primary rays differ from the restatement"
    # argument checks: pixel range outside the frame
    assert lib.rtk_cuda_generate_primary_rays(C.byref(cam), seed, 3, W * H - 3, 8, d_rays, dev.stream) != 0
    assert lib.rtk_cuda_generate_primary_rays(C.byref(cam), seed, 64, 0, 8, d_rays, dev.stream) != 0

    amax = np.float32(max(np.abs(np.array(info.bounds_min[:])).max(), np.abs(np.array(info.bounds_max[:])).max()))
    push = np.float32(amax * np.float32(2.0 ** -13))
    tris = s["tris"]
    total_hits = 0
    for bounce in range(4):
        flags = api.RTK_CUDA_BOUNCE_RELAUNCH if bounce % 2 else 0
        assert lib.rtk_trace_rays_compact_device(sc.ptr, d_rays, d_hit, n, dev.stream) == 0, lib.last_error()
        rays = dev.get(h_rays, api.RAY_DTYPE, n)
        hit = dev.get(h_hit, api.HIT16_DTYPE, n)
        assert_same(hit, orc.trace_brute(tris, rays), f"bounce {bounce}")
        total_hits += int((hit["prim"] != api.RTK_CUDA_MISS).sum())
        assert lib.rtk_cuda_generate_bounce_rays(sc.ptr, d_rays, d_hit, d_next, d_alive, n, seed, bounce, 1000, flags, dev.stream) == 0, lib.last_error()
        got = dev.get(h_next, api.RAY_DTYPE, n)
        alive = dev.get(h_alive, np.uint8, n)
        ref, state = wf.bounce_rays(tris, rays, hit, seed, bounce, 1000, push, flags)
        assert np.array_equal(alive, state)
        if flags:
            assert (alive != 0).all()
        else:
            dead = alive == 0
            assert np.array_equal(dead, hit["prim"] == api.RTK_CUDA_MISS)
            assert (got["max_t"][dead] == 0).all() and got[dead].tobytes() == ref[dead].tobytes()
        live = alive != 0
        scale = float(amax)
        assert np.allclose(got["o"][live], ref["o"][live], rtol=1e-5, atol=1e-5 * scale)
        assert np.allclose(got["d"][live], ref["d"][live], rtol=1e-5, atol=2e-5)
        assert (got["max_t"][live] == api.RTK_INF).all() and (got["min_t"] == 0).all()
        # directions are unit length and leave the surface on the side the path arrived from
        assert np.allclose(np.linalg.norm(got["d"][live].astype(np.float64), axis=1), 1.0, atol=1e-4)
        d_rays, d_next = d_next, d_rays
        h_rays, h_next = h_next, h_rays
    assert total_hits > n          # the Cornell box is closed on five sides: most paths keep hitting
    assert lib.rtk_cuda_generate_bounce_rays(sc.ptr, d_rays, d_hit, d_next, d_alive, n, seed, 16, 0, 0, dev.stream) != 0
    sc.free()


def case_refit(lib, orc, dev):
    """deform the vertices of a built scene, refit (or rebuild) on the device, and get
    exactly the hits of the oracle on the deformed triangles.  `dev` owns the device buffers."""
    s = scenes.config_scene("C3", 0.006)
    m = s["meshes"][0]
    pos0, idx = m["positions"].astype(np.float32), m["indices"].astype(np.uint32)
    rays = scenes.bounce_rays(s, 2500)
    keep, meshes = [], (api.rtk_cuda_mesh * 1)()

    def upload(pos):
        hp, dp = dev.put(pos)
        hi, di = dev.put(idx)
        keep[:] = [hp, hi]
        meshes[0].d_positions, meshes[0].d_indices = dp, di
        meshes[0].num_vertices, meshes[0].num_triangles = len(pos), len(idx)
    upload(pos0)
    ptr = lib.rtk_cuda_build_scene(meshes, 1, dev.stream)
    assert ptr, lib.last_error()
    sc = api.Scene(lib, ptr)
    h_rays, d_rays = dev.put(rays)
    h_hit, d_hit = dev.empty(16 * len(rays))

    def trace():
        assert lib.rtk_trace_rays_compact_device(sc.ptr, d_rays, d_hit, len(rays), dev.stream) == 0, lib.last_error()
        return dev.get(h_hit, api.HIT16_DTYPE, len(rays))
    assert_same(trace(), orc.trace_brute(pos0[idx.astype(np.int64)], rays), "before the update")
    rng = np.random.default_rng(17)
    for step, mode in enumerate([api.RTK_CUDA_UPDATE_REFIT, api.RTK_CUDA_UPDATE_REFIT, api.RTK_CUDA_UPDATE_REBUILD, api.RTK_CUDA_UPDATE_REFIT]):
        pos = pos0.copy()
        pos[:, 1] += (0.02 * (step + 1) * np.sin(9.0 * pos0[:, 0] + step) * np.cos(7.0 * pos0[:, 2])).astype(np.float32)
        pos += (rng.random(pos.shape).astype(np.float32) - 0.5) * np.float32(2e-3)
        if step == 1:
            pos *= np.float32(3.0)                      # the scene bounds change as well
        upload(pos)
        assert lib.rtk_cuda_update_scene(sc.ptr, meshes, 1, mode, dev.stream) == 0, lib.last_error()
        want = orc.trace_brute(pos[idx.astype(np.int64)], rays)
        assert_same(trace(), want, f"update {step} mode {mode}")
        assert (want["prim"] != api.RTK_CUDA_MISS).any()
    # a different triangle count is refused
    meshes[0].num_triangles = len(idx) - 1
    assert lib.rtk_cuda_update_scene(sc.ptr, meshes, 1, api.RTK_CUDA_UPDATE_REFIT, dev.stream) != 0
    sc.free()


def filtered_oracle(orc, tris, rays, keep):
    """the oracle on the scene WITHOUT the triangles that are switched off, numbered as before"""
    kept = np.flatnonzero(keep).astype(np.uint32)
    want = orc.trace_brute(np.ascontiguousarray(tris[kept]), rays) if len(kept) else None
    if want is None:
        want = np.zeros(len(rays), dtype=api.HIT16_DTYPE)
        want["prim"] = api.RTK_CUDA_MISS
        return want
    hit = want["prim"] != api.RTK_CUDA_MISS
    want["prim"][hit] = kept[want["prim"][hit]]
    return want


def case_triangle_filter(lib, orc, dev):
    """the device-side triangle predicate (bitset over global triangle numbers).
    With a filter every entry point must answer exactly as the oracle does on the scene without the
    switched-off triangles (numbering unchanged): compact device trace, exhaustive kernel, occlusion
    query, host batch; the filter survives refits and rebuilds; removing it restores the scene."""
    s = scenes.config_scene("C4", 0.0012)                       # two meshes: global numbering matters
    tris, first = s["tris"], s["mesh_first"]
    n = len(tris)
    rays = scenes.mixed_rays(s, 2400, block=256)
    sc = lib.build_scene(s["meshes"])
    h_rays, d_rays = dev.put(rays)
    h_hit, d_hit = dev.empty(16 * len(rays))
    h_occ, d_occ = dev.empty(len(rays), fill=7)

    def trace(brute=False):
        fn = lib.rtk_trace_rays_bruteforce_device if brute else lib.rtk_trace_rays_compact_device
        assert fn(sc.ptr, d_rays, d_hit, len(rays), dev.stream) == 0, lib.last_error()
        return dev.get(h_hit, api.HIT16_DTYPE, len(rays))

    def check(keep, what):
        want = filtered_oracle(orc, tris, rays, keep)
        assert_same(trace(), want, what + ": traversal")
        assert_same(trace(brute=True), want, what + ": exhaustive kernel")
        assert lib.rtk_occluded_rays_device(sc.ptr, d_rays, d_occ, len(rays), dev.stream) == 0, lib.last_error()
        occ = dev.get(h_occ, np.uint8, len(rays)).astype(bool)
        assert np.array_equal(occ, want["prim"] != api.RTK_CUDA_MISS), what + ": occlusion query"
        hits, mask, nh = sc.trace_rays(rays)
        assert_same(api.hits_to_hit16(hits, mask, first), want, what + ": host batch")
        m = mask.astype(bool)
        if m.any():                                              # expanded rows still carry per-mesh numbering
            g = want["prim"][m].astype(np.int64)
            mi = np.searchsorted(np.asarray(first), g, side="right") - 1
            assert np.array_equal(hits["mesh_index"][m], mi) and np.array_equal(hits["triangle_index"][m], g - np.asarray(first)[mi])
        return want

    everything = np.ones(n, dtype=bool)
    base = check(everything, "no filter")
    assert 0 < (base["prim"] != api.RTK_CUDA_MISS).sum() < len(rays)
    rng = np.random.default_rng(41)
    half = rng.random(n) < 0.5
    sc.set_triangle_filter(half)
    w = check(half, "random half")
    assert (w["prim"] != base["prim"]).any()
    # depth peeling: switch off exactly the triangles the rays hit first and look behind them
    peel = everything.copy()
    peel[base["prim"][base["prim"] != api.RTK_CUDA_MISS]] = False
    h_bits, d_bits = dev.put(api.pack_triangle_filter(peel))
    assert lib.rtk_cuda_set_triangle_filter_device(sc.ptr, d_bits, (n + 31) // 32, dev.stream) == 0, lib.last_error()
    w = check(peel, "peeled (device bitset)")
    hit = (w["prim"] != api.RTK_CUDA_MISS) & (base["prim"] != api.RTK_CUDA_MISS)
    assert hit.any() and (w["t"][hit] >= base["t"][hit]).all()
    # one whole mesh off, then everything off
    only1 = everything.copy()
    only1[:first[1]] = False
    sc.set_triangle_filter(only1)
    w = check(only1, "mesh 0 off")
    assert (w["prim"][w["prim"] != api.RTK_CUDA_MISS] >= first[1]).all()
    sc.set_triangle_filter(~everything)
    check(~everything, "all off")
    # a bitset that is too short is refused and changes nothing
    short = np.zeros(1, dtype=np.uint32)
    assert lib.rtk_cuda_set_triangle_filter(sc.ptr, short.ctypes.data, (n + 31) // 32 - 1) != 0
    check(~everything, "after the refused call")
    # the filter outlives a rebuild; removing it gives the unfiltered scene back
    sc.set_triangle_filter(half)
    assert lib.rtk_cuda_rebuild_scene(sc.ptr, dev.stream) == 0, lib.last_error()
    check(half, "after a rebuild")
    sc.set_triangle_filter(None)
    assert_same(trace(), base, "filter removed")
    sc.set_triangle_filter(None)                                 # removing twice is fine
    sc.free()

    # ... and a refit: deformed device mesh, filter in place
    s = scenes.config_scene("C3", 0.005)
    m = s["meshes"][0]
    pos0, idx = m["positions"].astype(np.float32), m["indices"].astype(np.uint32)
    rays = scenes.bounce_rays(s, 1500)
    meshes = (api.rtk_cuda_mesh * 1)()
    hp, dp = dev.put(pos0)
    hi, di = dev.put(idx)
    meshes[0].d_positions, meshes[0].d_indices = dp, di
    meshes[0].num_vertices, meshes[0].num_triangles = len(pos0), len(idx)
    ptr = lib.rtk_cuda_build_scene(meshes, 1, dev.stream)
    assert ptr, lib.last_error()
    sc = api.Scene(lib, ptr)
    keep = np.random.default_rng(43).random(len(idx)) < 0.7
    sc.set_triangle_filter(keep)
    h_rays, d_rays = dev.put(rays)
    h_hit, d_hit = dev.empty(16 * len(rays))
    for step, mode in enumerate([api.RTK_CUDA_UPDATE_REFIT, api.RTK_CUDA_UPDATE_REBUILD, api.RTK_CUDA_UPDATE_REFIT]):
        pos = pos0.copy()
        pos[:, 1] += (0.03 * (step + 1) * np.sin(8.0 * pos0[:, 0] + step) * np.cos(6.0 * pos0[:, 2])).astype(np.float32)
        hp2, dp2 = dev.put(pos)
        meshes[0].d_positions = dp2
        assert lib.rtk_cuda_update_scene(sc.ptr, meshes, 1, mode, dev.stream) == 0, lib.last_error()
        assert lib.rtk_trace_rays_compact_device(sc.ptr, d_rays, d_hit, len(rays), dev.stream) == 0, lib.last_error()
        got = dev.get(h_hit, api.HIT16_DTYPE, len(rays))
        want = filtered_oracle(orc, pos[idx.astype(np.int64)], rays, keep)
        assert_same(got, want, f"filtered, update {step} mode {mode}")
        assert (want["prim"] != api.RTK_CUDA_MISS).any()
    sc.set_triangle_filter(None)
    assert lib.rtk_trace_rays_compact_device(sc.ptr, d_rays, d_hit, len(rays), dev.stream) == 0, lib.last_error()
    assert_same(dev.get(h_hit, api.HIT16_DTYPE, len(rays)), orc.trace_brute(pos[idx.astype(np.int64)], rays), "filter removed after refits")
    sc.free()


def bake_instance(pos, xf):
    """the library's instance transform restated in plain fp32: ((m0*x + m1*y) + m2*z) + m3 per row"""
    pos = np.asarray(pos, dtype=np.float32)
    m = np.asarray(xf, dtype=np.float32).reshape(3, 4)
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    return np.stack([((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3] for r in range(3)], axis=1).astype(np.float32)


def case_instancing(lib, orc, dev):
    """SURVEY 8(f) N4, instancing (baked): instances of two meshes under rigid, scaled and mirrored
    transforms; hits must be the oracle's on the transformed triangles, mesh_index the instance
    number, triangle_index the triangle within its mesh, vertices in world space; instances then
    move (refit and rebuild)."""
    a = scenes.config_scene("C3", 0.002)["meshes"][0]                       # a terrain patch
    b = scenes.config_scene("C1")["meshes"][0]                               # the Cornell box
    src = []
    for m in (a, b):
        pos = m["positions"].astype(np.float32)
        pos = (pos - pos.min(0)) / np.float32((pos.max(0) - pos.min(0)).max())     # unit-sized
        src.append((np.ascontiguousarray(pos), m["indices"].astype(np.uint32)))
    meshes, keep = (api.rtk_cuda_mesh * 2)(), []
    for i, (pos, idx) in enumerate(src):
        hp, dp = dev.put(pos)
        hi, di = dev.put(idx)
        keep += [hp, hi]
        meshes[i].d_positions, meshes[i].d_indices = dp, di
        meshes[i].num_vertices, meshes[i].num_triangles = len(pos), len(idx)

    def rot_y(t):
        c, s_ = np.cos(t), np.sin(t)
        return np.array([[c, 0, s_], [0, 1, 0], [-s_, 0, c]])

    def layout(phase):
        out = []
        for i in range(7):
            lin = rot_y(0.7 * i + phase) * (0.6 + 0.15 * i)
            if i == 3:
                lin = lin @ np.diag([1.0, -1.0, 1.0])                       # mirrored
            if i == 5:
                lin = lin @ np.diag([2.0, 0.5, 1.0])                        # non-uniform scale
            t = np.array([1.7 * (i % 3) + 0.3 * phase, 0.4 * (i // 3), 1.9 * (i // 3) - 0.2 * phase])
            out.append((i % 2, np.concatenate([lin, t[:, None]], axis=1).astype(np.float32)))
        return out

    def pack(lay):
        inst = (api.rtk_cuda_instance * len(lay))()
        for i, (mi, xf) in enumerate(lay):
            inst[i].mesh = mi
            inst[i].transform[:] = [float(v) for v in xf.reshape(-1)]
        return inst

    def world(lay):
        tris, first = [], [0]
        for mi, xf in lay:
            pos, idx = src[mi]
            tris.append(bake_instance(pos, xf)[idx.astype(np.int64)])
            first.append(first[-1] + len(idx))
        return np.ascontiguousarray(np.concatenate(tris)), np.array(first)

    lay = layout(0.0)
    ptr = lib.rtk_cuda_build_instanced_scene(meshes, 2, pack(lay), len(lay), dev.stream)
    assert ptr, lib.last_error()
    sc = api.Scene(lib, ptr)
    info = sc.info()
    tris, first = world(lay)
    assert info.num_triangles == len(tris) and info.num_meshes == len(lay)
    rays = scenes.bounce_rays({"tris": tris}, 2500)
    rays["o"][::2] += np.float32(0.5) * rays["d"][::2]                       # some start further out
    rays["d"][1::2] *= np.float32(-1.0)
    for step, mode in enumerate([None, api.RTK_CUDA_UPDATE_REFIT, api.RTK_CUDA_UPDATE_REBUILD]):
        if mode is not None:
            lay = layout(0.35 * step)
            assert lib.rtk_cuda_update_instanced_scene(sc.ptr, meshes, 2, pack(lay), len(lay), mode, dev.stream) == 0, lib.last_error()
            tris, first = world(lay)
        want = orc.trace_brute(tris, rays)
        hits, mask, nh = sc.trace_rays(rays)
        assert_same(api.hits_to_hit16(hits, mask, first), want, f"instanced scene, step {step}")
        m = mask.astype(bool)
        assert 50 < nh < len(rays)
        g = want["prim"][m].astype(np.int64)
        ii = np.searchsorted(first, g, side="right") - 1
        assert np.array_equal(hits["mesh_index"][m], ii), "mesh_index is the instance number"
        assert np.array_equal(hits["triangle_index"][m], g - first[ii])
        assert len(np.unique(ii)) >= 4
        assert hits["vertex"]["position"][m].tobytes() == tris[g].tobytes(), "vertices are in world space"
        local_idx = np.concatenate([src[mi][1] for mi, _ in lay])[g]
        assert np.array_equal(hits["vertex"]["index"][m], local_idx), "original vertex indices of the mesh"
    # refused: unknown mesh, wrong instance count, non-finite transform
    bad = pack(lay)
    bad[2].mesh = 9
    assert not lib.rtk_cuda_build_instanced_scene(meshes, 2, bad, len(lay), dev.stream)
    assert lib.rtk_cuda_update_instanced_scene(sc.ptr, meshes, 2, pack(lay), len(lay) - 1, api.RTK_CUDA_UPDATE_REFIT, dev.stream) != 0
    bad = pack(lay)
    bad[0].transform[5] = float("nan")
    assert lib.rtk_cuda_update_instanced_scene(sc.ptr, meshes, 2, bad, len(lay), api.RTK_CUDA_UPDATE_REFIT, dev.stream) != 0
    swapped = pack(lay)
    swapped[0].mesh, swapped[1].mesh = 1, 0                                  # different triangle counts
    assert lib.rtk_cuda_update_instanced_scene(sc.ptr, meshes, 2, swapped, len(lay), api.RTK_CUDA_UPDATE_REFIT, dev.stream) != 0
    sc.free()


def case_pathological(lib, orc):
    """scenes a builder can trip over: zero-extent bounds (all points / all collinear / all coplanar),
    one triangle that dwarfs the rest, coordinates whose products overflow or underflow fp32, signed
    zeros, two clusters a million units apart -- both builders, hits bit-exact as everywhere else"""
    rng = np.random.default_rng(5)
    n = 500

    def rays_for(tris, count=250):
        flat = tris.reshape(-1, 3).astype(np.float64)
        lo, hi = flat.min(0), flat.max(0)
        ext = np.maximum(hi - lo, max(np.abs(hi).max() * 1e-3, 1e-3))
        r = np.zeros(count, dtype=api.RAY_DTYPE)
        o = lo + (rng.random((count, 3)) * 3 - 1) * ext
        tgt = tris[rng.integers(0, len(tris), count)].astype(np.float64).mean(1)
        d = (tgt - o).astype(np.float32)
        d[np.abs(d).max(1) == 0] = (0, 0, 1)
        r["o"], r["d"], r["max_t"] = o.astype(np.float32), d, api.RTK_INF
        return r
    cases = {}
    cases["all points at one place"] = np.repeat(np.repeat(rng.random((1, 1, 3)), 3, 1), n, 0)
    line = np.zeros((n, 3, 3))
    line[:, :, 0] = rng.random((n, 3))
    cases["all collinear on the x axis"] = line
    plane = rng.random((n, 3, 3))
    plane[:, :, 1] = 0.25
    cases["all in one plane"] = plane
    big = rng.random((n, 1, 3)) + 0.01 * rng.random((n, 3, 3))
    big[0] = [(-50, -50, 0.5), (50, -50, 0.5), (0, 80, 0.5)]
    cases["one huge triangle"] = big
    cases["1e15 coordinates"] = rng.random((n, 3, 3)) * 1e15
    cases["1e-20 coordinates"] = rng.random((n, 3, 3)) * 1e-20
    neg = rng.random((n, 1, 3)) - 0.5 + 0.05 * rng.random((n, 3, 3))
    neg[::3, :, 2] = -0.0
    cases["signed zeros"] = neg
    cases["two far clusters"] = np.concatenate([rng.random((n // 2, 3, 3)), rng.random((n // 2, 3, 3)) + 1e6])
    # the same kind of trouble above the size one CTA finishes on its own (512 triangles): the level-by-level part of
    # the SAH builder with no valid split at all (position halves), with splits that peel a sliver off a geometric
    # progression level after level (unbalanced node lists, one-chunk and many-chunk nodes side by side), and with
    # two clusters that separate at the root
    m = 3000
    cases["3000 triangles at one place"] = np.repeat(np.repeat(rng.random((1, 1, 3)), 3, 1), m, 0)
    prog = np.zeros((m, 3, 3))
    x = 1.01 ** np.arange(m)
    prog[:, :, 0] = x[:, None] * (1 + 0.004 * rng.random((m, 3)))
    prog[:, :, 1:] = rng.random((m, 3, 2)) * x[:, None, None] * 0.01
    cases["geometric progression along x"] = prog
    cases["two far clusters of 1500"] = np.concatenate([rng.random((m // 2, 3, 3)), rng.random((m // 2, 3, 3)) + 1e6])
    total = 0
    for mode in (api.RTK_CUDA_BUILD_LBVH, api.RTK_CUDA_BUILD_SAH):
        for name, tris in cases.items():
            tris = np.ascontiguousarray(tris.astype(np.float32))
            r = rays_for(tris)
            got, _, _ = trace_hit16(lib, soup_mesh(tris), r, mode=mode)
            want = orc.trace_brute(tris, r)
            assert_same(got, want, f"{name} (build mode {mode})")
            total += int((want["prim"] != api.RTK_CUDA_MISS).sum())
    assert total > 1000
    lib.rtk_cuda_set_build_mode(api.RTK_CUDA_BUILD_SAH)


def case_deep_stack(lib, orc, dev=None):
    """Thousands of coincident triangles: the builder cannot separate them (forced halving,
    rtk.c:1429-1443), every box overlaps every other, so a ray has to visit all of them: the
    traversal stack outgrows its shared-memory part and spills to the global slab.  All hits tie
    exactly and the lowest triangle number must win."""
    base = np.array([[(0, 0, 1), (1, 0, 1), (0, 1, 1)]], dtype=np.float32)
    tris = np.repeat(base, 21500, axis=0)
    tris[20000:] += np.float32(0.25)                   # a second coincident cluster, further away
    rays = np.zeros(64, dtype=api.RAY_DTYPE)
    rng = np.random.default_rng(23)
    rays["o"] = np.concatenate([rng.random((64, 2)) * 0.6, np.zeros((64, 1))], axis=1).astype(np.float32)
    rays["d"] = (0, 0, 1)
    rays["max_t"] = api.RTK_INF
    got, _, _ = trace_hit16(lib, soup_mesh(tris), rays)
    want = orc.trace_brute(tris, rays)
    assert_same(got, want, "coincident triangles")
    hit = want["prim"] != api.RTK_CUDA_MISS
    assert hit.any() and np.isin(want["prim"][hit], (0, 20000)).all() and (want["prim"] == 0).any()
    # the statistics variant reports how deep the stack went
    sc = lib.build_scene(soup_mesh(tris))
    info = sc.info()
    assert info.num_wide_nodes > 8
    if dev is not None:
        h_rays, d_rays = dev.put(rays)
        h_hit, d_hit = dev.empty(16 * len(rays))
        st = api.rtk_cuda_trace_stats()
        assert lib.rtk_trace_stats_device(sc.ptr, d_rays, d_hit, len(rays), C.byref(st), dev.stream) == 0, lib.last_error()
        assert_same(dev.get(h_hit, api.HIT16_DTYPE, len(rays)), want, "statistics kernel")
        assert st.stack_max > 16, st.stack_max                 # beyond the shared-memory part (RTK_STACK_SMEM)
        assert st.tri_tests >= 1500 * int(hit.sum())
    sc.free()


def case_host_batch_chunks(lib, orc, nrays=30000, chunk_log2=12):
    """rtk_trace_rays over many chunks: the pipeline's buffer rotation (4 chunks in flight, upload
    stream ahead of the kernels, dense rows put back in place by the host threads) must give the
    rows of a one-chunk run, leave the rows of misses untouched, and agree with the oracle."""
    s = scenes.config_scene("C3", 0.004)
    rays = scenes.bounce_rays(s, nrays)
    sc = lib.build_scene(s["meshes"])
    old = os.environ.pop("RTK_B200_HOST_CHUNK_LOG2", None)
    try:
        # a small pageable batch first: the staging of the host path (bounce buffers of the upload included) is
        # sized by the chunk and has to grow with the next, larger batch
        small = min(nrays, 4500)
        _, m0, n0 = sc.trace_rays(np.ascontiguousarray(rays[:small]))
        hits1 = np.zeros(nrays, dtype=api.HIT_DTYPE)
        hits1.view(np.uint8)[:] = 0x5A
        _, mask1, n1 = sc.trace_rays(rays, hits=hits1)
        assert np.array_equal(m0, mask1[:small]) and n0 == int(mask1[:small].sum())
        os.environ["RTK_B200_HOST_CHUNK_LOG2"] = str(chunk_log2)
        hits2 = np.zeros(nrays, dtype=api.HIT_DTYPE)
        hits2.view(np.uint8)[:] = 0x5A
        _, mask2, n2 = sc.trace_rays(rays, hits=hits2)
        # no mask requested: rows still land where they belong
        hits3 = np.zeros(nrays, dtype=api.HIT_DTYPE)
        r = lib.rtk_trace_rays(sc.ptr, rays.ctypes.data, hits3.ctypes.data, None, nrays)
    finally:
        os.environ.pop("RTK_B200_HOST_CHUNK_LOG2", None)
        if old is not None:
            os.environ["RTK_B200_HOST_CHUNK_LOG2"] = old
    assert n1 == n2 == r and 0 < n1 < nrays
    assert np.array_equal(mask1, mask2)
    assert hits1.tobytes() == hits2.tobytes()
    m = mask1.astype(bool)
    assert (hits2.view(np.uint8).reshape(-1, 68)[~m] == 0x5A).all()          # misses untouched (rtk.c:571-576)
    assert hits3[m].tobytes() == hits2[m].tobytes()
    k = min(nrays, 1500)
    assert_same(api.hits_to_hit16(hits2, mask2, s["mesh_first"])[:k], orc.trace_brute(s["tris"], rays[:k]), "chunked host batch")
    # compact results (rtk_trace_rays_compact): a record for every ray, one chunk and many chunks
    full = api.hits_to_hit16(hits2, mask2, s["mesh_first"])
    c1 = sc.trace_rays_compact(rays)
    os.environ["RTK_B200_HOST_CHUNK_LOG2"] = str(chunk_log2)
    try:
        c2 = sc.trace_rays_compact(rays, out=np.full(nrays, 0x5A5A5A5A, dtype=np.uint32).repeat(4).view(api.HIT16_DTYPE))
        c3 = sc.trace_rays_compact(rays[:nrays - 37])                           # ragged last chunk
    finally:
        os.environ.pop("RTK_B200_HOST_CHUNK_LOG2", None)
        if old is not None:
            os.environ["RTK_B200_HOST_CHUNK_LOG2"] = old
    assert_same(c1, full, "compact host batch")
    assert_same(c2, full, "compact host batch, many chunks")
    assert_same(c3, full[:nrays - 37], "compact host batch, ragged")
    assert len(sc.trace_rays_compact(rays[:0])) == 0
    assert lib.rtk_trace_rays_compact(sc.ptr, None, None, 5) != 0
    sc.free()


# ---------------------------------------------------------------------------------------------
# round 2: page-locked caller arrays (rows written in place by the device), several devices behind
# one process, concurrent streams on one scene, blob validation, the overflow report
# ---------------------------------------------------------------------------------------------

class PinnedArrays:
    """numpy views of rtk_cuda_host_alloc memory (page-locked, device-writable)"""

    def __init__(self, lib):
        self.lib, self.ptrs = lib, []

    def empty(self, n, dtype, fill=None):
        dt = np.dtype(dtype)
        nbytes = max(n * dt.itemsize, 16)
        p = self.lib.rtk_cuda_host_alloc(nbytes)
        assert p, self.lib.last_error()
        self.ptrs.append(p)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))[:n * dt.itemsize].view(dt)
        if fill is not None:
            a.view(np.uint8)[:] = fill
        return a

    def free(self):
        for p in self.ptrs:
            self.lib.rtk_cuda_host_free(p)
        self.ptrs = []


def case_direct_rows(lib, orc, nrays=30000, chunk_log2=12):
    """rtk_trace_rays with page-locked hits / mask arrays: the resolve kernel writes the rows of the rays
    that hit straight into the caller's memory.  Must equal the staged path (pageable arrays) byte for
    byte, leave the rows of misses untouched (rtk.c:571-576), survive many chunks and ragged tails, and
    agree with the oracle."""
    s = scenes.config_scene("C3", 0.004)
    rays = scenes.bounce_rays(s, nrays)
    sc = lib.build_scene(s["meshes"])
    pin = PinnedArrays(lib)
    old = os.environ.pop("RTK_B200_HOST_CHUNK_LOG2", None)
    try:
        hits0 = np.zeros(nrays, dtype=api.HIT_DTYPE)
        hits0.view(np.uint8)[:] = 0x5A
        _, mask0, n0 = sc.trace_rays(rays, hits=hits0)                       # pageable: staged path
        p_rays = pin.empty(nrays, api.RAY_DTYPE)
        p_rays[:] = rays
        for log2 in (None, chunk_log2):
            if log2 is not None:
                os.environ["RTK_B200_HOST_CHUNK_LOG2"] = str(log2)
            p_hits = pin.empty(nrays, api.HIT_DTYPE, fill=0x5A)
            p_mask = pin.empty(nrays, np.uint8, fill=0x77)
            r = lib.rtk_trace_rays(sc.ptr, p_rays.ctypes.data, p_hits.ctypes.data, p_mask.ctypes.data, nrays)
            assert r == n0, (r, n0, lib.last_error())
            assert np.array_equal(p_mask, mask0)
            assert p_hits.tobytes() == hits0.tobytes(), "direct rows differ from the staged rows"
            # no mask requested, pageable rays, ragged count
            p_hits2 = pin.empty(nrays, api.HIT_DTYPE, fill=0x5A)
            k = nrays - 37
            r = lib.rtk_trace_rays(sc.ptr, rays.ctypes.data, p_hits2.ctypes.data, None, k)
            assert r == int(mask0[:k].sum())
            assert p_hits2[:k].tobytes() == hits0[:k].tobytes()
            assert (p_hits2[k:].view(np.uint8) == 0x5A).all()
        m = mask0.astype(bool)
        assert 0 < m.sum() < nrays
        assert (hits0.view(np.uint8).reshape(-1, 68)[~m] == 0x5A).all()
        kk = min(nrays, 1500)
        assert_same(api.hits_to_hit16(p_hits, p_mask, s["mesh_first"])[:kk], orc.trace_brute(s["tris"], rays[:kk]), "direct rows")
    finally:
        os.environ.pop("RTK_B200_HOST_CHUNK_LOG2", None)
        if old is not None:
            os.environ["RTK_B200_HOST_CHUNK_LOG2"] = old
        pin.free()
        sc.free()


def case_multi_device(lib, orc, ndev, dev=None, nrays=600000, oracle_rays=1200):
    """`lib` was initialised with rtk_cuda_init_devices over `ndev` devices.  One rtk_build_scene, one
    rtk_trace_rays: the batch is split over the devices and the result must be what a single device
    gives -- rows, mask, count, compact records -- and the oracle's on a sample.  Small batches, pageable
    and page-locked arrays, the triangle filter and a rebuild all have to reach every replica."""
    assert lib.rtk_cuda_device_count() == ndev
    s = scenes.config_scene("C3", 0.004)
    rays = scenes.bounce_rays(s, nrays)
    sc = lib.build_scene(s["meshes"])
    pin = PinnedArrays(lib)
    try:
        # reference result: compact records of the whole batch through the device entry point on the
        # first device (single-device code path), expanded on the host side of the test
        want16 = sc.trace_rays_compact(rays)                   # split over the devices as well ...
        idx = np.linspace(0, nrays - 1, oracle_rays).astype(np.int64)
        assert_same(want16[idx], orc.trace_brute(s["tris"], rays[idx]), "multi-device compact records")
        # every share's boundary region against the oracle too (off-by-one in the range split)
        per = (nrays // ndev + 127) // 128 * 128
        edges = np.unique(np.clip(np.concatenate([np.arange(k * per - 3, k * per + 3) for k in range(1, ndev)] + [np.arange(nrays - 4, nrays)]), 0, nrays - 1))
        assert_same(want16[edges], orc.trace_brute(s["tris"], rays[edges]), "multi-device share boundaries")
        # rows, pageable arrays (staged path on every device)
        hits = np.zeros(nrays, dtype=api.HIT_DTYPE)
        hits.view(np.uint8)[:] = 0x5A
        _, mask, nh = sc.trace_rays(rays, hits=hits)
        assert nh == int(mask.sum()) == int((want16["prim"] != api.RTK_CUDA_MISS).sum())
        assert_same(api.hits_to_hit16(hits, mask, s["mesh_first"]), want16, "multi-device rows (staged)")
        assert (hits.view(np.uint8).reshape(-1, 68)[~mask.astype(bool)] == 0x5A).all()
        # rows, page-locked arrays (written in place by every device)
        p_rays = pin.empty(nrays, api.RAY_DTYPE)
        p_rays[:] = rays
        p_hits = pin.empty(nrays, api.HIT_DTYPE, fill=0x5A)
        p_mask = pin.empty(nrays, np.uint8, fill=0x77)
        r = lib.rtk_trace_rays(sc.ptr, p_rays.ctypes.data, p_hits.ctypes.data, p_mask.ctypes.data, nrays)
        assert r == nh, (r, nh, lib.last_error())
        assert p_hits.tobytes() == hits.tobytes() and np.array_equal(p_mask, mask)
        # single rays from several threads land on different devices
        import threading
        errors = []

        def single(tid):
            try:
                for i in range(tid, 64, 4):
                    h = sc.trace_ray(rays[i])
                    if want16["prim"][i] == api.RTK_CUDA_MISS:
                        assert h is None
                    else:
                        assert h is not None and h["t"] == want16["t"][i]
            except Exception as ex:                                   # noqa: BLE001
                errors.append(ex)
        th = [threading.Thread(target=single, args=(t,)) for t in range(4)]
        [t.start() for t in th]
        [t.join() for t in th]
        if errors:
            raise errors[0]
        # a filter and a rebuild must reach the replicas
        keep = np.ones(len(s["tris"]), dtype=bool)
        keep[want16["prim"][want16["prim"] != api.RTK_CUDA_MISS]] = False       # switch off every first hit
        sc.set_triangle_filter(keep)
        got = sc.trace_rays_compact(rays)
        hitm = got["prim"] != api.RTK_CUDA_MISS
        assert not np.isin(got["prim"][hitm], np.nonzero(~keep)[0]).any(), "a replica still traces switched-off triangles"
        k = 400
        tr = s["tris"].copy()
        tr[~keep] = np.nan
        assert_same(got[idx[:k]], orc.trace_brute(tr, rays[idx[:k]]), "filtered, multi-device")
        sc.set_triangle_filter(None)
        assert lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0, lib.last_error()
        assert_same(sc.trace_rays_compact(rays), want16, "after filter removal and rebuild")
        assert lib.rtk_cuda_scene_status(sc.ptr) == 0
        # device entry points follow the buffer: a buffer on device k is traced by device k's replica
        if dev is not None and hasattr(dev, "on_device"):
            for k in range(ndev):
                with dev.on_device(k):
                    h_r, d_r = dev.put(rays[:5000])
                    h_o, d_o = dev.empty(16 * 5000)
                    assert lib.rtk_trace_rays_compact_device(sc.ptr, d_r, d_o, 5000, dev.stream) == 0, lib.last_error()
                    assert_same(dev.get(h_o, api.HIT16_DTYPE, 5000), want16[:5000], f"device entry point on device {k}")
    finally:
        pin.free()
        sc.free()


def case_two_streams(lib, orc, dev, make_stream=None):
    """Two queries on ONE scene in flight at once on different streams (and from two host threads): each
    launch has its own ray cursor and stack scratch, so neither may skip or repeat rays."""
    import threading
    s = scenes.config_scene("C3", 0.004)
    n = 200000
    rays = scenes.bounce_rays(s, n)
    sc = lib.build_scene(s["meshes"])
    want = sc.trace_rays_compact(rays)
    k = 800
    assert_same(want[:k], orc.trace_brute(s["tris"], rays[:k]), "reference run")
    h_r, d_r = dev.put(rays)
    streams = [make_stream() if make_stream else None for _ in range(2)]
    outs = [dev.empty(16 * n, fill=0x11) for _ in range(4)]
    errors = []

    def worker(t):
        try:
            for rep in range(2):
                h_o, d_o = outs[2 * t + rep]
                if lib.rtk_trace_rays_compact_device(sc.ptr, d_r, d_o, n, streams[t]) != 0:
                    raise RuntimeError(lib.last_error())
        except Exception as ex:                                   # noqa: BLE001
            errors.append(ex)
    th = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errors:
        raise errors[0]
    dev.sync()
    for i, (h_o, _) in enumerate(outs):
        assert_same(dev.get(h_o, api.HIT16_DTYPE, n), want, f"concurrent query {i}")
    sc.free()


def case_blob_validation(lib, orc):
    """Blobs come from disk: truncated or corrupted ones are refused with a reason instead of being
    uploaded (host out-of-bounds reads) or traced (device out-of-bounds reads, endless loops)."""
    s = scenes.config_scene("C3", 0.002)
    rays = scenes.bounce_rays(s, 500)
    sc, blob = lib.build_blob(s["meshes"])
    want = sc.trace_rays_compact(rays)
    size = len(blob)
    good = blob.copy()
    sc.free()

    def attempt(mutate, what):
        buf = np.zeros(size + 256, dtype=np.uint8)
        off = (-buf.ctypes.data) % 128
        buf[off:off + size] = good
        view = buf[off:off + size]
        mutate(view)
        hits = np.zeros(len(rays), dtype=api.HIT_DTYPE)
        r = lib.rtk_trace_rays(buf.ctypes.data + off, rays.ctypes.data, hits.ctypes.data, None, len(rays))
        if r != C.c_size_t(-1).value:
            lib.rtk_cuda_detach_scene(buf.ctypes.data + off)
        return r, lib.last_error(), buf

    r, err, buf = attempt(lambda v: None, "untouched")
    assert r == int((want["prim"] != api.RTK_CUDA_MISS).sum()), err
    hdr = api.rtk_scene.from_buffer(good)          # header block: 128 bytes; payload table follows
    sub = good[128:256].view(np.uint64)             # magic2, id, 3 x u32 pairs ..., offsets at the end
    u32 = good[128:256].view(np.uint32)

    def set_size(v):
        v[24:32] = np.frombuffer(np.uint64(64).tobytes(), dtype=np.uint8)          # size_in_bytes < header block
    r, err, _ = attempt(set_size, "size below the header")
    assert r == C.c_size_t(-1).value and "truncated" in err, err

    def cut(v):
        v[24:32] = np.frombuffer(np.uint64(size // 2).tobytes(), dtype=np.uint8)    # claims half its size
    r, err, _ = attempt(cut, "truncated")
    assert r == C.c_size_t(-1).value and ("corrupt" in err or "truncated" in err), err

    def big_nodes(v):
        v[128 + 24:128 + 28] = np.frombuffer(np.uint32(0x7fffff00).tobytes(), dtype=np.uint8)   # num_nodes
    r, err, _ = attempt(big_nodes, "node count")
    assert r == C.c_size_t(-1).value and "corrupt" in err, err

    # child reference of the root pointing at the root itself: a cycle
    off_nodes = int(good[128:256].view(np.uint64)[10])

    def cycle(v):
        node0 = v[128 + off_nodes:128 + off_nodes + 256].view(np.uint32)
        k = [j for j in range(8) if node0[8 * j + 3] != 0xffffffff and not (node0[8 * j + 3] & 0x80000000)]
        assert k, "root has no internal child in this scene"
        node0[8 * k[0] + 3] = 0
    r, err, _ = attempt(cycle, "cycle")
    assert r == C.c_size_t(-1).value and "child reference" in err, err

    def fake_empty(v):
        # an "empty" reference on a slot that still carries a real box: the traversal would follow it as a leaf
        node0 = v[128 + off_nodes:128 + off_nodes + 256].view(np.uint32)
        k = [j for j in range(8) if node0[8 * j + 3] != 0xffffffff]
        node0[8 * k[-1] + 3] = 0xffffffff
    r, err, _ = attempt(fake_empty, "empty slot with a box")
    assert r == C.c_size_t(-1).value and "empty child slot" in err, err

    off_tv0 = int(good[128:256].view(np.uint64)[11])

    def bad_prim(v):
        v[128 + off_tv0 + 12:128 + off_tv0 + 16] = np.frombuffer(np.uint32(0x0ffffff0).tobytes(), dtype=np.uint8)
    r, err, _ = attempt(bad_prim, "triangle number")
    assert r == C.c_size_t(-1).value and "triangle number" in err, err


def case_overflow_report(lib, orc, dev):
    """A traversal that runs out of stack is REPORTED: host entry points fail with RTK_CUDA_ERR_OVERFLOW,
    the device path raises the scene's sticky status, and a rebuild clears it."""
    base = np.array([[(0, 0, 1), (1, 0, 1), (0, 1, 1)]], dtype=np.float32)
    tris = np.repeat(base, 16000, axis=0)              # coincident: every box overlaps every other
    rays = np.zeros(2100, dtype=api.RAY_DTYPE)
    rng = np.random.default_rng(5)
    rays["o"] = np.concatenate([rng.random((2100, 2)) * 0.6, np.zeros((2100, 1))], axis=1).astype(np.float32)
    rays["o"][300:] += 5.0                             # most rays pass the cluster by: the test stays cheap on the emulator
    rays["d"] = (0, 0, 1)
    rays["max_t"] = api.RTK_INF
    sc = lib.build_scene(soup_mesh(tris))
    try:
        want = sc.trace_rays_compact(rays)
        assert lib.rtk_cuda_scene_status(sc.ptr) == 0
        assert lib.rtk_cuda_debug_limit_stack(1) == 0           # 16 entries in shared memory + 1: far too few here
        hits = np.zeros(len(rays), dtype=api.HIT_DTYPE)
        r = lib.rtk_trace_rays(sc.ptr, rays.ctypes.data, hits.ctypes.data, None, len(rays))
        assert r == C.c_size_t(-1).value and "stack" in lib.last_error(), (r, lib.last_error())
        assert lib.rtk_cuda_scene_status(sc.ptr) == api.RTK_CUDA_ERR_OVERFLOW
        # device path: returns OK (asynchronous), the status tells
        h_r, d_r = dev.put(rays)
        h_o, d_o = dev.empty(16 * len(rays))
        assert lib.rtk_trace_rays_compact_device(sc.ptr, d_r, d_o, len(rays), dev.stream) == 0
        dev.sync()
        assert lib.rtk_cuda_scene_status(sc.ptr) == api.RTK_CUDA_ERR_OVERFLOW
        # back to automatic sizing + rebuild: the flag is cleared and the answers are right again
        assert lib.rtk_cuda_debug_limit_stack(0) == 0
        assert lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0
        assert lib.rtk_cuda_scene_status(sc.ptr) == 0
        assert_same(sc.trace_rays_compact(rays), want, "after the rebuild")
        assert lib.rtk_cuda_scene_status(sc.ptr) == 0
    finally:
        lib.rtk_cuda_debug_limit_stack(0)
        sc.free()
