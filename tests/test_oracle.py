"""The oracle itself: pinned against the committed golden vectors (generated from the unmodified
reference, tests/golden/make_golden.py) and, where oracle/_ref/ is available, against the
reference builds directly."""
import numpy as np
import pytest

import parity_cases as pc
from rtk_b200 import scenes


def test_oracle_matches_golden_vectors(orc):
    kats, (rt, rr, rh) = pc.load_kats()
    assert len(kats) == 14
    for name, tris, ray, expect, literal, nreal in kats:
        got = orc.trace_brute(tris, ray)
        pc.assert_same(got, pc.expect_hit16(expect), name)
    pc.assert_same(orc.trace_brute(rt, rr), rh, "golden random soup")


def test_oracle_matches_unmodified_reference_leaf_code(orc):
    """bit-exact against rtk.c's own leaf test (flat single-leaf blobs) on random,
    degenerate and tie-heavy inputs whose triangle counts are multiples of four"""
    if not orc.have_reference():
        pytest.skip("oracle/_ref/librtk_ref.so not built (needs the reference tree)")
    rng = np.random.default_rng(1)
    tris = (rng.random((3000, 1, 3)) + 0.05 * (rng.random((3000, 3, 3)) * 2 - 1)).astype(np.float32)
    rays = np.zeros(4000, dtype=orc.RAY_DTYPE)
    rays["o"] = rng.random((4000, 3)).astype(np.float32)
    rays["d"] = rng.normal(size=(4000, 3)).astype(np.float32)
    rays["max_t"] = 3.402823e38
    a, b = orc.trace_brute(tris, rays), orc.trace_flat_reference(tris, rays)
    assert (a["prim"] != orc.MISS).sum() > 1000
    pc.assert_same(a, b, "random soup")
    # coplanar grid, rays through vertices / edges: every group of four is promoted to fp64 by
    # its own exact zeros, so own-lane and group-coupled promotion coincide
    g = scenes._quad_grid((0, 0, 1), (1, 0, 0), (0, 1, 0), 8, 8)
    grid = g[0][g[1].astype(np.int64)]
    gx, gy = np.meshgrid(np.arange(0, 33) / 32.0, np.arange(0, 33) / 32.0)
    r2 = np.zeros(gx.size, dtype=orc.RAY_DTYPE)
    r2["o"] = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], -1)
    r2["d"] = (0, 0, 1)
    r2["max_t"] = 3.402823e38
    a, b = orc.trace_brute(grid, r2), orc.trace_flat_reference(grid, r2)
    assert (a["prim"] == b["prim"]).all() and (a["prim"] != orc.MISS).all()
    assert np.allclose(a["t"], b["t"], rtol=1e-6) and np.allclose(a["u"], b["u"], atol=1e-6)


def test_patched_reference_traversal_agrees_with_oracle(orc):
    """the timed CPU baseline (rtk.c traversal with the 6-line stack fix over a blob from the
    oracle's binned-SAH packer) finds the same triangles as brute force"""
    if not orc.have_reference(patched=True):
        pytest.skip("oracle/_ref/librtk_ref_patched.so not built")
    s = scenes.config_scene("C3", 0.05)
    rays = scenes.bounce_rays(s, 6000)
    blob = orc.ReferenceBlob(s["tris"])
    assert blob.stats.num_nodes4 > 100 and blob.stats.max_depth <= 64
    got, sec = blob.trace(rays, patched=True)
    want = orc.trace_brute(s["tris"], rays)
    hit = want["prim"] != orc.MISS
    assert hit.sum() > 1000
    assert ((got["prim"] != orc.MISS) == hit).all()
    # literal rtk.c: traversal-order ties and group-coupled promotion may pick another triangle
    # at the same distance; t agrees to the north star's tolerance
    assert (got["prim"] == want["prim"]).mean() > 0.999
    assert np.allclose(got["t"][hit], want["t"][hit], rtol=1e-5)
    blob.close()


def test_scene_generators_are_deterministic():
    a, b = scenes.config_scene("C3", 0.01), scenes.config_scene("C3", 0.01)
    assert a["tris"].tobytes() == b["tris"].tobytes()
    r1, r2 = scenes.bounce_rays(a, 1000, first=500), scenes.bounce_rays(a, 2000)[500:1500]
    assert r1.tobytes() == r2.tobytes()                 # counter-based: any slice regenerates alone
    c1 = scenes.config_scene("C1")
    assert len(c1["tris"]) == 992 and c1["meshes"][0]["indices"].dtype == np.uint16
    c4 = scenes.config_scene("C4", 0.001)
    assert len(c4["meshes"]) == 2 and c4["mesh_first"][-1] == len(c4["tris"])


def test_blocked_brute_force_equals_plain_brute_force(orc):
    """orc.trace_brute switches to the blocked, pre-filtered loop for big jobs; the scalar code stays the
    arbiter of every hit there, so the two must agree bit for bit -- on random soups, on the exactly
    coplanar grid (fp64 promotion, shared edges and vertices, ties), on degenerate and NaN triangles,
    and on counts that are not multiples of the vector width or the block size."""
    rng = np.random.default_rng(11)
    for ntris, nrays in ((4099, 333), (2048, 64), (7, 40), (1, 9), (20001, 70)):
        tris = (rng.random((ntris, 1, 3)) + 0.08 * (rng.random((ntris, 3, 3)) * 2 - 1)).astype(np.float32)
        if ntris > 100:
            tris[5] = tris[4]                                  # duplicate: a tie in t
            tris[17, 2] = tris[17, 1]                          # degenerate
            tris[33, 0, 1] = np.nan
        rays = np.zeros(nrays, dtype=orc.RAY_DTYPE)
        rays["o"] = rng.random((nrays, 3)).astype(np.float32)
        rays["d"] = rng.normal(size=(nrays, 3)).astype(np.float32)
        rays["d"][::7, 1] = 0.0
        rays["max_t"] = 3.402823e38
        rays["max_t"][::5] = 0.4
        rays["min_t"][::3] = 0.1
        a = orc.trace_brute(tris, rays, blocked=False)
        b = orc.trace_brute(tris, rays, blocked=True)
        pc.assert_same(b, a, f"blocked vs plain, {ntris} triangles")
    g = scenes._quad_grid((0, 0, 1), (1, 0, 0), (0, 1, 0), 16, 16)
    grid = g[0][g[1].astype(np.int64)]
    gx, gy = np.meshgrid(np.arange(0, 65) / 64.0, np.arange(0, 65) / 64.0)
    r2 = np.zeros(gx.size, dtype=orc.RAY_DTYPE)
    r2["o"] = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], -1)
    r2["d"] = (0, 0, 1)
    r2["max_t"] = 3.402823e38
    a, b = orc.trace_brute(grid, r2, blocked=False), orc.trace_brute(grid, r2, blocked=True)
    assert (a["prim"] != orc.MISS).all()
    pc.assert_same(b, a, "blocked vs plain, coplanar grid")
    s = scenes.config_scene("C1")
    rays = scenes.config_rays("C1", s)[::97]
    pc.assert_same(orc.trace_brute(s["tris"], rays, blocked=True), orc.trace_brute(s["tris"], rays, blocked=False), "C1")
