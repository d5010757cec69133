"""GPU tier (-m gpu): the product library librtk_b200.so on a real B200, through the C ABI,
against the CPU oracle (bit-exact) and against size-independent properties at full size."""
import ctypes as C

import numpy as np
import pytest

import parity_cases as pc
from rtk_b200 import api, scenes

pytestmark = pytest.mark.gpu


def test_known_answer_vectors(gpu_lib, orc):
    pc.case_kats(gpu_lib, orc)


def test_c1_all_rays(gpu_lib, orc):
    """config 1 in full: 992 triangles, 512x512 primary rays, every ray against the oracle"""
    assert pc.case_config(gpu_lib, orc, "C1", 1.0, 0) > 200000


@pytest.mark.parametrize("name,scale,nrays", [("C2", 0.02, 20000), ("C3", 0.02, 20000), ("C4", 0.004, 20000)])
def test_configs_reduced(gpu_lib, orc, name, scale, nrays):
    assert pc.case_config(gpu_lib, orc, name, scale, nrays) > 0


@pytest.mark.parametrize("mode", [api.RTK_CUDA_BUILD_LBVH, api.RTK_CUDA_BUILD_SAH])
def test_build_modes(gpu_lib, orc, mode):
    assert pc.case_config(gpu_lib, orc, "C3", 0.01, 8000, mode=mode) > 0
    gpu_lib.rtk_cuda_set_build_mode(api.RTK_CUDA_BUILD_LBVH)


def test_edge_scenes(gpu_lib, orc):
    pc.case_edge_scenes(gpu_lib, orc)


def test_ties_and_watertightness(gpu_lib, orc):
    pc.case_ties(gpu_lib, orc)


def test_ray_limits(gpu_lib, orc):
    pc.case_ray_limits(gpu_lib, orc)


def test_invariances(gpu_lib, orc):
    pc.case_invariances(gpu_lib, orc)


def test_mesh_formats(gpu_lib, orc):
    pc.case_mesh_formats(gpu_lib, orc)


def test_api_semantics(gpu_lib, orc):
    pc.case_api_semantics(gpu_lib, orc)


def _device_trace(lib, sc, rays_np, brute=False, stats=False):
    import torch
    rays_np = np.ascontiguousarray(rays_np)
    rays = torch.from_numpy(rays_np.view(np.uint8).reshape(-1, 32)).cuda()
    out = torch.zeros((len(rays_np), 16), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res = None
    if brute:
        r = lib.rtk_trace_rays_bruteforce_device(sc.ptr, rays.data_ptr(), out.data_ptr(), len(rays_np), st)
    elif stats:
        res = api.rtk_cuda_trace_stats()
        r = lib.rtk_trace_stats_device(sc.ptr, rays.data_ptr(), out.data_ptr(), len(rays_np), C.byref(res), st)
    else:
        r = lib.rtk_trace_rays_compact_device(sc.ptr, rays.data_ptr(), out.data_ptr(), len(rays_np), st)
    assert r == 0, lib.last_error()
    torch.cuda.synchronize()
    h = out.cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
    return (h, res) if stats else h


def test_full_size_c3_sampled_oracle_and_gpu_bruteforce(gpu_lib, orc):
    """1M-triangle terrain at full size: (a) 2048 sampled rays against the CPU oracle,
    (b) 262144 rays against the exhaustive GPU kernel (same arithmetic, no BVH), (c) device and
    host entry points agree, (d) statistics kernel returns the same hits."""
    s = scenes.config_scene("C3")
    sc = gpu_lib.build_scene(s["meshes"])
    info = sc.info()
    assert info.num_triangles == 1_000_000
    rays = scenes.bounce_rays(s, 1 << 18)
    trav = _device_trace(gpu_lib, sc, rays)
    brute = _device_trace(gpu_lib, sc, rays, brute=True)
    pc.assert_same(trav, brute, "traversal vs exhaustive GPU kernel")
    sub = rays[:2048]
    pc.assert_same(trav[:2048], orc.trace_brute(s["tris"], sub), "traversal vs CPU oracle")
    hits, mask, nh = sc.trace_rays(rays[:50000])
    pc.assert_same(api.hits_to_hit16(hits, mask, s["mesh_first"]), trav[:50000], "host vs device entry point")
    sh, st = _device_trace(gpu_lib, sc, rays[:65536], stats=True)
    pc.assert_same(sh, trav[:65536], "stats kernel")
    assert st.rays == 65536 and st.hits == int((trav[:65536]["prim"] != api.RTK_CUDA_MISS).sum())
    assert st.node_visits > 0 and st.tri_tests > 0
    sc.free()


def test_full_size_c2_coherent(gpu_lib, orc):
    s = scenes.config_scene("C2")
    sc = gpu_lib.build_scene(s["meshes"])
    rays = scenes.config_rays("C2", s)
    trav = _device_trace(gpu_lib, sc, rays)
    idx = np.random.default_rng(3).choice(len(rays), 1024, replace=False)
    pc.assert_same(trav[idx], orc.trace_brute(s["tris"], rays[idx]), "C2 sampled vs oracle")
    sub = np.ascontiguousarray(rays[:: 16])
    pc.assert_same(trav[::16], _device_trace(gpu_lib, sc, sub, brute=True), "C2 vs exhaustive GPU kernel")
    sc.free()


def test_rebuild_and_device_mesh_build(gpu_lib, orc):
    import torch
    s = scenes.config_scene("C4", 0.02)
    rays = scenes.mixed_rays(s, 30000, block=2048)
    want = orc.trace_brute(s["tris"], rays[:3000])
    keep, meshes = [], (api.rtk_cuda_mesh * len(s["meshes"]))()
    for i, m in enumerate(s["meshes"]):
        p = torch.from_numpy(m["positions"]).cuda()
        ix = torch.from_numpy(m["indices"].astype(np.int64)).to(torch.int32).cuda()
        keep += [p, ix]
        meshes[i].d_positions, meshes[i].d_indices = p.data_ptr(), ix.data_ptr()
        meshes[i].num_vertices, meshes[i].num_triangles = len(m["positions"]), len(m["indices"])
    torch.cuda.synchronize()
    ptr = gpu_lib.rtk_cuda_build_scene(meshes, len(s["meshes"]), None)
    assert ptr, gpu_lib.last_error()
    sc = api.Scene(gpu_lib, ptr)
    a = _device_trace(gpu_lib, sc, rays)
    pc.assert_same(a[:3000], want, "device-mesh build")
    assert gpu_lib.rtk_cuda_rebuild_scene(sc.ptr, None) == 0
    b = _device_trace(gpu_lib, sc, rays)
    pc.assert_same(b, a, "rebuild")
    sc.free()


def test_host_batch_pinned_vs_pageable(gpu_lib, orc):
    """rtk_trace_rays with pinned buffers (asynchronous, overlapped chunk pipeline) and with
    pageable buffers gives the same rows; rows of misses are left untouched (rtk.c:571-576)"""
    import torch
    s = scenes.config_scene("C3", 0.05)
    rays = scenes.bounce_rays(s, 300000)
    sc = gpu_lib.build_scene(s["meshes"])
    hits_p, mask_p, n_p = sc.trace_rays(rays)                       # pageable numpy arrays
    t_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32)).pin_memory()
    t_hits = torch.full((len(rays), 68), 0xAB, dtype=torch.uint8).pin_memory()
    t_mask = torch.full((len(rays),), 0xCD, dtype=torch.uint8).pin_memory()
    n_d = gpu_lib.rtk_trace_rays(sc.ptr, t_rays.data_ptr(), t_hits.data_ptr(), t_mask.data_ptr(), len(rays))
    assert n_d == n_p and 0 < n_p < len(rays)
    mask_d = t_mask.numpy()
    hits_d = t_hits.numpy().reshape(-1).view(api.HIT_DTYPE)
    assert np.array_equal(mask_d, mask_p)
    m = mask_p.astype(bool)
    assert hits_d[m].tobytes() == hits_p[m].tobytes()
    assert (t_hits.numpy()[~m] == 0xAB).all() and (hits_p.view(np.uint8).reshape(-1, 68)[~m] == 0).all()
    pc.assert_same(api.hits_to_hit16(hits_d, mask_d, s["mesh_first"])[:1500], orc.trace_brute(s["tris"], rays[:1500]), "pinned path vs oracle")
    sc.free()


def test_occlusion_query(gpu_lib, orc):
    import torch

    def alloc(rays, n):
        r = torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(-1, 32)).cuda()
        out = torch.full((n,), 7, dtype=torch.uint8, device="cuda")

        def fetch(keep=(r, out)):
            torch.cuda.synchronize()
            return keep[1].cpu().numpy()
        return r.data_ptr(), out.data_ptr(), fetch
    pc.case_occlusion(gpu_lib, orc, alloc)


class TorchDevice:
    """device buffers for the shared cases: torch owns the memory, the library sees raw pointers"""
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
        torch.cuda.synchronize()
        return handle[:count * np.dtype(dtype).itemsize].cpu().numpy().view(dtype)


def test_wavefront_generators(gpu_lib, orc):
    pc.case_wavefront(gpu_lib, orc, TorchDevice())


def test_refit_and_rebuild_of_a_deformed_mesh(gpu_lib, orc):
    pc.case_refit(gpu_lib, orc, TorchDevice())


def test_deep_stack_spills(gpu_lib, orc):
    pc.case_deep_stack(gpu_lib, orc, TorchDevice())


def test_host_batch_many_chunks(gpu_lib, orc):
    pc.case_host_batch_chunks(gpu_lib, orc, nrays=300000, chunk_log2=14)


def test_triangle_filter(gpu_lib, orc):
    pc.case_triangle_filter(gpu_lib, orc, TorchDevice())


def test_concurrent_host_threads(gpu_lib, orc):
    pc.case_threads(gpu_lib, orc, nthreads=6)


def test_baked_instancing(gpu_lib, orc):
    pc.case_instancing(gpu_lib, orc, TorchDevice())


def test_pathological_scenes(gpu_lib, orc):
    pc.case_pathological(gpu_lib, orc)
