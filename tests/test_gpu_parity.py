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


def test_trees_are_what_they_were(gpu_lib):
    """the device builds the trees the emulator builds from the same sources (and built before the build kernels were restructured)"""
    pc.case_tree_stats(gpu_lib)


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


TorchDevice = pc.TorchDevice


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


# ---- round 2 ------------------------------------------------------------------------------------

def test_direct_rows_into_page_locked_arrays(gpu_lib, orc):
    pc.case_direct_rows(gpu_lib, orc, nrays=300000, chunk_log2=14)


def test_two_streams_one_scene(gpu_lib, orc):
    import torch
    streams = []

    def make_stream():
        streams.append(torch.cuda.Stream())
        return streams[-1].cuda_stream
    pc.case_two_streams(gpu_lib, orc, TorchDevice(), make_stream)


def test_blob_validation(gpu_lib, orc):
    pc.case_blob_validation(gpu_lib, orc)


def test_overflow_is_reported(gpu_lib, orc):
    pc.case_overflow_report(gpu_lib, orc, TorchDevice())


def test_probes(gpu_lib):
    g = C.c_double(0)
    assert gpu_lib.rtk_cuda_measure_gather_bandwidth(96 << 20, 256, 3, C.byref(g)) == 0 and g.value > 500, g.value
    assert gpu_lib.rtk_cuda_measure_host_link(1, 64 << 20, 3, 2, C.byref(g)) == 0 and g.value > 5, g.value


def test_multi_device_in_one_process():
    """rtk_cuda_init_devices: ONE process drives several GPUs behind rtk_build_scene / rtk_trace_rays.
    On a box with 2+ GPUs the real ones (up to 4); on a one-GPU box the same physical device twice
    (test hook RTK_B200_TEST_DUP_DEVICES), which still runs every line of the multi-device layer:
    replicas, per-device worker threads and pipelines, range split, merge.  A process of its own, because
    a process's device list is fixed at initialisation."""
    import os
    import subprocess
    import sys
    import torch
    have = torch.cuda.device_count()
    devs = list(range(min(have, 4))) if have >= 2 else [0, 0]
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    code = (
        "import os, sys, ctypes as C\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "import parity_cases as pc\n"
        "from rtk_b200 import api\n"
        "from oracle import orc\n"
        "lib = api.load()\n"
        f"devs = (C.c_int * {len(devs)})({', '.join(str(d) for d in devs)})\n"
        f"assert lib.rtk_cuda_init_devices(devs, {len(devs)}) == 0, lib.last_error()\n"
        "g = C.c_double(0)\n"
        f"assert lib.rtk_cuda_measure_host_link({len(devs)}, 32 << 20, 3, 2, C.byref(g)) == 0 and g.value > 5, g.value\n"
        f"pc.case_multi_device(lib, orc, {len(devs)}, {'pc.TorchDevice()' if have >= 2 else 'None'}, nrays={len(devs)} * 700000 + 999, oracle_rays=2000)\n"
        "print('multi-device ok')\n"
    )
    env = dict(os.environ)
    if have < 2:
        env["RTK_B200_TEST_DUP_DEVICES"] = "1"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "multi-device ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_full_size_c4_and_one_c5_band(gpu_lib, orc):
    """BASELINE configs 4 and 5 at FULL size: the 10M-triangle terrain in two meshes.
    C4: 2^21 mixed rays (coherent / bounce / short segments) -- 2048 sampled rays against the CPU oracle
    (mesh_index / triangle_index through the host entry point included), 65536 rays against the
    exhaustive GPU kernel.  C5: one band of the 4K frame (3840 x 270 pixels x 16 spp = 16.6M rays) generated
    on the device and followed for 4 bounces; the bounce-3 rays, as the device produced them, against the
    oracle (512) and the exhaustive kernel (16384)."""
    import torch
    s = scenes.config_scene("C4")
    assert len(s["tris"]) == 10_000_000 and len(s["meshes"]) == 2
    sc = gpu_lib.build_scene(s["meshes"])
    info = sc.info()
    assert info.num_triangles == 10_000_000 and info.num_meshes == 2
    n = 1 << 21
    rays = scenes.mixed_rays(s, n, seed=0xD4)
    trav = _device_trace(gpu_lib, sc, rays)
    assert 0.2 < (trav["prim"] != api.RTK_CUDA_MISS).mean() < 0.999
    idx = np.sort(np.random.default_rng(4).choice(n, 2048, replace=False))
    want = orc.trace_brute(s["tris"], rays[idx])
    pc.assert_same(trav[idx], want, "C4 full size vs CPU oracle")
    kb = 65536
    sub = np.ascontiguousarray(rays[::n // kb][:kb])
    pc.assert_same(_device_trace(gpu_lib, sc, sub), _device_trace(gpu_lib, sc, sub, brute=True), "C4 full size vs exhaustive GPU kernel")
    # host entry point: mesh_index / triangle_index of the two 5M-triangle meshes
    hits, mask, nh = sc.trace_rays(np.ascontiguousarray(rays[idx]))
    pc.assert_same(api.hits_to_hit16(hits, mask, s["mesh_first"]), want, "C4 rows vs oracle")
    m = mask.astype(bool)
    assert set(np.unique(hits["mesh_index"][m])) <= {0, 1} and len(np.unique(hits["mesh_index"][m])) == 2
    assert (hits["triangle_index"][m] < 5_000_000).all()
    # ---- C5: one band, rays never leave the device ----
    W, H, SPP, BOUNCES, BANDS = 3840, 2160, 16, 4, 8
    npx = W * (H // BANDS)
    nr = npx * SPP
    tris = s["tris"].reshape(-1, 3)
    lo, hi = tris.min(0), tris.max(0)
    c = (lo + hi) / 2
    eye = np.array([c[0], hi[1] + 0.8 * (hi[2] - lo[2]), lo[2] - 0.6 * (hi[2] - lo[2])], dtype=np.float32)
    fwd = (c - eye) / np.linalg.norm(c - eye)
    right = np.cross([0.0, 1.0, 0.0], fwd)
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    cam = api.rtk_cuda_camera()
    cam.eye[:], cam.forward[:], cam.right[:], cam.up[:] = [float(x) for x in eye], [float(x) for x in fwd], [float(x) for x in right], [float(x) for x in up]
    cam.tan_half_fov, cam.width, cam.height = float(np.tan(np.radians(50.0) / 2)), W, H
    st = torch.cuda.current_stream().cuda_stream
    bufs = [torch.empty((nr, 32), dtype=torch.uint8, device="cuda") for _ in range(2)]
    h16 = torch.empty((nr, 16), dtype=torch.uint8, device="cuda")
    band = 3
    for sp in range(SPP):
        assert gpu_lib.rtk_cuda_generate_primary_rays(C.byref(cam), 0xD5, sp, band * npx, npx, bufs[0].data_ptr() + 32 * npx * sp, st) == 0
    cur = 0
    for b in range(BOUNCES):
        assert gpu_lib.rtk_trace_rays_compact_device(sc.ptr, bufs[cur].data_ptr(), h16.data_ptr(), nr, st) == 0, gpu_lib.last_error()
        if b + 1 < BOUNCES:
            assert gpu_lib.rtk_cuda_generate_bounce_rays(sc.ptr, bufs[cur].data_ptr(), h16.data_ptr(), bufs[cur ^ 1].data_ptr(), None, nr,
                                                         0xD5, b, band * nr, api.RTK_CUDA_BOUNCE_RELAUNCH, st) == 0
            cur ^= 1
    torch.cuda.synchronize()
    got = h16.cpu().numpy().view(api.HIT16_DTYPE).reshape(-1)
    last = bufs[cur]
    pick = np.sort(np.random.default_rng(5).choice(nr, 16384, replace=False))
    r_last = last.cpu().numpy().view(api.RAY_DTYPE).reshape(-1)[pick]
    assert (got["prim"] != api.RTK_CUDA_MISS).mean() > 0.5
    pc.assert_same(got[pick], _device_trace(gpu_lib, sc, r_last, brute=True), "C5 bounce 3 vs exhaustive GPU kernel")
    pc.assert_same(got[pick[:512]], orc.trace_brute(s["tris"], r_last[:512]), "C5 bounce 3 vs CPU oracle")
    assert gpu_lib.rtk_cuda_scene_status(sc.ptr) == 0
    sc.free()


def test_c_example_program(tmp_path):
    """examples/trace_batch.c, compiled with gcc against include/*.h and run on the device: a 64 x 64 grid of quads at
    z = 1, 2^20 rays along +z of which those over the grid must hit at t = 1"""
    import os
    import subprocess
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    lib_dir = os.path.join(root, "rtk_b200")
    exe = str(tmp_path / "trace_batch")
    subprocess.run(["gcc", "-Wall", "-std=gnu11", os.path.join(root, "examples", "trace_batch.c"), "-I", os.path.join(root, "include"),
                    "-L", lib_dir, "-lrtk_b200", "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # x, y in [-0.1, 1.1): the grid covers [0, 1]^2 -> (1/1.2)^2 of the rays (edges included), and ray 524800 is over it
    found = int(r.stdout.split()[0])
    assert abs(found / (1 << 20) - (1 / 1.2) ** 2) < 0.01, r.stdout
    assert "ray 524800: hit t=1 " in r.stdout and "rtk_trace_ray agrees: t=1" in r.stdout, r.stdout
