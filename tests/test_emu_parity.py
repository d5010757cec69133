"""CPU tier: the product's own CUDA sources and host layer, compiled against the SIMT emulator
(tests/emu) and driven through the C ABI, against the CPU oracle.  Small sizes only -- the GPU
tier (test_gpu_parity.py, -m gpu) runs the same cases on the real library."""
import pytest

import parity_cases as pc


def test_known_answer_vectors(emu_lib, orc):
    pc.case_kats(emu_lib, orc)


@pytest.mark.parametrize("name,scale,nrays", [("C1", 1.0, 1500), ("C2", 0.004, 1200), ("C3", 0.008, 1500), ("C4", 0.001, 1500)])
def test_configs_small(emu_lib, orc, name, scale, nrays):
    assert pc.case_config(emu_lib, orc, name, scale, nrays) > 0


def test_trees_are_what_they_were(emu_lib):
    pc.case_tree_stats(emu_lib)


def test_edge_scenes(emu_lib, orc):
    pc.case_edge_scenes(emu_lib, orc)


def test_ties_and_watertightness(emu_lib, orc):
    pc.case_ties(emu_lib, orc)


def test_ray_limits(emu_lib, orc):
    pc.case_ray_limits(emu_lib, orc)


def test_invariances(emu_lib, orc):
    pc.case_invariances(emu_lib, orc)


def test_mesh_formats(emu_lib, orc):
    pc.case_mesh_formats(emu_lib, orc)


def test_api_semantics(emu_lib, orc):
    pc.case_api_semantics(emu_lib, orc)


def test_occlusion_query(emu_lib, orc):
    import numpy as np

    def alloc(rays, n):                                 # emulated device memory is host memory
        r = np.ascontiguousarray(rays)
        out = np.full(n, 7, dtype=np.uint8)
        return r.ctypes.data, out.ctypes.data, lambda keep=(r, out): keep[1]
    pc.case_occlusion(emu_lib, orc, alloc)


def test_wavefront_generators(emu_lib, orc):
    pc.case_wavefront(emu_lib, orc, pc.HostDevice())


def test_refit_and_rebuild_of_a_deformed_mesh(emu_lib, orc):
    pc.case_refit(emu_lib, orc, pc.HostDevice())


def test_concurrent_host_threads(emu_lib, orc):
    pc.case_threads(emu_lib, orc)


def test_baked_instancing(emu_lib, orc):
    pc.case_instancing(emu_lib, orc, pc.HostDevice())


def test_pathological_scenes(emu_lib, orc):
    pc.case_pathological(emu_lib, orc)


def test_triangle_filter(emu_lib, orc):
    pc.case_triangle_filter(emu_lib, orc, pc.HostDevice())


def test_read_bandwidth_probe(emu_lib):
    import ctypes as C
    g = C.c_double(0)
    assert emu_lib.rtk_cuda_measure_read_bandwidth(1 << 16, 2, C.byref(g)) == 0
    assert g.value > 0
    assert emu_lib.rtk_cuda_measure_read_bandwidth(16, 2, C.byref(g)) != 0          # too small: refused


def test_deep_stack_spills(emu_lib, orc):
    pc.case_deep_stack(emu_lib, orc, pc.HostDevice())


def test_host_batch_many_chunks(emu_lib, orc):
    pc.case_host_batch_chunks(emu_lib, orc, nrays=30000, chunk_log2=12)


def test_randomised_adversarial_scenes(emu_lib, orc):
    """a fixed slice of tools/fuzz_emu.py (slivers, degenerate and duplicated triangles, lattice-aligned
    geometry, rays through vertices / along axes / starting on surfaces, tight t windows; both
    builders, with and without a triangle filter, closest hit and occlusion)"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import fuzz_emu
    hits = 0
    for seed in range(1000, 1040):
        hits += fuzz_emu.one(emu_lib, seed)[2]
    assert hits > 500
    emu_lib.rtk_cuda_set_build_mode(1)                  # back to the default (binned SAH)


# ---- round 2 ------------------------------------------------------------------------------------

def test_direct_rows_into_page_locked_arrays(emu_lib, orc):
    pc.case_direct_rows(emu_lib, orc, nrays=20000, chunk_log2=12)


def test_two_streams_one_scene(emu_lib, orc):
    pc.case_two_streams(emu_lib, orc, pc.HostDevice())


def test_blob_validation(emu_lib, orc):
    pc.case_blob_validation(emu_lib, orc)


def test_overflow_is_reported(emu_lib, orc):
    pc.case_overflow_report(emu_lib, orc, pc.HostDevice())


def test_probes(emu_lib):
    import ctypes as C
    g = C.c_double(0)
    assert emu_lib.rtk_cuda_measure_gather_bandwidth(1 << 18, 256, 1, C.byref(g)) == 0 and g.value > 0
    assert emu_lib.rtk_cuda_measure_gather_bandwidth(1 << 18, 128, 1, C.byref(g)) == 0 and g.value > 0
    assert emu_lib.rtk_cuda_measure_gather_bandwidth(1 << 18, 100, 1, C.byref(g)) != 0
    assert emu_lib.rtk_cuda_measure_host_link(1, 1 << 16, 3, 2, C.byref(g)) == 0 and g.value > 0
    assert emu_lib.rtk_cuda_measure_host_link(2, 1 << 16, 3, 2, C.byref(g)) != 0          # one device in use


def test_multi_device_in_one_process():
    """rtk_cuda_init_devices over 3 emulated devices (SIMT_DEVICES=3): a process of its own, because the
    device list of a process is fixed at initialisation."""
    import os
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    code = (
        "import os, sys, ctypes as C\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r}); sys.path.insert(0, {os.path.join(root, 'tests', 'emu')!r})\n"
        "import build_emu, parity_cases as pc\n"
        "from rtk_b200 import api\n"
        "from oracle import orc\n"
        "lib = api.Library(build_emu.build())\n"
        "devs = (C.c_int * 3)(0, 1, 2)\n"
        "assert lib.rtk_cuda_init_devices(devs, 3) == 0, lib.last_error()\n"
        "assert lib.rtk_cuda_init_devices(devs, 3) == 0\n"
        "assert lib.rtk_cuda_init(1) != 0 and 'already bound' in lib.last_error()\n"
        "g = C.c_double(0)\n"
        "assert lib.rtk_cuda_measure_host_link(3, 1 << 16, 3, 2, C.byref(g)) == 0 and g.value > 0\n"
        "pc.case_multi_device(lib, orc, 3, pc.HostDevice(), nrays=3 * (1 << 14) + 1000, oracle_rays=500)\n"
        "lib.rtk_cuda_shutdown()\n"
        "two = (C.c_int * 2)(2, 0)\n"
        "assert lib.rtk_cuda_init_devices(two, 2) == 0, lib.last_error()\n"
        "pc.case_multi_device(lib, orc, 2, None, nrays=2 * (1 << 14), oracle_rays=200)\n"
        "print('multi-device ok')\n"
    )
    env = dict(os.environ, SIMT_DEVICES="3", RTK_B200_HOST_MIN_SHARE_LOG2="14", RTK_B200_HOST_CHUNK_LOG2="12")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "multi-device ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("bits", ["8", "32"])
def test_sah_tree_does_not_depend_on_the_input_order(bits):
    """RTK_B200_SAH_SORT_BITS only changes the ORDER the SAH builder meets the triangles in (8 / 24 / 32 Morton bits); bins,
    counts and boxes are functions of the triangle sets, so the pinned trees must come out whatever the order.  A
    process of its own: the knob is read at initialisation."""
    import os
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    code = (
        "import sys\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r}); sys.path.insert(0, {os.path.join(root, 'tests', 'emu')!r})\n"
        "import build_emu, parity_cases as pc\n"
        "from rtk_b200 import api\n"
        "lib = api.Library(build_emu.build())\n"
        "assert lib.rtk_cuda_init(0) == 0, lib.last_error()\n"
        "pc.case_tree_stats(lib)\n"
        "print('trees ok')\n"
    )
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RTK_B200_SAH_SORT_BITS=bits), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "trees ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
