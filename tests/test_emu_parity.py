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
