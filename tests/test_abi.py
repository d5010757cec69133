"""ABI of the drop-in boundary: struct sizes / offsets of include/rtk.h as C and as C++
(measured on the reference header), and the exported symbols of the
shared library.  No GPU, no compute calls."""
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

PROBE = r"""
#include "rtk.h"
#include "rtk_cuda.h"
#include <stdio.h>
#include <stddef.h>
#define S(t) printf(#t " %zu\n", sizeof(t))
#define O(t, f) printf(#t "." #f " %zu\n", offsetof(t, f))
int main(void) {
  S(rtk_vec3); S(rtk_vertex); S(rtk_ray); S(rtk_hit); S(rtk_buffer); S(rtk_mesh); S(rtk_scene);
  S(rtk_scene_desc); S(rtk_task); S(rtk_cuda_hit16);
  O(rtk_ray, direction); O(rtk_ray, min_t); O(rtk_ray, max_t);
  O(rtk_hit, t); O(rtk_hit, u); O(rtk_hit, v); O(rtk_hit, vertex); O(rtk_hit, mesh_index); O(rtk_hit, triangle_index);
  O(rtk_mesh, num_triangles); O(rtk_mesh, position); O(rtk_mesh, index); O(rtk_mesh, position_cb);
  O(rtk_scene, endian); O(rtk_scene, sizeof_real); O(rtk_scene, version); O(rtk_scene, size_in_bytes);
  O(rtk_scene, node_offset); O(rtk_scene, leaf_offset); O(rtk_scene, vertex_offset);
  O(rtk_task, cost); O(rtk_task, index); O(rtk_task, arg);
  printf("enum %d %d %d %d %d %d\n", RTK_TYPE_DEFAULT, RTK_TYPE_F32, RTK_TYPE_F64, RTK_TYPE_REAL, RTK_TYPE_U16, RTK_TYPE_U32);
  printf("inf %a\n", (double)RTK_INF);
  return 0;
}
"""

EXPECT = {
    "rtk_vec3": 12, "rtk_vertex": 16, "rtk_ray": 32, "rtk_hit": 68, "rtk_buffer": 24, "rtk_mesh": 96,
    "rtk_scene": 56, "rtk_scene_desc": 32, "rtk_task": 40, "rtk_cuda_hit16": 16,
    "rtk_ray.direction": 12, "rtk_ray.min_t": 24, "rtk_ray.max_t": 28,
    "rtk_hit.t": 0, "rtk_hit.u": 4, "rtk_hit.v": 8, "rtk_hit.vertex": 12, "rtk_hit.mesh_index": 60,
    "rtk_hit.triangle_index": 64,
    "rtk_mesh.num_triangles": 8, "rtk_mesh.position": 16, "rtk_mesh.index": 40, "rtk_mesh.position_cb": 64,
    "rtk_scene.endian": 8, "rtk_scene.sizeof_real": 10, "rtk_scene.version": 12, "rtk_scene.size_in_bytes": 24,
    "rtk_scene.node_offset": 32, "rtk_scene.leaf_offset": 40, "rtk_scene.vertex_offset": 48,
    "rtk_task.cost": 16, "rtk_task.index": 24, "rtk_task.arg": 32,
}


@pytest.mark.parametrize("compiler,lang", [("gcc", "c"), ("g++", "c++")])
def test_struct_layout(compiler, lang):
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "probe." + ("c" if lang == "c" else "cpp"))
        with open(src, "w") as f:
            f.write(PROBE)
        exe = os.path.join(d, "probe")
        subprocess.run([compiler, "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = {}
    for line in out.splitlines():
        k, v = line.split(" ", 1)
        got[k] = v
    for k, v in EXPECT.items():
        assert int(got[k]) == v, (k, got[k], v)
    assert got["enum"] == "0 1 2 3 4 5"
    assert got["inf"] == "0x1.fffffap+127"          # reference rtk.h:11, SURVEY 8(a)


def test_reference_header_agrees():
    """where the reference tree is present (authoring container), its own header gives the same
    numbers: the re-authored include/rtk.h is ABI-identical"""
    ref = "/root/reference/rtk.h"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this box")
    probe = PROBE.replace('#include "rtk_cuda.h"\n', "").replace(" S(rtk_cuda_hit16);", "")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "probe.c")
        with open(src, "w") as f:
            f.write(probe)
        exe = os.path.join(d, "probe")
        subprocess.run(["gcc", "-I", "/root/reference", src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(line.split(" ", 1) for line in out.splitlines())
    for k, v in EXPECT.items():
        if k != "rtk_cuda_hit16":
            assert int(got[k]) == v, (k, got[k], v)


def test_library_exports_every_declared_symbol():
    """librtk_b200.so loads and exports every function include/*.h declare (ctypes resolves each
    one in api.Library); no compute call is made, so this runs without a GPU."""
    sys.path.insert(0, ROOT)
    import re
    from rtk_b200 import api, build
    path = build.build()
    lib = api.Library(path)
    declared = set()
    for h in ("rtk.h", "rtk_cuda.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(rtk_[a-z0-9_]+)\s*\(", text))
    declared -= {"rtk_position_callback_fn", "rtk_index_callback_fn", "rtk_log_fn", "rtk_task_fn", "rtk_filter_fn"}
    assert declared == set(api.SYMBOLS), declared ^ set(api.SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in nm.splitlines() if " T " in line}
    assert declared <= exported, declared - exported
    assert lib.path == path


def test_no_device_fails_loudly():
    """without a CUDA device the product refuses to work instead of falling back to a CPU path"""
    sys.path.insert(0, ROOT)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from rtk_b200 import api\n"
        "lib = api.load()\n"
        "r = lib.rtk_cuda_init(0)\n"
        "assert r == -1, r\n"
        "assert 'no CPU fallback' in lib.last_error(), lib.last_error()\n"
        "try:\n"
        "    lib.build_scene([])\n"
        "    raise SystemExit('built a scene without a GPU')\n"
        "except api.RtkError as e:\n"
        "    print('ok', e)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_c_example_links_and_fails_loudly_without_a_device(tmp_path):
    """examples/trace_batch.c is the reference user's program after the switch: plain C against
    include/*.h, linked with librtk_b200.so.  Without a CUDA device it must say so and exit non-zero (no CPU path)."""
    import shutil
    lib = os.path.join(ROOT, "rtk_b200", "librtk_b200.so")
    if not os.path.exists(lib):
        pytest.skip("librtk_b200.so not built")
    exe = str(tmp_path / "trace_batch")
    subprocess.run(["gcc", "-Wall", "-Werror", "-std=gnu11", os.path.join(ROOT, "examples", "trace_batch.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", os.path.dirname(lib), "-lrtk_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe], check=True)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = shutil.which("nvidia-smi") is not None
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if has_gpu:
        assert r.returncode == 0 and "rays hit" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode != 0 and "no CPU fallback" in r.stderr, r.stdout + r.stderr
