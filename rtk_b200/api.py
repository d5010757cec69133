"""ctypes mirror of include/rtk.h and include/rtk_cuda.h.

Same names, same argument meaning and the same error behaviour as the C ABI (which in turn
mirrors the reference's rtk.h): builders return None (NULL) on failure, rtk_trace_ray returns
False and leaves the hit untouched on a miss.  The product library is librtk_b200.so, compiled
for sm_100a by rtk_b200/build.py; there is no CPU path here -- `load()` raises when the library
is missing and every call fails loudly when no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librtk_b200.so")

RTK_INF = np.float32(3.402823e+38)
RTK_CUDA_MISS = 0xFFFFFFFF
(RTK_TYPE_DEFAULT, RTK_TYPE_F32, RTK_TYPE_F64, RTK_TYPE_REAL, RTK_TYPE_U16, RTK_TYPE_U32) = range(6)
RTK_CUDA_BUILD_LBVH, RTK_CUDA_BUILD_SAH = 0, 1
(RTK_CUDA_OK, RTK_CUDA_ERR_NO_DEVICE, RTK_CUDA_ERR_CUDA, RTK_CUDA_ERR_ARGUMENT, RTK_CUDA_ERR_SCENE,
 RTK_CUDA_ERR_MEMORY, RTK_CUDA_ERR_OVERFLOW) = (0, -1, -2, -3, -4, -5, -6)


class rtk_vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class rtk_vertex(C.Structure):
    _fields_ = [("position", rtk_vec3), ("index", C.c_uint32)]


class rtk_ray(C.Structure):
    _fields_ = [("origin", rtk_vec3), ("direction", rtk_vec3), ("min_t", C.c_float), ("max_t", C.c_float)]


class rtk_hit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("vertex", rtk_vertex * 3),
                ("mesh_index", C.c_uint32), ("triangle_index", C.c_uint32)]


class rtk_buffer(C.Structure):
    _fields_ = [("data", C.c_void_p), ("stride", C.c_size_t), ("type", C.c_int)]


class rtk_mesh(C.Structure):
    pass


rtk_position_callback_fn = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(rtk_mesh), C.POINTER(rtk_vec3),
                                       C.POINTER(C.c_uint32), C.c_size_t)
rtk_index_callback_fn = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(rtk_mesh), C.POINTER(C.c_uint32),
                                    C.c_size_t, C.c_size_t)
rtk_mesh._fields_ = [("user", C.c_void_p), ("num_triangles", C.c_size_t),
                     ("position", rtk_buffer), ("index", rtk_buffer),
                     ("position_cb", rtk_position_callback_fn), ("position_cb_user", C.c_void_p),
                     ("index_cb", rtk_index_callback_fn), ("index_cb_user", C.c_void_p)]


class rtk_scene(C.Structure):
    _fields_ = [("magic", C.c_char * 8), ("endian", C.c_uint16), ("sizeof_real", C.c_uint8), ("pad_0", C.c_uint8),
                ("version", C.c_uint32), ("pad_1", C.c_uint32), ("size_in_bytes", C.c_uint64),
                ("node_offset", C.c_uint64), ("leaf_offset", C.c_uint64), ("vertex_offset", C.c_uint64)]


rtk_log_fn = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_char_p)


class rtk_scene_desc(C.Structure):
    _fields_ = [("meshes", C.POINTER(rtk_mesh)), ("num_meshes", C.c_size_t),
                ("log_fn", rtk_log_fn), ("log_user", C.c_void_p)]


class rtk_task(C.Structure):
    _fields_ = [("build", C.c_void_p), ("fn", C.c_void_p), ("cost", C.c_double),
                ("index", C.c_size_t), ("arg", C.c_size_t)]


rtk_filter_fn = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.POINTER(rtk_ray), C.POINTER(rtk_hit))


class rtk_cuda_mesh(C.Structure):
    _fields_ = [("d_positions", C.c_void_p), ("d_indices", C.c_void_p),
                ("num_vertices", C.c_size_t), ("num_triangles", C.c_size_t)]


class rtk_cuda_instance(C.Structure):
    _fields_ = [("mesh", C.c_uint32), ("transform", C.c_float * 12)]


class rtk_cuda_camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("forward", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3),
                ("tan_half_fov", C.c_float), ("width", C.c_uint32), ("height", C.c_uint32)]


class rtk_cuda_scene_info(C.Structure):
    _fields_ = [("num_triangles", C.c_uint64), ("num_meshes", C.c_uint64), ("num_wide_nodes", C.c_uint64),
                ("num_leaves", C.c_uint64), ("wide_depth", C.c_uint32), ("build_mode", C.c_uint32),
                ("device_bytes", C.c_uint64), ("build_device_ms", C.c_double), ("build_total_ms", C.c_double),
                ("sah_cost", C.c_double), ("bounds_min", C.c_float * 3), ("bounds_max", C.c_float * 3)]


class rtk_cuda_trace_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("hits", C.c_uint64), ("node_visits", C.c_uint64),
                ("leaf_visits", C.c_uint64), ("tri_tests", C.c_uint64), ("stack_max", C.c_uint64)]


RTK_CUDA_BOUNCE_RELAUNCH = 1
RTK_CUDA_UPDATE_REFIT, RTK_CUDA_UPDATE_REBUILD = 0, 1

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("min_t", "<f4"), ("max_t", "<f4")])
VERTEX_DTYPE = np.dtype([("position", "<f4", 3), ("index", "<u4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("vertex", VERTEX_DTYPE, 3),
                      ("mesh_index", "<u4"), ("triangle_index", "<u4")])
HIT16_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 68 and HIT16_DTYPE.itemsize == 16

# every symbol include/rtk.h and include/rtk_cuda.h declare: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    # rtk.h
    "rtk_start_build": (_P, [C.POINTER(rtk_scene_desc), C.POINTER(rtk_task)]),
    "rtk_run_task": (C.c_size_t, [C.POINTER(rtk_task), C.POINTER(rtk_task), C.c_size_t]),
    "rtk_get_build_size": (C.c_size_t, [_P]),
    "rtk_finish_build_to": (_P, [_P, _P, C.c_size_t]),
    "rtk_finish_build": (_P, [_P]),
    "rtk_build_scene": (_P, [C.POINTER(rtk_scene_desc)]),
    "rtk_free_scene": (None, [_P]),
    "rtk_trace_ray": (C.c_bool, [_P, C.POINTER(rtk_ray), C.POINTER(rtk_hit)]),
    "rtk_trace_ray_filter": (C.c_bool, [_P, C.POINTER(rtk_ray), C.POINTER(rtk_hit), rtk_filter_fn, _P]),
    # rtk_cuda.h
    "rtk_cuda_init": (C.c_int, [C.c_int]),
    "rtk_cuda_init_devices": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "rtk_cuda_device_count": (C.c_int, []),
    "rtk_cuda_host_alloc": (_P, [C.c_size_t]),
    "rtk_cuda_host_free": (None, [_P]),
    "rtk_cuda_host_alloc_batch": (_P, [C.c_size_t, C.c_size_t]),
    "rtk_cuda_host_register": (C.c_int, [_P, C.c_size_t]),
    "rtk_cuda_host_unregister": (C.c_int, [_P]),
    "rtk_cuda_scene_status": (C.c_int, [_P]),
    "rtk_cuda_debug_limit_stack": (C.c_int, [C.c_int]),
    "rtk_cuda_measure_gather_bandwidth": (C.c_int, [C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
    "rtk_cuda_measure_host_link": (C.c_int, [C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "rtk_cuda_shutdown": (None, []),
    "rtk_cuda_last_error": (C.c_char_p, []),
    "rtk_cuda_set_build_mode": (C.c_int, [C.c_int]),
    "rtk_cuda_set_cull_mode": (C.c_int, [C.c_int]),
    "rtk_cuda_reserve_sms": (C.c_int, [C.c_int]),
    "rtk_cuda_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rtk_cuda_measure_read_bandwidth": (C.c_int, [C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
    "rtk_trace_rays": (C.c_size_t, [_P, _P, _P, _P, C.c_size_t]),
    "rtk_trace_rays_compact": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rtk_trace_rays_device": (C.c_int, [_P, _P, _P, _P, C.c_size_t, _P]),
    "rtk_trace_rays_compact_device": (C.c_int, [_P, _P, _P, C.c_size_t, _P]),
    "rtk_resolve_hits_device": (C.c_int, [_P, _P, _P, _P, C.c_size_t, _P]),
    "rtk_occluded_rays_device": (C.c_int, [_P, _P, _P, C.c_size_t, _P]),
    "rtk_cuda_peer_window_create": (C.c_int, [C.c_size_t, C.POINTER(_P), _P]),
    "rtk_cuda_peer_window_open": (C.c_int, [_P, C.POINTER(_P)]),
    "rtk_cuda_peer_window_close": (C.c_int, [_P]),
    "rtk_cuda_peer_window_destroy": (C.c_int, [_P]),
    "rtk_cuda_peer_push": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "rtk_cuda_set_triangle_filter": (C.c_int, [_P, _P, C.c_size_t]),
    "rtk_cuda_set_triangle_filter_device": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "rtk_trace_rays_bruteforce_device": (C.c_int, [_P, _P, _P, C.c_size_t, _P]),
    "rtk_trace_stats_device": (C.c_int, [_P, _P, _P, C.c_size_t, C.POINTER(rtk_cuda_trace_stats), _P]),
    "rtk_cuda_generate_primary_rays": (C.c_int, [C.POINTER(rtk_cuda_camera), C.c_uint64, C.c_uint32, C.c_size_t, C.c_size_t, _P, _P]),
    "rtk_cuda_generate_bounce_rays": (C.c_int, [_P, _P, _P, _P, _P, C.c_size_t, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, _P]),
    "rtk_cuda_build_scene": (_P, [C.POINTER(rtk_cuda_mesh), C.c_size_t, _P]),
    "rtk_cuda_rebuild_scene": (C.c_int, [_P, _P]),
    "rtk_cuda_update_scene": (C.c_int, [_P, C.POINTER(rtk_cuda_mesh), C.c_size_t, C.c_int, _P]),
    "rtk_cuda_build_instanced_scene": (_P, [C.POINTER(rtk_cuda_mesh), C.c_size_t, C.POINTER(rtk_cuda_instance), C.c_size_t, _P]),
    "rtk_cuda_update_instanced_scene": (C.c_int, [_P, C.POINTER(rtk_cuda_mesh), C.c_size_t, C.POINTER(rtk_cuda_instance), C.c_size_t, C.c_int, _P]),
    "rtk_cuda_get_scene_info": (C.c_int, [_P, C.POINTER(rtk_cuda_scene_info)]),
    "rtk_cuda_attach_scene": (C.c_int, [_P]),
    "rtk_cuda_detach_scene": (C.c_int, [_P]),
}


class RtkError(RuntimeError):
    pass


class Library:
    """One loaded copy of the C ABI.  Attribute access gives the raw C functions by their rtk.h
    names; the methods below are thin conveniences over numpy arrays."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise RtkError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  rtk_b200 has no CPU fallback.")
        self.path = path
        self.c = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.c, name)          # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def last_error(self):
        e = self.rtk_cuda_last_error()
        return e.decode() if e else ""

    # ---- scene creation from numpy meshes (rtk_build_scene / rtk_start_build ...) ------------
    def make_desc(self, meshes, log_fn=None):
        """meshes: list of dicts {positions (nv,3) f32|f64, indices (nt,3) u16|u32 or None}.
        Returns (desc, keepalive)."""
        arr = (rtk_mesh * max(1, len(meshes)))()
        keep = []
        for i, m in enumerate(meshes):
            pos = np.ascontiguousarray(m["positions"])
            if pos.dtype == np.float64:
                ptype = RTK_TYPE_F64
            else:
                pos = np.ascontiguousarray(pos, dtype=np.float32)
                ptype = m.get("position_type", RTK_TYPE_F32)
            keep.append(pos)
            arr[i].position.data = pos.ctypes.data
            arr[i].position.stride = m.get("position_stride", 0)
            arr[i].position.type = ptype
            idx = m.get("indices")
            if idx is not None:
                if idx.dtype == np.uint16:
                    itype = RTK_TYPE_U16
                    idx = np.ascontiguousarray(idx)
                else:
                    idx = np.ascontiguousarray(idx, dtype=np.uint32)
                    itype = m.get("index_type", RTK_TYPE_U32)
                keep.append(idx)
                arr[i].index.data = idx.ctypes.data
                arr[i].index.stride = m.get("index_stride", 0)
                arr[i].index.type = itype
                arr[i].num_triangles = len(idx)
            else:
                arr[i].num_triangles = m.get("num_triangles", len(pos) // 3)
        desc = rtk_scene_desc()
        desc.meshes = arr
        desc.num_meshes = len(meshes)
        if log_fn is not None:
            desc.log_fn = log_fn
            keep.append(log_fn)
        keep.append(arr)
        return desc, keep

    def build_scene(self, meshes, mode=None):
        """rtk_build_scene over numpy meshes -> Scene (device-resident handle)."""
        if mode is not None:
            self.rtk_cuda_set_build_mode(mode)
        desc, keep = self.make_desc(meshes)
        ptr = self.rtk_build_scene(C.byref(desc))
        del keep
        if not ptr:
            raise RtkError("rtk_build_scene failed: " + self.last_error())
        return Scene(self, ptr)

    def build_blob(self, meshes, mode=None):
        """rtk_start_build / rtk_run_task pump / rtk_get_build_size / rtk_finish_build_to into a
        numpy buffer -> (Scene, buffer).  The buffer is the relocatable blob."""
        if mode is not None:
            self.rtk_cuda_set_build_mode(mode)
        desc, keep = self.make_desc(meshes)
        first = rtk_task()
        b = self.rtk_start_build(C.byref(desc), C.byref(first))
        if not b:
            raise RtkError("rtk_start_build failed: " + self.last_error())
        queue = (rtk_task * 16)()
        pending = [first]
        while pending:
            t = pending.pop()
            n = self.rtk_run_task(C.byref(t), queue, 16)
            for i in range(n):
                c = rtk_task()
                C.memmove(C.byref(c), C.byref(queue[i]), C.sizeof(rtk_task))
                pending.append(c)
        size = self.rtk_get_build_size(b)
        if not size:
            raise RtkError("build failed: " + self.last_error())
        buf = np.zeros(size + 128, dtype=np.uint8)
        off = (-buf.ctypes.data) % 128
        ptr = self.rtk_finish_build_to(b, buf.ctypes.data + off, size)
        del keep
        if not ptr:
            raise RtkError("rtk_finish_build_to failed: " + self.last_error())
        return Scene(self, ptr, owner=buf), buf[off:off + size]


class Scene:
    """A built scene (rtk_scene *)."""

    def __init__(self, lib, ptr, owner=None):
        self.lib, self.ptr, self.owner = lib, ptr, owner

    def header(self):
        return rtk_scene.from_address(self.ptr)

    def info(self):
        i = rtk_cuda_scene_info()
        r = self.lib.rtk_cuda_get_scene_info(self.ptr, C.byref(i))
        if r:
            raise RtkError("rtk_cuda_get_scene_info: " + self.lib.last_error())
        return i

    def trace_ray(self, ray):
        """rtk_trace_ray: one numpy RAY_DTYPE record -> HIT_DTYPE record or None."""
        r = np.ascontiguousarray(ray, dtype=RAY_DTYPE).reshape(1)
        h = np.zeros(1, dtype=HIT_DTYPE)
        ok = self.lib.rtk_trace_ray(self.ptr, C.cast(r.ctypes.data, C.POINTER(rtk_ray)),
                                    C.cast(h.ctypes.data, C.POINTER(rtk_hit)))
        return h[0] if ok else None

    def trace_rays(self, rays, hits=None, mask=None):
        """rtk_trace_rays with host arrays -> (hits HIT_DTYPE, mask uint8, number of hits)."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = len(rays)
        if hits is None:
            hits = np.zeros(n, dtype=HIT_DTYPE)
        if mask is None:
            mask = np.zeros(n, dtype=np.uint8)
        r = self.lib.rtk_trace_rays(self.ptr, rays.ctypes.data, hits.ctypes.data, mask.ctypes.data, n)
        if r == C.c_size_t(-1).value:
            raise RtkError("rtk_trace_rays failed: " + self.lib.last_error())
        return hits, mask, int(r)

    def trace_rays_compact(self, rays, out=None):
        """rtk_trace_rays_compact with host arrays -> HIT16_DTYPE record per ray."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        if out is None:
            out = np.zeros(len(rays), dtype=HIT16_DTYPE)
        r = self.lib.rtk_trace_rays_compact(self.ptr, rays.ctypes.data, out.ctypes.data, len(rays))
        if r:
            raise RtkError("rtk_trace_rays_compact failed: " + self.lib.last_error())
        return out

    def set_triangle_filter(self, keep):
        """rtk_cuda_set_triangle_filter: `keep` is a boolean array over the scene's global triangle
        numbers (True = the triangle takes part in queries) or None to remove the filter."""
        if keep is None:
            r = self.lib.rtk_cuda_set_triangle_filter(self.ptr, None, 0)
        else:
            bits = pack_triangle_filter(keep)
            r = self.lib.rtk_cuda_set_triangle_filter(self.ptr, bits.ctypes.data, len(bits))
        if r:
            raise RtkError("rtk_cuda_set_triangle_filter: " + self.lib.last_error())

    def free(self):
        if self.ptr:
            self.lib.rtk_free_scene(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


_lib = None


def load():
    """The product library (librtk_b200.so, sm_100a).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        _lib = Library(LIB_PATH)
    return _lib


def pack_triangle_filter(keep):
    """Boolean array over global triangle numbers -> the uint32 bitset of rtk_cuda_set_triangle_filter
    (bit i & 31 of word i >> 5)."""
    keep = np.asarray(keep, dtype=bool)
    words = (len(keep) + 31) // 32
    padded = np.zeros(words * 32, dtype=np.uint8)
    padded[:len(keep)] = keep
    return np.ascontiguousarray(np.packbits(padded, bitorder="little").view("<u4"))


def hits_to_hit16(hits, mask, mesh_first):
    """Expanded hits -> compact (t,u,v,global triangle number) records for comparisons."""
    out = np.zeros(len(hits), dtype=HIT16_DTYPE)
    m = mask.astype(bool)
    out["prim"] = RTK_CUDA_MISS
    out["t"][m] = hits["t"][m]
    out["u"][m] = hits["u"][m]
    out["v"][m] = hits["v"][m]
    mf = np.asarray(mesh_first, dtype=np.uint64)
    out["prim"][m] = (mf[hits["mesh_index"][m]] + hits["triangle_index"][m]).astype(np.uint32)
    return out
