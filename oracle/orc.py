"""ctypes bindings of the CPU oracle (oracle/liborc.so) and of the reference builds
under oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (rtk_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBORC = os.path.join(HERE, "liborc.so")
REF_SO = os.path.join(HERE, "_ref", "librtk_ref.so")
REF_PATCHED_SO = os.path.join(HERE, "_ref", "librtk_ref_patched.so")

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("min_t", "<f4"), ("max_t", "<f4")])
HIT16_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])
MISS = 0xFFFFFFFF


class BlobStats(C.Structure):
    _fields_ = [("size_bytes", C.c_uint64), ("num_nodes4", C.c_uint64), ("num_leaves", C.c_uint64),
                ("num_build_nodes", C.c_uint64), ("max_depth", C.c_uint32), ("build_seconds", C.c_double)]


def build(force=False):
    """Compile liborc.so and (when /root/reference exists) oracle/_ref/*.so."""
    if force or not os.path.exists(LIBORC) or \
            os.path.getmtime(LIBORC) < os.path.getmtime(os.path.join(HERE, "rtk_oracle.c")) or \
            (os.path.exists("/root/reference/rtk.c") and not os.path.exists(REF_PATCHED_SO)):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIBORC)
        L.orc_trace_brute.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.orc_trace_brute.restype = None
        L.orc_trace_brute_blocked.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.orc_trace_brute_blocked.restype = C.c_int
        L.orc_ray_triangle.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.POINTER(C.c_float),
                                       C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_ray_triangle.restype = C.c_int
        L.orc_build_reference_blob.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.POINTER(BlobStats)]
        L.orc_build_reference_blob.restype = C.c_void_p
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free.restype = None
        L.orc_trace_flat_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                               C.c_void_p, C.c_int]
        L.orc_trace_flat_reference.restype = None
        L.orc_trace_reference_blob.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                               C.c_void_p, C.c_int]
        L.orc_trace_reference_blob.restype = C.c_double
        L.orc_num_cores.restype = C.c_int
        _lib = L
    return _lib


def num_cores():
    return int(lib().orc_num_cores())


def _tri9(tris):
    a = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9)
    return a


def _rays(rays):
    a = np.ascontiguousarray(rays)
    assert a.dtype == RAY_DTYPE, a.dtype
    return a


def trace_brute(tris, rays, threads=0, blocked=None):
    """Canonical brute force: (ntris,3,3) float32, rays RAY_DTYPE -> HIT16_DTYPE array.
    Big jobs (>= 2^26 ray-triangle pairs) take the blocked, pre-filtered loop, which gives the same
    answers (the scalar code stays the arbiter of every hit; tests/test_oracle.py compares the two)."""
    t9, r = _tri9(tris), _rays(rays)
    out = np.zeros(len(r), dtype=HIT16_DTYPE)
    if blocked is None:
        blocked = len(t9) * len(r) >= (1 << 26)
    if blocked:
        lib().orc_trace_brute_blocked(t9.ctypes.data, len(t9), r.ctypes.data, len(r), out.ctypes.data, threads)
    else:
        lib().orc_trace_brute(t9.ctypes.data, len(t9), r.ctypes.data, len(r), out.ctypes.data, threads)
    return out


_ref_handles = {}


def have_reference(patched=False):
    return os.path.exists(REF_PATCHED_SO if patched else REF_SO)


def _ref_fn(patched):
    path = REF_PATCHED_SO if patched else REF_SO
    if path not in _ref_handles:
        build()
        h = C.CDLL(path)
        _ref_handles[path] = h
    return C.cast(_ref_handles[path].rtk_trace_ray, C.c_void_p)


def trace_flat_reference(tris, rays, threads=0):
    """Brute force through the UNMODIFIED reference leaf code (60-triangle flat blobs)."""
    t9, r = _tri9(tris), _rays(rays)
    out = np.zeros(len(r), dtype=HIT16_DTYPE)
    lib().orc_trace_flat_reference(_ref_fn(False), t9.ctypes.data, len(t9), r.ctypes.data, len(r),
                                   out.ctypes.data, threads)
    return out


class ReferenceBlob:
    """A scene in the reference's blob format, built by the oracle's binned-SAH restatement."""

    def __init__(self, tris, mesh_of=None, tri_in_mesh=None, vidx3=None):
        t9 = _tri9(tris)
        self.stats = BlobStats()
        keep = []

        def p(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.uint32)
            keep.append(a)
            return a.ctypes.data
        self.ptr = lib().orc_build_reference_blob(t9.ctypes.data, len(t9), p(mesh_of), p(tri_in_mesh),
                                                  p(vidx3), C.byref(self.stats))
        if not self.ptr:
            raise MemoryError("orc_build_reference_blob failed")
        self.ntris = len(t9)

    def trace(self, rays, patched=True, threads=0, mesh_first=None, want_hits=True):
        """rtk_trace_ray of the (patched) reference build over every ray; returns (hits, seconds)."""
        r = _rays(rays)
        out = np.zeros(len(r), dtype=HIT16_DTYPE) if want_hits else None
        mf = None if mesh_first is None else np.ascontiguousarray(mesh_first, dtype=np.uint32)
        sec = lib().orc_trace_reference_blob(_ref_fn(patched), self.ptr, r.ctypes.data, len(r),
                                             out.ctypes.data if want_hits else None,
                                             mf.ctypes.data if mf is not None else None, threads)
        return out, sec

    def close(self):
        if self.ptr:
            lib().orc_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
