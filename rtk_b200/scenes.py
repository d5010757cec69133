"""Deterministic synthetic scenes and ray sets for the BASELINE.json configs.

Everything is generated on the host from a counter-based hash, splitmix64(seed ^ counter) with
float = (x >> 40) * 2**-24, so that the CPU oracle, the reference build and the GPU see the same
bytes and any element can be regenerated independently of the others.

A scene is a dict:
    meshes      list of {"positions": (nv,3) float32, "indices": (nt,3) uint16/uint32 or None}
    tris        (N,3,3) float32 -- all triangles, meshes concatenated (the global numbering of
                the reference's build items, rtk.c:1131-1170)
    mesh_first  (num_meshes+1,) uint32 -- first global triangle number of each mesh
Rays are numpy structured arrays with the layout of rtk_ray (rtk.h:29-34).
"""
import numpy as np

RTK_INF = np.float32(3.402823e+38)
RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("min_t", "<f4"), ("max_t", "<f4")])

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def splitmix64(x):
    """Vectorised splitmix64 finaliser over uint64 arrays."""
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=np.uint64) + _M1
        z = (z ^ (z >> np.uint64(30))) * _M2
        z = (z ^ (z >> np.uint64(27))) * _M3
        return z ^ (z >> np.uint64(31))


def u01(seed, counter):
    """float32 in [0,1): (splitmix64(seed ^ counter) >> 40) * 2**-24."""
    h = splitmix64(np.uint64(seed) ^ np.asarray(counter, dtype=np.uint64))
    return ((h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def _stream(seed, n, k, streams):
    """u01 for element indices [0,n) of stream k out of `streams` per element."""
    idx = np.arange(n, dtype=np.uint64) * np.uint64(streams) + np.uint64(k)
    return u01(seed, idx)


def _finish(meshes):
    tris, first = [], [0]
    for m in meshes:
        p = m["positions"]
        if m["indices"] is None:
            t = p.reshape(-1, 3, 3)
        else:
            t = p[m["indices"].astype(np.int64)]
        tris.append(np.ascontiguousarray(t, dtype=np.float32))
        first.append(first[-1] + len(t))
    return {"meshes": meshes, "tris": np.concatenate(tris, axis=0),
            "mesh_first": np.asarray(first, dtype=np.uint32)}


def make_rays(o, d, min_t=0.0, max_t=RTK_INF):
    n = len(o)
    r = np.zeros(n, dtype=RAY_DTYPE)
    r["o"] = o
    r["d"] = d
    r["min_t"] = min_t
    r["max_t"] = max_t
    return r


# --------------------------------------------------------------------------------------------
# C1: Cornell box, 992 triangles, one mesh, U16 indices
# --------------------------------------------------------------------------------------------

def _quad_grid(p0, du, dv, nu, nv):
    """(nu+1)*(nv+1) vertices p0 + i*du/nu + j*dv/nv and 2*nu*nv index triples."""
    i, j = np.meshgrid(np.arange(nu + 1), np.arange(nv + 1), indexing="ij")
    p = (np.asarray(p0, np.float64)[None, None, :]
         + (i[..., None] / nu) * np.asarray(du, np.float64)
         + (j[..., None] / nv) * np.asarray(dv, np.float64))
    verts = p.reshape(-1, 3).astype(np.float32)
    idx = []
    for a in range(nu):
        for b in range(nv):
            v00 = a * (nv + 1) + b
            v01 = v00 + 1
            v10 = v00 + (nv + 1)
            v11 = v10 + 1
            idx.append((v00, v10, v11))
            idx.append((v00, v11, v01))
    return verts, np.asarray(idx, dtype=np.uint32)


def cornell_box():
    """5 walls x 8x8x2 + 2 boxes x 5 faces x 4x4x2 + light 4x4x2 = 992 triangles."""
    X, Y, Z = 556.0, 548.0, 559.0
    quads = [
        ((0, 0, 0), (X, 0, 0), (0, 0, Z), 8),        # floor
        ((0, Y, 0), (X, 0, 0), (0, 0, Z), 8),        # ceiling
        ((0, 0, Z), (X, 0, 0), (0, Y, 0), 8),        # back wall
        ((0, 0, 0), (0, 0, Z), (0, Y, 0), 8),        # right wall (x = 0)
        ((X, 0, 0), (0, 0, Z), (0, Y, 0), 8),        # left wall  (x = X)
    ]

    def box(x0, x1, y1, z0, z1):
        return [
            ((x0, y1, z0), (x1 - x0, 0, 0), (0, 0, z1 - z0), 4),    # top
            ((x0, 0, z0), (x1 - x0, 0, 0), (0, y1, 0), 4),          # front
            ((x0, 0, z1), (x1 - x0, 0, 0), (0, y1, 0), 4),          # back
            ((x0, 0, z0), (0, 0, z1 - z0), (0, y1, 0), 4),          # side
            ((x1, 0, z0), (0, 0, z1 - z0), (0, y1, 0), 4),          # side
        ]
    quads += box(130.0, 295.0, 165.0, 65.0, 230.0)
    quads += box(265.0, 430.0, 330.0, 295.0, 460.0)
    quads.append(((213.0, Y - 0.5, 227.0), (130.0, 0, 0), (0, 0, 105.0), 4))   # light
    verts, idx, base = [], [], 0
    for p0, du, dv, n in quads:
        v, i = _quad_grid(p0, du, dv, n, n)
        verts.append(v)
        idx.append(i + base)
        base += len(v)
    positions = np.concatenate(verts, axis=0)
    indices = np.concatenate(idx, axis=0).astype(np.uint16)
    assert len(indices) == 992
    return _finish([{"positions": positions, "indices": indices}])


def pinhole_rays(width, height, eye, fov_deg, jitter_seed=None):
    """Row-major pixel-centre primary rays looking down +z (vertical fov)."""
    py, px = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32), indexing="ij")
    if jitter_seed is None:
        jx = jy = np.float32(0.5)
    else:
        n = width * height
        jx = _stream(jitter_seed, n, 0, 2).reshape(height, width)
        jy = _stream(jitter_seed, n, 1, 2).reshape(height, width)
    th = np.float32(np.tan(np.radians(fov_deg) / 2.0))
    aspect = np.float32(width / height)
    dx = ((px + jx) / np.float32(width) * 2 - 1) * th * aspect
    dy = (1 - (py + jy) / np.float32(height) * 2) * th
    d = np.stack([dx, dy, np.ones_like(dx)], axis=-1).reshape(-1, 3).astype(np.float32)
    o = np.broadcast_to(np.asarray(eye, dtype=np.float32), d.shape)
    return make_rays(o, d)


def cornell_rays(width=512, height=512):
    return pinhole_rays(width, height, (278.0, 273.0, -800.0), 39.3)


# --------------------------------------------------------------------------------------------
# C2: random triangle soup
# --------------------------------------------------------------------------------------------

def soup(ntris=1_000_000, seed=0xC2, size=0.01):
    """centroid ~ U[0,1)^3, vertices = centroid + size * U[-1,1)^3; non-indexed float32."""
    c = np.stack([_stream(seed, ntris, k, 12) for k in range(3)], axis=-1)            # (N,3)
    off = np.stack([_stream(seed, ntris, 3 + k, 12) for k in range(9)], axis=-1)       # (N,9)
    off = (off * np.float32(2) - np.float32(1)) * np.float32(size)
    tris = (c[:, None, :] + off.reshape(ntris, 3, 3)).astype(np.float32)
    return _finish([{"positions": np.ascontiguousarray(tris.reshape(-1, 3)), "indices": None}])


def soup_primary_rays(width=1920, height=1080):
    return pinhole_rays(width, height, (0.5, 0.5, -1.5), 40.0)


# --------------------------------------------------------------------------------------------
# C3 / C4: procedural terrain (value-noise heightfield)
# --------------------------------------------------------------------------------------------

def _value_noise(seed, x, z, octaves=5, base_freq=16.0):
    """5-octave value noise in [0,1): hashed lattice values, smoothstep interpolation."""
    out = np.zeros(x.shape, dtype=np.float64)
    amp, freq, norm = 0.5, base_freq, 0.0
    for o in range(octaves):
        fx, fz = x * freq, z * freq
        ix, iz = np.floor(fx).astype(np.int64), np.floor(fz).astype(np.int64)
        tx, tz = fx - ix, fz - iz
        tx = tx * tx * (3 - 2 * tx)
        tz = tz * tz * (3 - 2 * tz)

        def lat(a, b):
            key = ((a.astype(np.uint64) & np.uint64(0xFFFFF)) << np.uint64(24)) ^ \
                  ((b.astype(np.uint64) & np.uint64(0xFFFFF)) << np.uint64(4)) ^ np.uint64(o)
            return u01(seed, key).astype(np.float64)
        v00, v10, v01, v11 = lat(ix, iz), lat(ix + 1, iz), lat(ix, iz + 1), lat(ix + 1, iz + 1)
        out += amp * ((v00 * (1 - tx) + v10 * tx) * (1 - tz) + (v01 * (1 - tx) + v11 * tx) * tz)
        norm += amp
        amp *= 0.5
        freq *= 2.0
    return out / norm


def _grid_indices(nx, nz, row0=0, row1=None):
    """index triples of cells [row0,row1) x [0,nx) of an (nz+1) x (nx+1) vertex grid."""
    row1 = nz if row1 is None else row1
    r, c = np.meshgrid(np.arange(row0, row1, dtype=np.int64), np.arange(nx, dtype=np.int64), indexing="ij")
    v00 = (r * (nx + 1) + c).reshape(-1)
    v01 = v00 + 1
    v10 = v00 + (nx + 1)
    v11 = v10 + 1
    t = np.empty((len(v00), 2, 3), dtype=np.int64)
    t[:, 0] = np.stack([v00, v10, v11], axis=-1)
    t[:, 1] = np.stack([v00, v11, v01], axis=-1)
    return t.reshape(-1, 3)


def terrain(nx=1000, nz=500, seed=0xC3, amplitude=0.15, meshes=1):
    """(nx+1) x (nz+1) vertex heightfield, 2*nx*nz triangles, U32 indices.  x in [0, nx/nz], z in
    [0,1], y = amplitude * noise.  meshes=2 splits the rows into two meshes, each with its own
    vertex buffer and local indices (exercises mesh_index / triangle_index)."""
    ext_x = nx / nz
    gz, gx = np.meshgrid(np.arange(nz + 1, dtype=np.float64) / nz, np.arange(nx + 1, dtype=np.float64) / nz, indexing="ij")
    h = amplitude * _value_noise(seed, gx, gz)
    pos = np.stack([gx, h, gz], axis=-1).reshape(-1, 3).astype(np.float32)
    assert abs(pos[:, 0].max() - ext_x) < 1e-6
    out = []
    bounds = [round(nz * k / meshes) for k in range(meshes + 1)]
    for k in range(meshes):
        r0, r1 = bounds[k], bounds[k + 1]
        idx = _grid_indices(nx, nz, r0, r1)
        vbase = r0 * (nx + 1)
        vend = (r1 + 1) * (nx + 1)
        out.append({"positions": np.ascontiguousarray(pos[vbase:vend]),
                    "indices": (idx - vbase).astype(np.uint32)})
    return _finish(out)


def _normalize(v):
    n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=-1, keepdims=True))
    return (v / np.maximum(n, 1e-30)).astype(np.float32)


def bounce_rays(scene, n, seed=0xD3, first=0, up=(0.0, 1.0, 0.0)):
    """Diffuse-bounce rays [first, first+n): triangle ~U, point ~U(barycentric), origin = p +
    1e-3 * n_hat, direction cosine-weighted about n_hat (geometric normal flipped towards `up`).
    Generation order is the ray order (incoherent)."""
    tris = scene["tris"]
    N = len(tris)
    idx = np.arange(first, first + n, dtype=np.uint64)

    def s(k):
        return u01(seed, idx * np.uint64(5) + np.uint64(k))
    ti = np.minimum((s(0).astype(np.float64) * N).astype(np.int64), N - 1)
    r1, r2 = s(1).astype(np.float64), s(2).astype(np.float64)
    sq = np.sqrt(r1)
    b0, b1 = 1 - sq, sq * (1 - r2)
    b2 = 1 - b0 - b1
    t = tris[ti].astype(np.float64)
    p = b0[:, None] * t[:, 0] + b1[:, None] * t[:, 1] + b2[:, None] * t[:, 2]
    nrm = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    ln = np.sqrt((nrm ** 2).sum(-1, keepdims=True))
    nrm = np.where(ln > 0, nrm / np.maximum(ln, 1e-300), np.asarray(up)[None, :])
    flip = (nrm * np.asarray(up)[None, :]).sum(-1) < 0
    nrm[flip] *= -1
    # orthonormal basis
    a = np.where(np.abs(nrm[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    tx = np.cross(a, nrm)
    tx /= np.sqrt((tx ** 2).sum(-1, keepdims=True))
    ty = np.cross(nrm, tx)
    e1, e2 = s(3).astype(np.float64), s(4).astype(np.float64)
    rr, ph = np.sqrt(e1), 2 * np.pi * e2
    lx, ly, lz = rr * np.cos(ph), rr * np.sin(ph), np.sqrt(np.maximum(0.0, 1 - e1))
    d = lx[:, None] * tx + ly[:, None] * ty + lz[:, None] * nrm
    o = p + 1e-3 * nrm
    return make_rays(o.astype(np.float32), d.astype(np.float32))


def scene_bounds(scene):
    """(lo, hi) float32 corners of the scene's bounding box; computed once per scene dict"""
    if "_bounds" not in scene:
        t = scene["tris"].reshape(-1, 3)
        scene["_bounds"] = (t.min(0), t.max(0))
    return scene["_bounds"]


def segment_rays(scene, n, seed=0xD4, first=0, length=0.25):
    """Uniform segments: origin ~U(bbox), direction ~U(sphere) * length * diag, max_t = 1."""
    lo, hi = scene_bounds(scene)
    lo, hi = lo.astype(np.float64), hi.astype(np.float64)
    idx = np.arange(first, first + n, dtype=np.uint64)

    def s(k):
        return u01(seed, idx * np.uint64(5) + np.uint64(k)).astype(np.float64)
    o = lo + np.stack([s(0), s(1), s(2)], -1) * (hi - lo)
    z = 2 * s(3) - 1
    ph = 2 * np.pi * s(4)
    r = np.sqrt(np.maximum(0.0, 1 - z * z))
    d = np.stack([r * np.cos(ph), z, r * np.sin(ph)], -1) * length * np.sqrt(((hi - lo) ** 2).sum())
    return make_rays(o.astype(np.float32), d.astype(np.float32), 0.0, 1.0)


def terrain_primary_rays(scene, width, height, first=0, count=None):
    """Coherent camera rays looking down onto the terrain from above one corner: pixels
    [first, first + count) of the width x height image, row-major (default: all of them)."""
    lo, hi = scene_bounds(scene)
    c = (lo + hi) / 2
    eye = np.array([c[0], hi[1] + 0.8 * (hi[2] - lo[2]), lo[2] - 0.6 * (hi[2] - lo[2])], dtype=np.float32)
    fwd = c - eye
    fwd = fwd / np.linalg.norm(fwd)
    right = np.cross([0.0, 1.0, 0.0], fwd)
    right /= np.linalg.norm(right)
    upv = np.cross(fwd, right)
    if count is None:
        count = width * height - first
    pix = np.arange(first, first + count, dtype=np.int64)
    py, px = pix // width, pix % width
    th = np.tan(np.radians(50.0) / 2)
    sx = ((px + 0.5) / width * 2 - 1) * th * (width / height)
    sy = (1 - (py + 0.5) / height * 2) * th
    d = fwd[None, :] + sx[..., None] * right + sy[..., None] * upv
    d = d.reshape(-1, 3).astype(np.float32)
    return make_rays(np.broadcast_to(eye, d.shape), d)


def mixed_rays(scene, n, seed=0xD4, block=65536, first=0, count=None, threads=1):
    """C4: thirds of coherent primary / bounce / segment rays interleaved in blocks of `block`.
    Rays [first, first + count) of the n-ray set (first a multiple of `block`; default: all of them):
    every block regenerates alone, so ranks and worker threads can each make their own part."""
    nb = (n + block - 1) // block
    if count is None:
        count = n - first
    assert first % block == 0 and first + count <= n
    out = np.zeros(count, dtype=RAY_DTYPE)
    n_prim = sum(min(block, n - b * block) for b in range(0, nb, 3))
    side = int(np.ceil(np.sqrt(max(n_prim, 1))))
    scene_bounds(scene)

    def one(b):
        lo, hi = b * block, min(n, (b + 1) * block, first + count)
        k = b % 3
        if k == 0:
            pp = sum(min(block, n - bb * block) for bb in range(0, b, 3))
            r = terrain_primary_rays(scene, side, side, pp, hi - lo)
        elif k == 1:
            r = bounce_rays(scene, hi - lo, seed=seed, first=lo)
        else:
            r = segment_rays(scene, hi - lo, seed=seed ^ 0x55, first=lo)
        out[lo - first:hi - first] = r
    blocks = range(first // block, (first + count + block - 1) // block)
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, blocks))
    else:
        for b in blocks:
            one(b)
    return out


# --------------------------------------------------------------------------------------------
# named configs
# --------------------------------------------------------------------------------------------

def config_scene(name, scale=1.0):
    """Scene of BASELINE.json config `name` in {"C1","C2","C3","C4"}; scale < 1 shrinks the
    triangle count (tests), keeping the generator and seeds."""
    if name == "C1":
        return cornell_box()
    if name == "C2":
        return soup(max(4, int(1_000_000 * scale)), seed=0xC2)
    if name == "C3":
        s = np.sqrt(scale)
        return terrain(max(2, int(round(1000 * s))), max(1, int(round(500 * s))), seed=0xC3)
    if name == "C4":
        s = np.sqrt(scale)
        return terrain(max(2, int(round(2500 * s))), max(2, int(round(2000 * s))), seed=0xC4,
                       amplitude=0.25, meshes=2)
    raise ValueError(name)


def config_rays(name, scene, n=None):
    if name == "C1":
        return cornell_rays()
    if name == "C2":
        return soup_primary_rays()
    if name == "C3":
        return bounce_rays(scene, 16_777_216 if n is None else n, seed=0xD3)
    if name == "C4":
        return mixed_rays(scene, 67_108_864 if n is None else n, seed=0xD4)
    raise ValueError(name)
