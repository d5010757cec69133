"""numpy restatement of the wavefront ray generators (rtk_b200/csrc/k_wavefront.cuh).

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  The reference has no ray generator --
rtk.h only answers ray queries -- so this file is the specification of the two device kernels
that sit either side of the trace in a path-tracing wavefront:

  primary_rays   float32 arithmetic, one rounding per operation, same order as k_gen_primary:
                 the device result must match bit for bit.
  bounce_rays    float32 arithmetic in the order of k_gen_bounce; the device contracts
                 multiply-adds and uses its own sqrt/sin/cos, so the comparison is 1e-5 relative
                 (tests/parity_cases.py::case_wavefront).
"""
import numpy as np

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("min_t", "<f4"), ("max_t", "<f4")])
RTK_INF = np.float32(3.402823e+38)
MISS = 0xFFFFFFFF
RELAUNCH = 1

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)
F = np.float32


def _splitmix64(x):
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=np.uint64) + _M1
        z = (z ^ (z >> np.uint64(30))) * _M2
        z = (z ^ (z >> np.uint64(27))) * _M3
        return z ^ (z >> np.uint64(31))


def _u01(seed, counter):
    h = _splitmix64(np.uint64(seed) ^ np.asarray(counter, dtype=np.uint64))
    return ((h >> np.uint64(40)).astype(np.float32) * F(2.0 ** -24)).astype(np.float32)


def primary_rays(eye, forward, right, up, tan_half_fov, width, height, seed, sample, first_pixel, count):
    """k_gen_primary: jittered pinhole rays of pixels [first_pixel, first_pixel+count), row-major."""
    p = np.arange(first_pixel, first_pixel + count, dtype=np.uint64)
    px = (p % np.uint64(width)).astype(np.float32)
    py = (p // np.uint64(width)).astype(np.float32)
    ctr = (p * np.uint64(64) + np.uint64(sample)) * np.uint64(2)
    jx, jy = _u01(seed, ctr), _u01(seed, ctr + np.uint64(1))
    w, h, tan = F(width), F(height), F(tan_half_fov)
    aspect = F(w / h)
    fx = ((px + jx) / w).astype(np.float32)
    fy = ((py + jy) / h).astype(np.float32)
    sx = (((fx * F(2) - F(1)).astype(np.float32) * tan).astype(np.float32) * aspect).astype(np.float32)
    sy = ((F(1) - (fy * F(2)).astype(np.float32)).astype(np.float32) * tan).astype(np.float32)
    rays = np.zeros(count, dtype=RAY_DTYPE)
    rays["o"] = np.asarray(eye, dtype=np.float32)
    for k in range(3):
        a = (F(forward[k]) + (sx * F(right[k])).astype(np.float32)).astype(np.float32)
        rays["d"][:, k] = (a + (sy * F(up[k])).astype(np.float32)).astype(np.float32)
    rays["min_t"] = 0
    rays["max_t"] = RTK_INF
    return rays


def bounce_rays(tris, rays_in, hit16, seed, bounce, first_ray, push, flags=0):
    """k_gen_bounce.  tris: (N,3,3) float32 in original order; hit16: records with t and prim.
    Returns (rays, alive)."""
    tris = np.asarray(tris, dtype=np.float32)
    n = len(rays_in)
    i = np.arange(n, dtype=np.uint64)
    ctr = ((np.uint64(first_ray) + i) * np.uint64(16) + np.uint64(bounce)) * np.uint64(8) + np.uint64(0x5bd1e995)
    prim = hit16["prim"].astype(np.int64)
    miss = hit16["prim"] == MISS
    state = np.where(miss, 0, 1).astype(np.uint8)
    N = len(tris)
    if (flags & RELAUNCH) and N:
        pick = np.minimum((_u01(seed, ctr + np.uint64(2)) * F(N)).astype(np.float32).astype(np.uint32), N - 1)
        prim = np.where(miss, pick, prim)
        state[miss] = 2
    live = state != 0
    out = np.zeros(n, dtype=RAY_DTYPE)
    out["d"][~live] = (0, 0, 1)                     # float4 (0, 1, 0, 0): d.y = 0, d.z = 1, min_t = max_t = 0
    if not live.any():
        return out, state
    pr = np.where(live, prim, 0)
    a, b, c = tris[pr, 0], tris[pr, 1], tris[pr, 2]
    e1, e2 = (b - a).astype(np.float32), (c - a).astype(np.float32)
    nrm = np.cross(e1, e2).astype(np.float32)
    ln = np.sqrt((nrm * nrm).sum(-1, dtype=np.float32)).astype(np.float32)
    ok = ln > 0
    nrm = np.where(ok[:, None], nrm / np.where(ok, ln, 1)[:, None], np.array([[0, 1, 0]], dtype=np.float32)).astype(np.float32)
    d_in = rays_in["d"]
    t = hit16["t"].astype(np.float32)
    p_hit = (rays_in["o"] + t[:, None] * d_in).astype(np.float32)
    flip_hit = (nrm * d_in).sum(-1, dtype=np.float32) > 0
    sq = np.sqrt(_u01(seed, ctr + np.uint64(3))).astype(np.float32)
    r2 = _u01(seed, ctr + np.uint64(4))
    b0 = (F(1) - sq).astype(np.float32)
    b1 = (sq * (F(1) - r2)).astype(np.float32)
    b2 = (F(1) - b0 - b1).astype(np.float32)
    p_new = (b0[:, None] * a + b1[:, None] * b + b2[:, None] * c).astype(np.float32)
    flip_new = nrm[:, 1] < 0
    relaunched = state == 2
    p = np.where(relaunched[:, None], p_new, p_hit)
    flip = np.where(relaunched, flip_new, flip_hit)
    nrm = np.where(flip[:, None], -nrm, nrm)
    big = np.abs(nrm[:, 0]) > F(0.9)
    ax, ay = np.where(big, F(0), F(1)), np.where(big, F(1), F(0))
    tx = np.stack([ay * nrm[:, 2], -ax * nrm[:, 2], ax * nrm[:, 1] - ay * nrm[:, 0]], -1).astype(np.float32)
    tx = (tx / np.sqrt((tx * tx).sum(-1, dtype=np.float32))[:, None]).astype(np.float32)
    bt = np.cross(nrm, tx).astype(np.float32)
    u1, u2 = _u01(seed, ctr), _u01(seed, ctr + np.uint64(1))
    r, phi = np.sqrt(u1), (F(6.2831853) * u2).astype(np.float32)
    lx, ly = (r * np.cos(phi)).astype(np.float32), (r * np.sin(phi)).astype(np.float32)
    lz = np.sqrt(np.maximum(F(0), F(1) - u1)).astype(np.float32)
    d = (lx[:, None] * tx + ly[:, None] * bt + lz[:, None] * nrm).astype(np.float32)
    o = (p + F(push) * nrm).astype(np.float32)
    out["o"][live] = o[live]
    out["d"][live] = d[live]
    out["max_t"][live] = RTK_INF
    return out, state
