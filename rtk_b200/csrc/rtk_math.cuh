// rtk_math.cuh -- per-ray constants, the watertight ray/triangle test and the conservative
// ray/box slab test.  The triangle arithmetic reproduces the reference bit for bit
// (rtk.c:543-567 setup, :256-354 test); every multiply/add is an explicit round-to-nearest
// intrinsic so nvcc cannot contract them into FMAs (the reference's SSE code has none).
#pragma once
#include "rtk_common.cuh"

struct rtk_ray_ctx {
	// triangle test (reference rtk.c:550-566)
	float ox, oy, oz;        // origin permuted to (kx,ky,kz)
	float sx, sy, sz;        // shear constants
	int   kz;                // dominant axis; kx=(kz+1)%3, ky=(kz+2)%3
	// box test: t = fma(plane, id, c) with padded near/far origins folded into c
	float idx, idy, idz;
	float cnx, cny, cnz;
	float cfx, cfy, cfz;
	uint32_t sgn;            // bit a: direction component a is negative (near plane = hi)
	float min_t;
};

#ifndef RTK_TRI_SUB_FIRST
#define RTK_TRI_SUB_FIRST 1     // 1: the triangle test subtracts the (unpermuted) origin before the axis permutation
#endif

RTK_DEV float rtk_fast_rcp(float x)
{
#ifdef RTK_SIMT_EMU
	return 1.0f / x;
#else
	return __fdividef(1.0f, x);
#endif
}

// select component k of (x,y,z)
RTK_DEV float rtk_sel3(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

RTK_DEV void rtk_ray_setup(rtk_ray_ctx &r, float ox, float oy, float oz, float dx, float dy, float dz,
                           float min_t, float scene_abs_max)
{
	// rtk.c:550-555: kz = argmax |d|, x wins ties, then y
	float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
	float mx = ax > ay ? ax : ay;
	mx = mx > az ? mx : az;
	int kz = (ax == mx) ? 0 : ((ay == mx) ? 1 : 2);
	int kx = kz == 2 ? 0 : kz + 1;
	int ky = kx == 2 ? 0 : kx + 1;
	float dkx = rtk_sel3(dx, dy, dz, kx), dky = rtk_sel3(dx, dy, dz, ky), dkz = rtk_sel3(dx, dy, dz, kz);
	r.kz = kz;
	r.sx = __fdiv_rn(-dkx, dkz);      // rtk.c:561
	r.sy = __fdiv_rn(-dky, dkz);      // rtk.c:562
	r.sz = __fdiv_rn(1.0f, dkz);      // rtk.c:563
#if RTK_TRI_SUB_FIRST
	r.ox = ox; r.oy = oy; r.oz = oz;  // unpermuted: rtk_tri_test translates first, then permutes
#else
	r.ox = rtk_sel3(ox, oy, oz, kx);  // rtk.c:564-566
	r.oy = rtk_sel3(ox, oy, oz, ky);
	r.oz = rtk_sel3(ox, oy, oz, kz);
#endif
	r.min_t = min_t;

	// Conservative box test (DESIGN.md "node test").  The triangle test rounds in fp32, so it
	// can accept a triangle whose exact geometry the exact ray misses by a few ulps of the
	// origin-to-vertex distance.  Every box is therefore padded by 64 ulps of S = max(|o|, |scene|)
	// and direction components below 2^-24 |d|max are raised to that value (the ray moves by
	// less than one ulp of S inside the scene), which also keeps 1/d finite.
	float thr = mx * 5.9604645e-08f;                       // 2^-24
	float ddx = copysignf(fmaxf(ax, thr), dx);
	float ddy = copysignf(fmaxf(ay, thr), dy);
	float ddz = copysignf(fmaxf(az, thr), dz);
	// approximate reciprocals (2 ulp) are enough here: the error is covered by the padding
	r.idx = rtk_fast_rcp(ddx); r.idy = rtk_fast_rcp(ddy); r.idz = rtk_fast_rcp(ddz);
	float S = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fmaxf(fabsf(oz), scene_abs_max));
	float pad = S * 3.8146973e-06f;                        // 64 * 2^-24
	r.cnx = -((ox + copysignf(pad, dx)) * r.idx);
	r.cny = -((oy + copysignf(pad, dy)) * r.idy);
	r.cnz = -((oz + copysignf(pad, dz)) * r.idz);
	r.cfx = -((ox - copysignf(pad, dx)) * r.idx);
	r.cfy = -((oy - copysignf(pad, dy)) * r.idy);
	r.cfz = -((oz - copysignf(pad, dz)) * r.idz);
	r.sgn = (__float_as_uint(dx) >> 31) | ((__float_as_uint(dy) >> 31) << 1) | ((__float_as_uint(dz) >> 31) << 2);
}

// The same setup in two halves, for the traversal kernel's distributed ray preparation: the 32 lanes
// of a warp each prepare ONE ray of a staged batch (the divides, full lane utilisation) into three
// float4, and the lanes that later own the ray finish the cheap rest.
//   q0 = (o.x, o.y, o.z, min_t)   q1 = (sx, sy, sz, bits: kz | sgn << 2)   q2 = (1/d.x, 1/d.y, 1/d.z, max_t)
RTK_DEV void rtk_ray_prepare(float4 r0, float4 r1, float4 &q0, float4 &q1, float4 &q2)
{
	const float ox = r0.x, oy = r0.y, oz = r0.z, dx = r0.w, dy = r1.x, dz = r1.y;
	float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
	float mx = ax > ay ? ax : ay;
	mx = mx > az ? mx : az;
	int kz = (ax == mx) ? 0 : ((ay == mx) ? 1 : 2);            // rtk.c:550-555
	int kx = kz == 2 ? 0 : kz + 1;
	int ky = kx == 2 ? 0 : kx + 1;
	float dkx = rtk_sel3(dx, dy, dz, kx), dky = rtk_sel3(dx, dy, dz, ky), dkz = rtk_sel3(dx, dy, dz, kz);
	float thr = mx * 5.9604645e-08f;
	float ddx = copysignf(fmaxf(ax, thr), dx);
	float ddy = copysignf(fmaxf(ay, thr), dy);
	float ddz = copysignf(fmaxf(az, thr), dz);
	uint32_t sgn = (__float_as_uint(dx) >> 31) | ((__float_as_uint(dy) >> 31) << 1) | ((__float_as_uint(dz) >> 31) << 2);
	q0 = make_float4(ox, oy, oz, r1.z);
	q1 = make_float4(__fdiv_rn(-dkx, dkz), __fdiv_rn(-dky, dkz), __fdiv_rn(1.0f, dkz), __uint_as_float((uint32_t)kz | (sgn << 2)));
	q2 = make_float4(rtk_fast_rcp(ddx), rtk_fast_rcp(ddy), rtk_fast_rcp(ddz), r1.w);
}

// node-test half of the context from a prepared ray (the triangle-test half stays in shared memory)
RTK_DEV void rtk_ray_node_ctx(rtk_ray_ctx &r, float4 q0, float4 q1, float4 q2, float scene_abs_max)
{
	const uint32_t bits = __float_as_uint(q1.w);
	r.kz = (int)(bits & 3u);
	r.sgn = bits >> 2;
	r.min_t = q0.w;
	r.idx = q2.x; r.idy = q2.y; r.idz = q2.z;
	float S = fmaxf(fmaxf(fabsf(q0.x), fabsf(q0.y)), fmaxf(fabsf(q0.z), scene_abs_max));
	float pad = S * 3.8146973e-06f;                        // 64 * 2^-24
	float px = (r.sgn & 1u) ? -pad : pad, py = (r.sgn & 2u) ? -pad : pad, pz = (r.sgn & 4u) ? -pad : pad;
	r.cnx = -((q0.x + px) * r.idx);
	r.cny = -((q0.y + py) * r.idy);
	r.cnz = -((q0.z + pz) * r.idz);
	r.cfx = -((q0.x - px) * r.idx);
	r.cfy = -((q0.y - py) * r.idy);
	r.cfz = -((q0.z - pz) * r.idz);
}

// triangle-test half from the shared-memory record (q0, q1 as above)
RTK_DEV void rtk_ray_tri_ctx(rtk_ray_ctx &r, float4 q0, float4 q1)
{
	const int kz = (int)(__float_as_uint(q1.w) & 3u);
	r.kz = kz;
#if RTK_TRI_SUB_FIRST
	r.ox = q0.x; r.oy = q0.y; r.oz = q0.z;
#else
	const int kx = kz == 2 ? 0 : kz + 1;
	const int ky = kx == 2 ? 0 : kx + 1;
	r.ox = rtk_sel3(q0.x, q0.y, q0.z, kx);                  // rtk.c:564-566
	r.oy = rtk_sel3(q0.x, q0.y, q0.z, ky);
	r.oz = rtk_sel3(q0.x, q0.y, q0.z, kz);
#endif
	r.sx = q1.x; r.sy = q1.y; r.sz = q1.z;
	r.min_t = q0.w;
}

// One triangle, reference rtk.c:256-354 for a single lane with the own-lane fp64 rule
// (SURVEY 8(c)): returns true and t,u,v when min_t < t <= max_t_incl.  The upper bound is
// inclusive so that the caller can resolve exact ties towards the lowest triangle number; the
// caller rejects t == max_t_incl when no hit has been recorded yet (rtk.c:354 is strict).
RTK_DEV bool rtk_tri_test(const rtk_ray_ctx &r, float4 p0, float4 p1, float4 p2, float max_t_incl,
                          float &t_out, float &u_out, float &v_out)
{
	const int kz = r.kz;
#if RTK_TRI_SUB_FIRST
	// translate (rtk.c:256-280) BEFORE the axis permutation: component-wise, so the same nine differences, and the
	// origin needs no permutation of its own
	p0.x = __fsub_rn(p0.x, r.ox); p0.y = __fsub_rn(p0.y, r.oy); p0.z = __fsub_rn(p0.z, r.oz);
	p1.x = __fsub_rn(p1.x, r.ox); p1.y = __fsub_rn(p1.y, r.oy); p1.z = __fsub_rn(p1.z, r.oz);
	p2.x = __fsub_rn(p2.x, r.ox); p2.y = __fsub_rn(p2.y, r.oy); p2.z = __fsub_rn(p2.z, r.oz);
#endif
	// axis permutation (the reference's pshufb, rtk.c:232-243)
	float a0x = kz == 0 ? p0.y : (kz == 1 ? p0.z : p0.x);
	float a0y = kz == 0 ? p0.z : (kz == 1 ? p0.x : p0.y);
	float a0z = kz == 0 ? p0.x : (kz == 1 ? p0.y : p0.z);
	float a1x = kz == 0 ? p1.y : (kz == 1 ? p1.z : p1.x);
	float a1y = kz == 0 ? p1.z : (kz == 1 ? p1.x : p1.y);
	float a1z = kz == 0 ? p1.x : (kz == 1 ? p1.y : p1.z);
	float a2x = kz == 0 ? p2.y : (kz == 1 ? p2.z : p2.x);
	float a2y = kz == 0 ? p2.z : (kz == 1 ? p2.x : p2.y);
	float a2z = kz == 0 ? p2.x : (kz == 1 ? p2.y : p2.z);
#if !RTK_TRI_SUB_FIRST
	// translate, rtk.c:256-280
	a0x = __fsub_rn(a0x, r.ox); a0y = __fsub_rn(a0y, r.oy); a0z = __fsub_rn(a0z, r.oz);
	a1x = __fsub_rn(a1x, r.ox); a1y = __fsub_rn(a1y, r.oy); a1z = __fsub_rn(a1z, r.oz);
	a2x = __fsub_rn(a2x, r.ox); a2y = __fsub_rn(a2y, r.oy); a2z = __fsub_rn(a2z, r.oz);
#endif
	// shear, rtk.c:284-292 (separate multiply and add)
	float x0 = __fadd_rn(a0x, __fmul_rn(r.sx, a0z));
	float y0 = __fadd_rn(a0y, __fmul_rn(r.sy, a0z));
	float z0 = __fmul_rn(r.sz, a0z);
	float x1 = __fadd_rn(a1x, __fmul_rn(r.sx, a1z));
	float y1 = __fadd_rn(a1y, __fmul_rn(r.sy, a1z));
	float z1 = __fmul_rn(r.sz, a1z);
	float x2 = __fadd_rn(a2x, __fmul_rn(r.sx, a2z));
	float y2 = __fadd_rn(a2y, __fmul_rn(r.sy, a2z));
	float z2 = __fmul_rn(r.sz, a2z);
	// edge functions, rtk.c:298-300
	float u = __fsub_rn(__fmul_rn(x1, y2), __fmul_rn(y1, x2));
	float v = __fsub_rn(__fmul_rn(x2, y0), __fmul_rn(y2, x0));
	float w = __fsub_rn(__fmul_rn(x0, y1), __fmul_rn(y0, x1));
	// exact-zero fallback in double, rtk.c:301-336
	if (u == 0.0f || v == 0.0f || w == 0.0f) {
		double ud = __dsub_rn(__dmul_rn((double)x1, (double)y2), __dmul_rn((double)y1, (double)x2));
		double vd = __dsub_rn(__dmul_rn((double)x2, (double)y0), __dmul_rn((double)y2, (double)x0));
		double wd = __dsub_rn(__dmul_rn((double)x0, (double)y1), __dmul_rn((double)y0, (double)x1));
		u = __double2float_rn(ud); v = __double2float_rn(vd); w = __double2float_rn(wd);
	}
	// sign test, rtk.c:340-344
	bool neg = (u < 0.0f) || (v < 0.0f) || (w < 0.0f);
	bool pos = (u > 0.0f) || (v > 0.0f) || (w > 0.0f);
	if (neg && pos) return false;
	// rtk.c:346-353
	float det = __fadd_rn(__fadd_rn(u, v), w);
	float rcp = __fdiv_rn(1.0f, det);
	float z = __fmul_rn(u, z0);
	z = __fadd_rn(z, __fmul_rn(v, z1));
	z = __fadd_rn(z, __fmul_rn(w, z2));
	float t = __fmul_rn(z, rcp);
	if (!(t > r.min_t && t <= max_t_incl)) return false;  // rtk.c:354
	t_out = t;
	u_out = __fmul_rn(u, rcp);                            // rtk.c:363
	v_out = __fmul_rn(v, rcp);                            // rtk.c:364
	return true;
}
