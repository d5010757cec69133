// k_wavefront.cuh -- ray generation on the device, so that a path-tracing wavefront never leaves
// HBM (BASELINE.json config 5).  Not part of the reference, which only answers
// single-ray queries: these kernels sit either side of k_trace.
//
//   k_gen_primary   jittered pinhole camera rays for a range of pixels
//   k_gen_bounce    next-bounce rays from the compact hit records of the previous bounce:
//                   origin = hit point pushed off the surface along the geometric normal,
//                   direction = cosine-weighted hemisphere sample about that normal
//
// Random numbers are counter based -- splitmix64(seed ^ counter) >> 40 scaled by 2^-24, the same
// hash rtk_b200/scenes.py uses -- so every ray can be regenerated independently and the numpy
// restatement in oracle/wavefront_ref.py sees the same uniforms.
#pragma once
#include "rtk_common.cuh"

struct rtkd_camera {
	float eye[3], forward[3], right[3], up[3];
	float tan_half_fov;          // vertical
	uint32_t width, height;
};

RTK_DEV unsigned long long rtk_splitmix64(unsigned long long x)
{
	unsigned long long z = x + 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
RTK_DEV float rtk_u01(unsigned long long seed, unsigned long long counter)
{
	return __fmul_rn((float)(unsigned)(rtk_splitmix64(seed ^ counter) >> 40), 5.9604645e-08f);
}

// pixel p of the frame, sample index `sample`: explicit round-to-nearest fp32 operations (no FMA
// contraction) so that the numpy float32 restatement is bit-identical
__global__ void k_gen_primary(rtkd_camera cam, unsigned long long seed, uint32_t sample,
                              unsigned long long first_pixel, uint32_t count, float4 *rays)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	unsigned long long p = first_pixel + i;
	uint32_t px = (uint32_t)(p % cam.width), py = (uint32_t)(p / cam.width);
	unsigned long long ctr = (p * 64ull + sample) * 2ull;
	float jx = rtk_u01(seed, ctr), jy = rtk_u01(seed, ctr + 1);
	float aspect = __fdiv_rn((float)cam.width, (float)cam.height);
	// sx = ((px + jx) / width * 2 - 1) * tan * aspect ;  sy = (1 - (py + jy) / height * 2) * tan
	float fx = __fdiv_rn(__fadd_rn((float)px, jx), (float)cam.width);
	float fy = __fdiv_rn(__fadd_rn((float)py, jy), (float)cam.height);
	float sx = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(fx, 2.0f), 1.0f), cam.tan_half_fov), aspect);
	float sy = __fmul_rn(__fsub_rn(1.0f, __fmul_rn(fy, 2.0f)), cam.tan_half_fov);
	float d[3];
	for (int k = 0; k < 3; k++)
		d[k] = __fadd_rn(__fadd_rn(cam.forward[k], __fmul_rn(sx, cam.right[k])), __fmul_rn(sy, cam.up[k]));
	rays[2ull * i] = make_float4(cam.eye[0], cam.eye[1], cam.eye[2], d[0]);
	rays[2ull * i + 1] = make_float4(d[1], d[2], 0.0f, RTK_INF_F);
}

// bounce rays.  A path that missed either ends -- a dead ray (max_t = 0 can never be hit) and
// alive[i] = 0 -- or, with RTKD_BOUNCE_RELAUNCH, restarts from a uniformly chosen triangle
// (point ~ U(barycentric), normal turned towards +y; BASELINE.json config 5 keeps the ray count
// fixed that way) and alive[i] = 2.
#define RTKD_BOUNCE_RELAUNCH 1u

__global__ void k_gen_bounce(rtkd_arrays sc, const float4 *rays_in, const float4 *hit16, float4 *rays_out,
                             unsigned char *alive, uint32_t n, unsigned long long seed, uint32_t bounce,
                             unsigned long long first_ray, float push, uint32_t flags)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 h = hit16[i];
	uint32_t prim = __float_as_uint(h.w);
	const unsigned long long ctr = ((first_ray + i) * 16ull + bounce) * 8ull + 0x5bd1e995ull;
	float px, py, pz, nx, ny, nz;
	unsigned char state = 1;
	if (prim == RTK_MISS) {
		if (!(flags & RTKD_BOUNCE_RELAUNCH) || sc.num_tris == 0) {
			rays_out[2ull * i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			rays_out[2ull * i + 1] = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
			if (alive) alive[i] = 0;
			return;
		}
		state = 2;
		prim = rtk_umin((uint32_t)(rtk_u01(seed, ctr + 2) * (float)sc.num_tris), sc.num_tris - 1u);
	}
	const float4 *tv = sc.tri_orig + 3ull * prim;
	float4 a = __ldg(tv), b = __ldg(tv + 1), c = __ldg(tv + 2);
	float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z;
	float e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
	nx = e1y * e2z - e1z * e2y; ny = e1z * e2x - e1x * e2z; nz = e1x * e2y - e1y * e2x;
	float len = sqrtf(nx * nx + ny * ny + nz * nz);
	if (len > 0.0f) { nx /= len; ny /= len; nz /= len; } else { nx = 0.0f; ny = 1.0f; nz = 0.0f; }
	if (state == 1) {
		float4 r0 = rays_in[2ull * i], r1 = rays_in[2ull * i + 1];
		float t = h.x;
		px = r0.x + t * r0.w; py = r0.y + t * r1.x; pz = r0.z + t * r1.y;
		if (nx * r0.w + ny * r1.x + nz * r1.y > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }   // face the incoming ray
	} else {
		float sq = sqrtf(rtk_u01(seed, ctr + 3)), r2 = rtk_u01(seed, ctr + 4);
		float b0 = 1.0f - sq, b1 = sq * (1.0f - r2), b2 = 1.0f - b0 - b1;
		px = b0 * a.x + b1 * b.x + b2 * c.x; py = b0 * a.y + b1 * b.y + b2 * c.y; pz = b0 * a.z + b1 * b.z + b2 * c.z;
		if (ny < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
	}
	// orthonormal frame
	float ax = fabsf(nx) > 0.9f ? 0.0f : 1.0f, ay = fabsf(nx) > 0.9f ? 1.0f : 0.0f;
	float tx = ay * nz, ty = -ax * nz, tz = ax * ny - ay * nx;                  // cross((ax,ay,0), n)
	float tl = sqrtf(tx * tx + ty * ty + tz * tz);
	tx /= tl; ty /= tl; tz /= tl;
	float bx = ny * tz - nz * ty, by = nz * tx - nx * tz, bz = nx * ty - ny * tx;
	float u1 = rtk_u01(seed, ctr), u2 = rtk_u01(seed, ctr + 1);
	float r = sqrtf(u1), phi = 6.2831853f * u2;
	float lx = r * cosf(phi), ly = r * sinf(phi), lz = sqrtf(fmaxf(0.0f, 1.0f - u1));
	float ndx = lx * tx + ly * bx + lz * nx, ndy = lx * ty + ly * by + lz * ny, ndz = lx * tz + ly * bz + lz * nz;
	rays_out[2ull * i] = make_float4(px + push * nx, py + push * ny, pz + push * nz, ndx);
	rays_out[2ull * i + 1] = make_float4(ndy, ndz, 0.0f, RTK_INF_F);
	if (alive) alive[i] = state;
}
