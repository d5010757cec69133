// rtk_common.cuh -- shared definitions of the device code: launch macro, device data layout,
// small helpers.  Compiled by nvcc for sm_100a; the test suite additionally compiles the very
// same sources with g++ against tests/emu/simt.h (RTK_SIMT_EMU) to execute them on the CPU.
#pragma once

#ifndef RTK_SIMT_EMU
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define RTK_LAUNCH(kernel, grid, block, stream, ...) \
	kernel<<<dim3(grid), dim3(block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__)
// launch with an L2 access-policy window: [win_base, win_base + win_bytes) is fetched as PERSISTING
// (a fraction win_ratio of it when it exceeds the set-aside part of the L2)
#define RTK_LAUNCH_WIN(kernel, grid, block, strm_, win_base, win_bytes, win_ratio, ...) do { \
	cudaLaunchConfig_t _cfg; memset(&_cfg, 0, sizeof(_cfg)); \
	_cfg.gridDim = dim3(grid); _cfg.blockDim = dim3(block); _cfg.stream = (cudaStream_t)(strm_); \
	cudaLaunchAttribute _at[1]; memset(_at, 0, sizeof(_at)); \
	if ((win_base) && (win_bytes)) { \
		_at[0].id = cudaLaunchAttributeAccessPolicyWindow; \
		_at[0].val.accessPolicyWindow.base_ptr = (void*)(win_base); \
		_at[0].val.accessPolicyWindow.num_bytes = (size_t)(win_bytes); \
		_at[0].val.accessPolicyWindow.hitRatio = (win_ratio); \
		_at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; \
		_at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal; \
		_cfg.attrs = _at; _cfg.numAttrs = 1; \
	} \
	cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__); \
} while (0)
#else
#define RTK_LAUNCH_WIN(kernel, grid, block, stream, win_base, win_bytes, win_ratio, ...) \
	RTK_LAUNCH(kernel, grid, block, stream, __VA_ARGS__)
#endif

#define RTK_DEV __device__ __forceinline__

// NVTX ranges around the host-side stages (build phases, the chunks of a host batch, replication): they
// show up in Nsight timelines and cost nothing when no tool is attached
#ifndef RTK_SIMT_EMU
#include <nvtx3/nvToolsExt.h>
struct rtk_nvtx_range { rtk_nvtx_range(const char *n) { nvtxRangePushA(n); } ~rtk_nvtx_range() { nvtxRangePop(); } };
#else
struct rtk_nvtx_range { rtk_nvtx_range(const char *) {} };
#endif
#define RTK_NVTX_CAT2(a, b) a##b
#define RTK_NVTX_CAT(a, b) RTK_NVTX_CAT2(a, b)
#define RTK_NVTX(name) rtk_nvtx_range RTK_NVTX_CAT(_nvtx_, __LINE__)(name)

#define RTK_MISS 0xffffffffu
#define RTK_INF_F 3.402823e+38f          // RTK_INF, reference rtk.h:11

// ---------------------------------------------------------------------------------------------
// Device scene layout (all arrays in HBM, 256-byte aligned by cudaMalloc)
//
//   tri_orig   [3*N] float4  original-order triangles, one rtk_vertex (xyz + mesh vertex index)
//                            per corner: exactly the 48 bytes rtk_hit::vertex[3] wants
//                            (reference rtk.c:1162-1167, 376-378).  Read by the resolve kernel.
//   tv0,tv1,tv2 [8*leaves] float4  SoA copy for traversal, one 8-entry slot per leaf: corner k of
//                            the j-th triangle of leaf L is tvk[8L+j]; tv0[i].w carries the global
//                            triangle number (RTK_MISS in the unused tail of a slot).  The 8 lanes
//                            that test a leaf fetch one aligned 128-byte line per array.
//   nodes      [16*M] float4 8-wide nodes, 256 bytes each: child slot k is 32 bytes,
//                              (lo.x, lo.y, lo.z, ref) (hi.x, hi.y, hi.z, unused)
//                            ref: 0xffffffff empty | bit31 set: leaf, bits[30:3] first triangle
//                            (leaf order), bits[2:0] count-1 | else index of an 8-wide node.
//                            A lane fetches a child with ONE 256-bit load (LDG.E.256, new on
//                            sm_100); with L lanes per ray, lane c owns slots c, c+L, ... so the
//                            j-th load of a ray's lanes covers L*32 contiguous bytes.
//   mesh_first [num_meshes+1] first global triangle number of each mesh.
// ---------------------------------------------------------------------------------------------

#define RTK_WIDE 8                       // children per node == lanes per ray
#ifndef RTK_LEAF_MAX
#define RTK_LEAF_MAX 8                   // triangles per leaf (<= 8: the leaf reference keeps count-1 in 3 bits)
#endif
#define RTK_MAX_LEAVES (1u << 25)         // 8 * leaves must fit the 28-bit "first" field of a leaf reference
#define RTK_REF_EMPTY 0xffffffffu
#define RTK_REF_LEAF 0x80000000u

struct rtkd_arrays {
	const float4 *tri_orig;
	const float4 *tv0, *tv1, *tv2;
	const float4 *nodes;
	const uint32_t *mesh_first;
	uint32_t num_tris, num_meshes, num_nodes;
	uint32_t num_tv;                     // entries of tv0/tv1/tv2 = 8 * leaves
	float abs_max;                       // largest |coordinate| of the scene bounds
};

RTK_DEV uint32_t rtk_leaf_ref(uint32_t first, uint32_t count) { return RTK_REF_LEAF | (first << 3) | (count - 1u); }
RTK_DEV bool     rtk_ref_is_leaf(uint32_t ref) { return (ref & RTK_REF_LEAF) != 0; }
RTK_DEV uint32_t rtk_leaf_first(uint32_t ref) { return (ref & 0x7fffffffu) >> 3; }
RTK_DEV uint32_t rtk_leaf_count(uint32_t ref) { return (ref & 7u) + 1u; }

RTK_DEV float rtk_fmin(float a, float b) { return fminf(a, b); }
RTK_DEV float rtk_fmax(float a, float b) { return fmaxf(a, b); }
RTK_DEV uint32_t rtk_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
RTK_DEV uint32_t rtk_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
RTK_DEV int rtk_imin(int a, int b) { return a < b ? a : b; }
RTK_DEV int rtk_imax(int a, int b) { return a > b ? a : b; }

// order-preserving float <-> uint32 map (for atomicMin/Max on floats)
RTK_DEV uint32_t rtk_f2ord(float f)
{
	uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
RTK_DEV float rtk_ord2f(uint32_t u)
{
	return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// 32-byte read-only global load: one LDG.E.256 on sm_100a
RTK_DEV void rtk_ldg256(const float4 *p, float4 &a, float4 &b)
{
#ifdef RTK_SIMT_EMU
	a = p[0]; b = p[1];
#else
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#endif
}

// asynchronous 16-byte global->shared copy (LDGSTS); the emulator copies synchronously
RTK_DEV void rtk_cp_async16(void *smem_dst, const void *gmem_src)
{
#ifdef RTK_SIMT_EMU
	memcpy(smem_dst, gmem_src, 16);
#else
	unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem_src) : "memory");
#endif
}
RTK_DEV void rtk_cp_async_commit()
{
#ifndef RTK_SIMT_EMU
	asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
RTK_DEV void rtk_cp_async_wait_all()
{
#ifndef RTK_SIMT_EMU
	asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}
