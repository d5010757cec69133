// k_build.cuh -- BVH construction kernels (sm_100a).
//
// Replaces the reference's CPU build (rtk.c:1116-1182 triangle setup, :867-1019 binned-SAH
// recursion, :1570-1622 2->4 collapse) with a data-parallel pipeline:
//
//   k_decode_mesh     strided U16/U32/implicit indices + F32/F64 positions -> 48-byte corner
//                     records (rtk.c:1028-1114, 1150-1171)
//   k_scene_bounds    scene AABB, warp-shuffle + shared-memory reduction (rtk.c:1398-1404)
//   k_morton          63-bit Morton code of each triangle's AABB centre
//   k_radix_*         LSD radix sort, 8-bit digits, 64-bit keys + 32-bit payload
//   k_hierarchy       binary radix tree over the sorted codes (Karras 2012)
//   k_refit           bottom-up bounds with per-node arrival counters
//   k_collapse        binary tree -> 8-wide nodes + leaves of <= 8 triangles
//   k_emit_tris       leaf-ordered SoA triangle copy for traversal
//
// All kernels are memory-streaming integer/fp32 work; none is GEMM-shaped.
#pragma once
#include "rtk_common.cuh"

// ---------------------------------------------------------------------------------------------
// mesh decode
// ---------------------------------------------------------------------------------------------

struct rtkd_decode_args {
	const unsigned char *pos;     // device copy of the position buffer
	const unsigned char *idx;     // device copy of the index buffer or NULL
	unsigned long long pos_stride, idx_stride;
	int pos_f64;                  // positions are doubles (RTK_TYPE_F64), else floats
	int idx_bytes;                // 0 implicit, 2 uint16, 4 uint32
	int pregathered;              // positions are stored per corner (3 per triangle, in order)
	uint32_t ntris, first_prim;
	int has_xf;                   // instance transform: world = xf * (x, y, z, 1), row-major 3x4
	float xf[12];
};

RTK_DEV float4 rtk_fetch_corner(const rtkd_decode_args &a, unsigned long long slot, uint32_t index)
{
	const unsigned char *p = a.pos + slot * a.pos_stride;
	float x, y, z;
	if (a.pos_f64) {
		const double *d = (const double*)p;
		// the reference reads floats out of the f64 buffer here (defect D12, rtk.c:1098-1110);
		// the intended conversion is a round-to-nearest narrowing
		x = __double2float_rn(d[0]); y = __double2float_rn(d[1]); z = __double2float_rn(d[2]);
	} else {
		const float *f = (const float*)p;
		x = f[0]; y = f[1]; z = f[2];
	}
	if (a.has_xf) {
		// baked instance: ((m0*x + m1*y) + m2*z) + m3 per row, every operation rounded on its own (no
		// FMA contraction) so that a host restatement in plain fp32 reproduces the corners bit for bit
		const float wx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.xf[0], x), __fmul_rn(a.xf[1], y)), __fmul_rn(a.xf[2], z)), a.xf[3]);
		const float wy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.xf[4], x), __fmul_rn(a.xf[5], y)), __fmul_rn(a.xf[6], z)), a.xf[7]);
		const float wz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.xf[8], x), __fmul_rn(a.xf[9], y)), __fmul_rn(a.xf[10], z)), a.xf[11]);
		x = wx; y = wy; z = wz;
	}
	return make_float4(x, y, z, __uint_as_float(index));
}

__global__ void k_decode_mesh(rtkd_decode_args a, float4 *tri_orig)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.ntris) return;
	uint32_t i0, i1, i2;
	if (a.idx_bytes == 2) {
		const unsigned short *s = (const unsigned short*)(a.idx + (unsigned long long)i * a.idx_stride);
		i0 = s[0]; i1 = s[1]; i2 = s[2];
	} else if (a.idx_bytes == 4) {
		const uint32_t *s = (const uint32_t*)(a.idx + (unsigned long long)i * a.idx_stride);
		i0 = s[0]; i1 = s[1]; i2 = s[2];
	} else {
		i0 = 3u * i; i1 = i0 + 1u; i2 = i0 + 2u;          // rtk.c:1062-1068
	}
	unsigned long long s0 = a.pregathered ? 3ull * i : i0;
	unsigned long long s1 = a.pregathered ? 3ull * i + 1 : i1;
	unsigned long long s2 = a.pregathered ? 3ull * i + 2 : i2;
	float4 *dst = tri_orig + 3ull * (a.first_prim + i);
	dst[0] = rtk_fetch_corner(a, s0, i0);
	dst[1] = rtk_fetch_corner(a, s1, i1);
	dst[2] = rtk_fetch_corner(a, s2, i2);
}

// largest index of an index buffer (the reference never needs the vertex count; the upload does)
__global__ void k_max_index(const unsigned char *idx, unsigned long long stride, int idx_bytes, uint32_t ntris, uint32_t *out)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t m = 0;
	if (i < ntris) {
		if (idx_bytes == 2) {
			const unsigned short *s = (const unsigned short*)(idx + (unsigned long long)i * stride);
			m = rtk_umax(rtk_umax(s[0], s[1]), s[2]);
		} else {
			const uint32_t *s = (const uint32_t*)(idx + (unsigned long long)i * stride);
			m = rtk_umax(rtk_umax(s[0], s[1]), s[2]);
		}
	}
	for (int o = 16; o > 0; o >>= 1) m = rtk_umax(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// ---------------------------------------------------------------------------------------------
// scene bounds: bounds[0..2] = min (ordered-uint encoded), bounds[3..5] = max
// ---------------------------------------------------------------------------------------------

__global__ void k_scene_bounds(const float4 *tri_orig, uint32_t ntris, uint32_t *bounds)
{
	__shared__ float s_red[6][8];
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	float mn[3] = { +RTK_INF_F, +RTK_INF_F, +RTK_INF_F }, mx[3] = { -RTK_INF_F, -RTK_INF_F, -RTK_INF_F };
	if (i < ntris) {
		float4 a = tri_orig[3ull * i], b = tri_orig[3ull * i + 1], c = tri_orig[3ull * i + 2];
		mn[0] = rtk_fmin(rtk_fmin(a.x, b.x), c.x); mx[0] = rtk_fmax(rtk_fmax(a.x, b.x), c.x);
		mn[1] = rtk_fmin(rtk_fmin(a.y, b.y), c.y); mx[1] = rtk_fmax(rtk_fmax(a.y, b.y), c.y);
		mn[2] = rtk_fmin(rtk_fmin(a.z, b.z), c.z); mx[2] = rtk_fmax(rtk_fmax(a.z, b.z), c.z);
	}
	for (int k = 0; k < 3; k++)
		for (int o = 16; o > 0; o >>= 1) {
			mn[k] = rtk_fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
			mx[k] = rtk_fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
		}
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) for (int k = 0; k < 3; k++) { s_red[k][warp] = mn[k]; s_red[3 + k][warp] = mx[k]; }
	__syncthreads();
	if (threadIdx.x < 6) {
		int k = threadIdx.x;
		int nw = blockDim.x >> 5;
		float v = s_red[k][0];
		for (int w = 1; w < nw; w++) v = k < 3 ? rtk_fmin(v, s_red[k][w]) : rtk_fmax(v, s_red[k][w]);
		if (k < 3) atomicMin(&bounds[k], rtk_f2ord(v)); else atomicMax(&bounds[k], rtk_f2ord(v));
	}
}

// ---------------------------------------------------------------------------------------------
// Morton codes
// ---------------------------------------------------------------------------------------------

RTK_DEV unsigned long long rtk_expand21(unsigned long long x)
{
	x &= 0x1fffffull;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

__global__ void k_morton(const float4 *tri_orig, uint32_t ntris, const uint32_t *bounds,
                         unsigned long long *keys, uint32_t *vals)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= ntris) return;
	float4 a = tri_orig[3ull * i], b = tri_orig[3ull * i + 1], c = tri_orig[3ull * i + 2];
	float lo[3], ext[3], cen[3];
	for (int k = 0; k < 3; k++) {
		lo[k] = rtk_ord2f(bounds[k]);
		ext[k] = rtk_ord2f(bounds[3 + k]) - lo[k];
	}
	// AABB centre, the quantity the reference bins on (rtk.c:899)
	cen[0] = 0.5f * (rtk_fmin(rtk_fmin(a.x, b.x), c.x) + rtk_fmax(rtk_fmax(a.x, b.x), c.x));
	cen[1] = 0.5f * (rtk_fmin(rtk_fmin(a.y, b.y), c.y) + rtk_fmax(rtk_fmax(a.y, b.y), c.y));
	cen[2] = 0.5f * (rtk_fmin(rtk_fmin(a.z, b.z), c.z) + rtk_fmax(rtk_fmax(a.z, b.z), c.z));
	unsigned long long q[3];
	for (int k = 0; k < 3; k++) {
		float f = ext[k] > 0.0f ? (cen[k] - lo[k]) / ext[k] : 0.0f;
		f = rtk_fmin(rtk_fmax(f, 0.0f), 1.0f);
		// 21 bits per axis; double keeps all of them
		double s = (double)f * 2097151.0;
		q[k] = (unsigned long long)s;
	}
	keys[i] = (rtk_expand21(q[0]) << 2) | (rtk_expand21(q[1]) << 1) | rtk_expand21(q[2]);
	vals[i] = i;
}

// ---------------------------------------------------------------------------------------------
// LSD radix sort: 8-bit digits, 64-bit keys, 32-bit payload.
// One pass = k_radix_hist (per-block digit counts) + k_radix_scan (per-digit exclusive scan
// over blocks) + k_radix_scatter (stable ranking with warp match, then scatter).
// ---------------------------------------------------------------------------------------------

#define RTK_SORT_WARPS 8
#define RTK_SORT_THREADS (RTK_SORT_WARPS * 32)
#define RTK_SORT_ITEMS 16
#define RTK_SORT_TILE (RTK_SORT_THREADS * RTK_SORT_ITEMS)

__global__ void __launch_bounds__(RTK_SORT_THREADS) k_radix_hist(const unsigned long long *keys, uint32_t n, int shift,
                                                                uint32_t *counts, uint32_t nblocks)
{
	__shared__ uint32_t s_cnt[256];
	s_cnt[threadIdx.x] = 0;
	__syncthreads();
	uint32_t base = blockIdx.x * RTK_SORT_TILE;
	for (int r = 0; r < RTK_SORT_ITEMS; r++) {
		uint32_t i = base + r * RTK_SORT_THREADS + threadIdx.x;
		if (i < n) atomicAdd(&s_cnt[(uint32_t)(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// block d scans row d (nblocks entries) in place to exclusive offsets and writes the row total
__global__ void __launch_bounds__(256) k_radix_scan(uint32_t *counts, uint32_t nblocks, uint32_t *totals)
{
	__shared__ uint32_t s_warp[8];
	__shared__ uint32_t s_carry;
	uint32_t *row = counts + (size_t)blockIdx.x * nblocks;
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < nblocks; base += 256) {
		uint32_t i = base + threadIdx.x;
		uint32_t v = i < nblocks ? row[i] : 0;
		uint32_t x = v;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31) s_warp[warp] = x;
		__syncthreads();
		uint32_t woff = 0;
		for (int w = 0; w < warp; w++) woff += s_warp[w];
		uint32_t carry = s_carry;
		if (i < nblocks) row[i] = carry + woff + x - v;
		__syncthreads();
		if (threadIdx.x == 255) s_carry = carry + woff + x;
		__syncthreads();
	}
	if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(RTK_SORT_THREADS) k_radix_scatter(
	const unsigned long long *keys_in, const uint32_t *vals_in,
	unsigned long long *keys_out, uint32_t *vals_out, uint32_t n, int shift,
	const uint32_t *offsets, const uint32_t *totals, uint32_t nblocks)
{
	__shared__ uint32_t s_cnt[RTK_SORT_WARPS][256];
	__shared__ uint32_t s_base[256];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int w = 0; w < RTK_SORT_WARPS; w++) s_cnt[w][threadIdx.x] = 0;
	s_base[threadIdx.x] = totals[threadIdx.x];
	__syncthreads();
	// global start of each digit = exclusive scan of the 256 totals
	if (threadIdx.x == 0) {
		uint32_t run = 0;
		for (int d = 0; d < 256; d++) { uint32_t c = s_base[d]; s_base[d] = run; run += c; }
	}

	unsigned long long key[RTK_SORT_ITEMS];
	uint32_t rank[RTK_SORT_ITEMS];
	const uint32_t base = blockIdx.x * RTK_SORT_TILE + warp * (32 * RTK_SORT_ITEMS);
	const uint32_t lt = (1u << lane) - 1u;
	// phase 1: rank inside the warp, in key order (round r holds keys base + 32 r + lane)
#pragma unroll
	for (int r = 0; r < RTK_SORT_ITEMS; r++) {
		uint32_t i = base + r * 32 + lane;
		bool valid = i < n;
		key[r] = valid ? keys_in[i] : 0xffffffffffffffffull;
		uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & 255u) : 256u;
		uint32_t peers = __match_any_sync(0xffffffffu, d);
		uint32_t before = __popc(peers & lt);
		uint32_t old = 0;
		if (valid && before == 0) {
			old = s_cnt[warp][d];
			s_cnt[warp][d] = old + __popc(peers);
		}
		old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
		rank[r] = old + before;
		__syncwarp();
	}
	__syncthreads();
	// phase 2: per digit, exclusive scan over the warps, seeded with this block's global offset
	{
		uint32_t d = threadIdx.x;
		uint32_t run = s_base[d] + offsets[(size_t)d * nblocks + blockIdx.x];
		for (int w = 0; w < RTK_SORT_WARPS; w++) { uint32_t c = s_cnt[w][d]; s_cnt[w][d] = run; run += c; }
	}
	__syncthreads();
	// phase 3: scatter
#pragma unroll
	for (int r = 0; r < RTK_SORT_ITEMS; r++) {
		uint32_t i = base + r * 32 + lane;
		if (i < n) {
			uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
			uint32_t pos = s_cnt[warp][d] + rank[r];
			keys_out[pos] = key[r];
			vals_out[pos] = vals_in[i];
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Binary radix tree (Karras 2012).  Internal nodes 0..n-2, node 0 is the root.  A child id
// >= 0 is an internal node, < 0 is ~position of a sorted triangle.  Bounds are stored for
// internal node i at [i] and for sorted triangle j at [n-1+j].
// ---------------------------------------------------------------------------------------------

struct rtkd_bvh2 {
	int *left, *right;           // [n-1]
	int *parent;                 // [2n-1] parent internal node of node / leaf (n-1+j); root: -1
	int *first, *last;           // [n-1] covered range of sorted positions
	float4 *blo, *bhi;           // [2n-1]
	int *flags;                  // [n-1] arrival counters for the refit
};

RTK_DEV int rtk_delta(const unsigned long long *keys, int n, int i, int j)
{
	if (j < 0 || j >= n) return -1;
	unsigned long long a = keys[i], b = keys[j];
	if (a != b) return __clzll((long long)(a ^ b));
	return 64 + __clz(i ^ j);
}

__global__ void k_hierarchy(const unsigned long long *keys, int n, rtkd_bvh2 t)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
	int d = rtk_delta(keys, n, i, i + 1) - rtk_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
	int dmin = rtk_delta(keys, n, i, i - d);
	int lmax = 2;
	while (rtk_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
	int l = 0;
	for (int s = lmax >> 1; s >= 1; s >>= 1)
		if (rtk_delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
	int j = i + l * d;
	int dnode = rtk_delta(keys, n, i, j);
	int s = 0;
	int div = 2;
	for (;;) {
		int step = (l + div - 1) / div;
		if (rtk_delta(keys, n, i, i + (s + step) * d) > dnode) s += step;
		if (step <= 1) break;
		div <<= 1;
	}
	int gamma = i + s * d + rtk_imin(d, 0);
	int lo = rtk_imin(i, j), hi = rtk_imax(i, j);
	int lc = (lo == gamma) ? ~gamma : gamma;
	int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
	t.left[i] = lc; t.right[i] = rc;
	t.first[i] = lo; t.last[i] = hi;
	t.parent[lc >= 0 ? lc : (n - 1) + ~lc] = i;
	t.parent[rc >= 0 ? rc : (n - 1) + ~rc] = i;
	if (i == 0) t.parent[0] = -1;
}

__global__ void k_refit(const float4 *tri_orig, const uint32_t *vals, int n, rtkd_bvh2 t)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	uint32_t prim = vals[j];
	float4 a = tri_orig[3ull * prim], b = tri_orig[3ull * prim + 1], c = tri_orig[3ull * prim + 2];
	float4 lo = make_float4(rtk_fmin(rtk_fmin(a.x, b.x), c.x), rtk_fmin(rtk_fmin(a.y, b.y), c.y), rtk_fmin(rtk_fmin(a.z, b.z), c.z), 0.0f);
	float4 hi = make_float4(rtk_fmax(rtk_fmax(a.x, b.x), c.x), rtk_fmax(rtk_fmax(a.y, b.y), c.y), rtk_fmax(rtk_fmax(a.z, b.z), c.z), 0.0f);
	int self = (n - 1) + j;
	t.blo[self] = lo; t.bhi[self] = hi;
	int cur = t.parent[self];
	while (cur >= 0) {
		__threadfence();
		if (atomicAdd(&t.flags[cur], 1) == 0) return;      // first arrival: the sibling finishes
		__threadfence();
		int l = t.left[cur], r = t.right[cur];
		int li = l >= 0 ? l : (n - 1) + ~l, ri = r >= 0 ? r : (n - 1) + ~r;
		int other = (li == self) ? ri : li;
		float4 olo = __ldcg(&t.blo[other]), ohi = __ldcg(&t.bhi[other]);
		lo.x = rtk_fmin(lo.x, olo.x); lo.y = rtk_fmin(lo.y, olo.y); lo.z = rtk_fmin(lo.z, olo.z);
		hi.x = rtk_fmax(hi.x, ohi.x); hi.y = rtk_fmax(hi.y, ohi.y); hi.z = rtk_fmax(hi.z, ohi.z);
		t.blo[cur] = lo; t.bhi[cur] = hi;
		self = cur;
		cur = t.parent[cur];
	}
}

// ---------------------------------------------------------------------------------------------
// Collapse to 8-wide nodes.  One thread per wide node, one launch per level of wide nodes.
// A binary subtree covering <= RTK_LEAF_MAX triangles becomes a leaf (its triangles are
// contiguous in sorted order); otherwise the child with the largest surface area is opened
// until 8 slots are filled (the reference's 2->4 collapse, rtk.c:1570-1622, takes fixed
// grandchildren instead).
// ---------------------------------------------------------------------------------------------

RTK_DEV float rtk_half_area(float4 lo, float4 hi)
{
	float x = hi.x - lo.x, y = hi.y - lo.y, z = hi.z - lo.z;
	return x * y + y * z + z * x;
}

struct rtkd_collapse_args {
	const uint2 *work_in; const uint32_t *n_in;    // the level's item count lives on the device
	uint2 *work_out; uint32_t *n_out;
	uint32_t *node_alloc; uint32_t node_cap;
	uint32_t *leaf_count;        // also the leaf slot allocator
	uint2 *leaf_list;            // [slot] = (first sorted position, count): filled here, read by k_emit_leaves
	unsigned char *node_level;   // [wide node] = depth of the node (root 0): the refit walks levels bottom-up
	uint32_t level;              // depth of the nodes this launch fills
	double *sah_cost;            // accumulates area-weighted cost (divide by root area on host)
	float4 *nodes;
	int n;                       // triangles
	uint32_t *err;
	const int4 *rec;             // [binary node] (left, right, area of left, area of right) from k_collapse_prep
	const unsigned char *nleaf2; // [binary node] leaves below left | leaves below right << 4
};

#define RTK_BIDX(c, n) ((c) >= 0 ? (c) : ((n) - 1) + ~(c))

// number of leaves (maximal subtrees of at most RTK_LEAF_MAX triangles) below binary node c when c must be
// opened and that number is at most RTK_WIDE, else 0.  Only subtrees of at most RTK_WIDE * RTK_LEAF_MAX
// triangles can qualify, so the walk is short.
RTK_DEV int rtk_count_leaves(const rtkd_bvh2 &t, int c0)
{
	if (c0 < 0) return 0;
	const int span = t.last[c0] - t.first[c0] + 1;
	if (span <= RTK_LEAF_MAX || span > RTK_LEAF_MAX * RTK_WIDE) return 0;
	int todo[2 * RTK_WIDE + 2], ntodo = 0, leaves = 0;
	todo[ntodo++] = c0;
	while (ntodo > 0 && leaves + ntodo <= RTK_WIDE) {
		const int c1 = todo[--ntodo];
		if (c1 >= 0 && (t.last[c1] - t.first[c1] + 1) > RTK_LEAF_MAX) { todo[ntodo++] = t.left[c1]; todo[ntodo++] = t.right[c1]; }
		else leaves++;
	}
	return ntodo == 0 ? leaves : 0;
}

// What k_collapse wants to know when it opens binary node c, gathered into ONE 16-byte record (+ one byte) per node:
// both children, the half area of each child that can itself be opened (-1: a leaf) and the number of leaves below
// it.  k_collapse is one thread per wide node walking a chain of dependent loads; it used to take two rounds per
// opening (the children of c, then box / range / leaf count of each child from five arrays) and measured 17 us for
// the root alone, 276 us for the eight levels of a 1M-triangle tree.  This pass is one thread per binary node, all
// of them independent.  (Round 1 counted the leaves inside k_collapse with a local-memory stack per thread: that had
// doubled the kernel's time.)
__global__ void k_collapse_prep(rtkd_bvh2 t, const uint32_t *num_nodes_dev, uint32_t num_nodes_host, int n, int4 *rec, unsigned char *nleaf2)
{
	const uint32_t c0 = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t num = num_nodes_dev ? *num_nodes_dev : num_nodes_host;
	if (c0 >= num) return;
	if (c0 != 0 && t.last[c0] - t.first[c0] + 1 <= RTK_LEAF_MAX) return;   // becomes a leaf: never opened (the root always is)
	const int ch[2] = { t.left[c0], t.right[c0] };
	float area[2];
	int nl[2];
	for (int k = 0; k < 2; k++) {
		const int c = ch[k];
		const bool openable = c >= 0 && (t.last[c] - t.first[c] + 1) > RTK_LEAF_MAX;
		area[k] = openable ? rtk_half_area(t.blo[c], t.bhi[c]) : -1.0f;
		nl[k] = openable ? rtk_count_leaves(t, c) : 0;
	}
	rec[c0] = make_int4(ch[0], ch[1], __float_as_int(area[0]), __float_as_int(area[1]));
	nleaf2[c0] = (unsigned char)(nl[0] | (nl[1] << 4));
}

__global__ void k_collapse(rtkd_collapse_args a, rtkd_bvh2 t)
{
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = w < *a.n_in;          // lanes without a node stay for the warp-wide allocation below
	const int n = a.n;
	const int root = valid ? (int)a.work_in[w].x : 0;
	const uint32_t dst = valid ? a.work_in[w].y : 0u;
	int slot[RTK_WIDE];
	float area[RTK_WIDE];       // half area of a slot that can be opened, -1 for a leaf
	int nleaf[RTK_WIDE];        // leaves below an openable slot when that is at most RTK_WIDE, else 0
	int ns = valid ? 2 : 0;
	if (valid) {
		const int4 r = a.rec[root];
		const int q = a.nleaf2[root];
		slot[0] = r.x; slot[1] = r.y; area[0] = __int_as_float(r.z); area[1] = __int_as_float(r.w);
		nleaf[0] = q & 15; nleaf[1] = q >> 4;
	}
	while (valid && ns < RTK_WIDE) {
		// A subtree whose leaves ALL fit into the free slots is absorbed whole, largest area first: the
		// wide node it would have become -- with few children, near the bottom of the tree, where most
		// nodes are -- disappears (a third fewer wide nodes, 3 % fewer node visits per ray on the
		// terrain and 10 % on the soup, counted by the statistics kernel).  Otherwise the child with the
		// largest area is opened, as before.
		int best = -1; float ba = -1.0f;
		for (int k = 0; k < ns; k++) if (nleaf[k] && nleaf[k] <= RTK_WIDE - ns + 1 && area[k] > ba) { ba = area[k]; best = k; }
		if (best < 0) for (int k = 0; k < ns; k++) if (area[k] > ba) { ba = area[k]; best = k; }
		if (best < 0) break;
		const int c = slot[best];
		const int4 r = a.rec[c];                        // the one dependent load of an opening
		const int q = a.nleaf2[c];
		slot[best] = r.x; slot[ns] = r.y;
		area[best] = __int_as_float(r.z); area[ns] = __int_as_float(r.w);
		nleaf[best] = q & 15; nleaf[ns] = q >> 4;
		ns++;
	}
	// boxes and ranges of the slots: independent loads, issued before the allocation below so that they are in
	// flight while the warp waits for its atomics
	float4 lo[RTK_WIDE], hi[RTK_WIDE];
	uint32_t first[RTK_WIDE], count[RTK_WIDE];
#pragma unroll
	for (int k = 0; k < RTK_WIDE; k++) {
		lo[k] = make_float4(+RTK_INF_F, +RTK_INF_F, +RTK_INF_F, 0.0f);
		hi[k] = make_float4(-RTK_INF_F, -RTK_INF_F, -RTK_INF_F, 0.0f);
		first[k] = 0; count[k] = 0;
		if (k < ns) {
			const int c = slot[k];
			lo[k] = t.blo[RTK_BIDX(c, n)]; hi[k] = t.bhi[RTK_BIDX(c, n)];
			if (!(area[k] >= 0.0f)) {
				first[k] = c >= 0 ? (uint32_t)t.first[c] : (uint32_t)~c;
				count[k] = c >= 0 ? (uint32_t)(t.last[c] - t.first[c] + 1) : 1u;
			}
		}
	}
	// One allocation per kind for the whole WARP (it used to be two returning atomics per child, one after the
	// other; then one set per node -- still thousands of atomics on the same three addresses from the threads of
	// a level, which finish together): the children of a node get consecutive numbers, its leaves consecutive slots.
	uint32_t n_open = 0, n_leafs = 0;
	for (int k = 0; k < ns; k++) { if (area[k] >= 0.0f) n_open++; else n_leafs++; }
	__syncwarp();
	const int lane = threadIdx.x & 31;
	const uint32_t mine = n_open | (n_leafs << 16);          // at most 8 each: the warp's sums fit 16 bits
	uint32_t incl = mine;
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += y;
	}
	const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
	uint32_t base_idx = 0, base_o = 0, base_l = 0;
	if (lane == 31) {
		if (tot & 0xffffu) { base_idx = atomicAdd(a.node_alloc, tot & 0xffffu); base_o = atomicAdd(a.n_out, tot & 0xffffu); }
		if (tot >> 16) base_l = atomicAdd(a.leaf_count, tot >> 16);
	}
	base_idx = __shfl_sync(0xffffffffu, base_idx, 31);
	base_o = __shfl_sync(0xffffffffu, base_o, 31);
	base_l = __shfl_sync(0xffffffffu, base_l, 31);
	const uint32_t excl = incl - mine;
	uint32_t idx = base_idx + (excl & 0xffffu);
	uint32_t o = base_o + (excl & 0xffffu);
	uint32_t lslot = base_l + (excl >> 16);
	if (idx + n_open > a.node_cap) { atomicOr(a.err, 1u); idx = 0; }
	if (lslot + n_leafs > RTK_MAX_LEAVES) { atomicOr(a.err, 2u); lslot = 0; }
	float4 *node = a.nodes + 16ull * dst;
	double cost = 0.0;
#pragma unroll
	for (int k = 0; k < RTK_WIDE; k++) {
		if (!valid) break;
		uint32_t ref = RTK_REF_EMPTY;
		if (k < ns) {
			if (area[k] >= 0.0f) {
				a.work_out[o++] = make_uint2((uint32_t)slot[k], idx);
				a.node_level[idx] = (unsigned char)(a.level + 1u);
				ref = idx++;
			} else {
				// every leaf owns an 8-triangle slot of the traversal arrays: 128 aligned bytes per array
				a.leaf_list[lslot] = make_uint2(first[k], count[k]);
				ref = rtk_leaf_ref(lslot * RTK_LEAF_MAX, count[k]);
				lslot++;
			}
			cost += (double)rtk_half_area(lo[k], hi[k]);       // node step or one 8-lane triangle round: cost 1
		}
		node[2 * k] = make_float4(lo[k].x, lo[k].y, lo[k].z, __uint_as_float(ref));
		node[2 * k + 1] = make_float4(hi[k].x, hi[k].y, hi[k].z, 0.0f);
	}
	__syncwarp();
	for (int d = 16; d > 0; d >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, d);
	if (lane == 0 && cost != 0.0) atomicAdd(a.sah_cost, cost);
}

// scene with a single triangle: a root node with one leaf child
__global__ void k_single_root(const float4 *tri_orig, const uint32_t *vals, float4 *nodes)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	uint32_t prim = vals[0];
	float4 a = tri_orig[3ull * prim], b = tri_orig[3ull * prim + 1], c = tri_orig[3ull * prim + 2];
	for (int k = 0; k < RTK_WIDE; k++) {
		nodes[2 * k] = make_float4(+RTK_INF_F, +RTK_INF_F, +RTK_INF_F, __uint_as_float(RTK_REF_EMPTY));
		nodes[2 * k + 1] = make_float4(-RTK_INF_F, -RTK_INF_F, -RTK_INF_F, 0.0f);
	}
	nodes[0] = make_float4(rtk_fmin(rtk_fmin(a.x, b.x), c.x), rtk_fmin(rtk_fmin(a.y, b.y), c.y), rtk_fmin(rtk_fmin(a.z, b.z), c.z),
	                       __uint_as_float(rtk_leaf_ref(0, 1)));
	nodes[1] = make_float4(rtk_fmax(rtk_fmax(a.x, b.x), c.x), rtk_fmax(rtk_fmax(a.y, b.y), c.y), rtk_fmax(rtk_fmax(a.z, b.z), c.z), 0.0f);
}

// traversal triangles, SoA, one 8-entry slot per leaf (tv0[i].w = global triangle number): the
// triangles of a leaf are 128 contiguous, 128-byte aligned bytes in each of the three arrays, so
// the 8 lanes that test a leaf touch exactly one line per array.  Unused entries of a slot are
// never read by the traversal (the leaf reference carries the count); they are zero-filled with
// id RTK_MISS.
__global__ void k_emit_leaves(const float4 *tri_orig, const uint32_t *vals, const uint2 *leaf_list, uint32_t num_leaves,
                              float4 *tv0, float4 *tv1, float4 *tv2)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= num_leaves * RTK_LEAF_MAX) return;
	const uint2 lf = leaf_list[i / RTK_LEAF_MAX];
	const uint32_t j = i % RTK_LEAF_MAX;
	float4 a = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(RTK_MISS)), b = make_float4(0.0f, 0.0f, 0.0f, 0.0f), c = b;
	if (j < lf.y) {
		const uint32_t prim = vals[lf.x + j];
		a = tri_orig[3ull * prim]; b = tri_orig[3ull * prim + 1]; c = tri_orig[3ull * prim + 2];
		a.w = __uint_as_float(prim);
	}
	tv0[i] = a; tv1[i] = b; tv2[i] = c;
}

// ---------------------------------------------------------------------------------------------
// Refit: the vertices moved, the topology did not.  k_refit_tris reloads the
// traversal triangles from the decoded corners; k_refit_level then recomputes the child boxes of
// every wide node of one depth -- one thread per child slot -- from the leaf's triangles or from
// the 8 (already refitted) child boxes of the node below.  Levels run deepest first.
// ---------------------------------------------------------------------------------------------

__global__ void k_refit_tris(const float4 *tri_orig, uint32_t num_tv, float4 *tv0, float4 *tv1, float4 *tv2)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= num_tv) return;
	const uint32_t prim = __float_as_uint(tv0[i].w);
	if (prim == RTK_MISS) return;
	float4 a = tri_orig[3ull * prim], b = tri_orig[3ull * prim + 1], c = tri_orig[3ull * prim + 2];
	a.w = __uint_as_float(prim);
	tv0[i] = a; tv1[i] = b; tv2[i] = c;
}

__global__ void k_refit_level(float4 *nodes, const unsigned char *node_level, uint32_t num_nodes, uint32_t level,
                              const float4 *tv0, const float4 *tv1, const float4 *tv2)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t node = t / RTK_WIDE, k = t % RTK_WIDE;
	if (node >= num_nodes || node_level[node] != level) return;
	float4 *slot = nodes + 16ull * node + 2 * k;
	const uint32_t ref = __float_as_uint(slot[0].w);
	if (ref == RTK_REF_EMPTY) return;
	float lo[3] = { +RTK_INF_F, +RTK_INF_F, +RTK_INF_F }, hi[3] = { -RTK_INF_F, -RTK_INF_F, -RTK_INF_F };
	if (rtk_ref_is_leaf(ref)) {
		const uint32_t first = rtk_leaf_first(ref), cnt = rtk_leaf_count(ref);
		for (uint32_t j = 0; j < cnt; j++) {
			const float4 a = tv0[first + j], b = tv1[first + j], c = tv2[first + j];
			lo[0] = rtk_fmin(lo[0], rtk_fmin(rtk_fmin(a.x, b.x), c.x)); hi[0] = rtk_fmax(hi[0], rtk_fmax(rtk_fmax(a.x, b.x), c.x));
			lo[1] = rtk_fmin(lo[1], rtk_fmin(rtk_fmin(a.y, b.y), c.y)); hi[1] = rtk_fmax(hi[1], rtk_fmax(rtk_fmax(a.y, b.y), c.y));
			lo[2] = rtk_fmin(lo[2], rtk_fmin(rtk_fmin(a.z, b.z), c.z)); hi[2] = rtk_fmax(hi[2], rtk_fmax(rtk_fmax(a.z, b.z), c.z));
		}
	} else {
		const float4 *ch = nodes + 16ull * ref;
		for (int j = 0; j < RTK_WIDE; j++) {
			const float4 l = ch[2 * j], h = ch[2 * j + 1];
			if (__float_as_uint(l.w) == RTK_REF_EMPTY) continue;
			lo[0] = rtk_fmin(lo[0], l.x); lo[1] = rtk_fmin(lo[1], l.y); lo[2] = rtk_fmin(lo[2], l.z);
			hi[0] = rtk_fmax(hi[0], h.x); hi[1] = rtk_fmax(hi[1], h.y); hi[2] = rtk_fmax(hi[2], h.z);
		}
	}
	slot[0] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(ref));
	slot[1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
}

// ---------------------------------------------------------------------------------------------
// Triangle filter: a device-side predicate as a bitset over global triangle
// numbers.  The predicate is baked into the leaf slots instead of being evaluated per ray: corner 0
// of a triangle that is switched off becomes NaN, which the watertight test can never accept
// (v, w, det and t are NaN; rtk_tri_test's `t > min_t` is false), so the traversal kernels need no
// extra load, register or branch.  Corner 0 of every other triangle is restored from the decoded
// corners, which makes the pass idempotent and lets bits == NULL remove the filter.  Boxes are not
// touched (they stay conservative); a refit rewrites the slots first and re-applies the filter last.
// ---------------------------------------------------------------------------------------------

__global__ void k_apply_filter(const float4 *tri_orig, const uint32_t *bits, uint32_t num_tv, float4 *tv0)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= num_tv) return;
	const uint32_t prim = __float_as_uint(tv0[i].w);
	if (prim == RTK_MISS) return;
	float4 a = tri_orig[3ull * prim];
	if (bits && !((bits[prim >> 5] >> (prim & 31u)) & 1u)) {
		const float qnan = __uint_as_float(0x7fc00000u);
		a.x = qnan; a.y = qnan; a.z = qnan;
	}
	a.w = __uint_as_float(prim);
	tv0[i] = a;
}
