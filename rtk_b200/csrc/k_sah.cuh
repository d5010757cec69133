// k_sah.cuh -- binned-SAH BVH builder (sm_100a), the GPU re-design of the reference's top-down
// build (rtk.c:867-1019 _rtk_build_node_sah, :1421-1453 dispatcher, :813-865 equal split).
//
// Same algorithm: 32 bins on each of the 3 axes, triangles binned by the centre of their AABB
// relative to the node's AABB (rtk.c:892-902), suffix/prefix sweep for the cheapest split
// (rtk.c:910-945), partition by bin index (rtk.c:968-986), depth cap with forced equal splits
// (rtk.c:1429-1443).  The cost counts SIMD batches of the leaf test, ceil(n/8) for the 8-lane
// ray groups of k_trace where the reference uses ceil(n/4) for SSE (rtk.c:934-935); a node with
// more than RTK_LEAF_MAX triangles is always split and one with at most RTK_LEAF_MAX never is
// (one leaf visit tests up to 8 triangles at once), which is what rtk.c:948-949 reduces to when
// max_leaf_items equals the SIMD width.
//
// Parallel structure:
//   * triangles start in Morton order (k_morton + radix sort), their AABBs are stored in that
//     order, so the members of any subtree are close in memory;
//   * LARGE nodes (> RTK_SAH_SMALL triangles) are processed level by level: k_sah_bin_large
//     reduces 2048-triangle chunks into shared-memory bins and merges them into the node's global
//     bins with atomics, k_sah_split_large evaluates the sweep with one warp per node (warp
//     prefix scans), k_sah_partition_large scatters triangle indices with block-aggregated
//     cursors;
//   * every SMALL node (<= RTK_SAH_SMALL) is finished by ONE CTA entirely in shared memory
//     (k_sah_small): AABBs and the index permutation stay on chip for all remaining levels.
#pragma once
#include "rtk_common.cuh"
#include "k_build.cuh"

#define RTK_SAH_BINS 32
#define RTK_SAH_SMALL 512
#define RTK_SAH_CHUNK 2048
#ifndef RTK_SAH_COMPACT_MIN
#define RTK_SAH_COMPACT_MIN 32768    // large nodes of at least this many triangles are binned a warp per 32 consecutive positions
#endif
#define RTK_SAH_MAX_DEPTH 64          // RTK_BVH_MAX_DEPTH, rtk.c:5
#define RTK_SAH_BINWORDS 8            // lo xyz, hi xyz, count, pad
#define RTK_SAH_NODEBINS (3 * RTK_SAH_BINS * RTK_SAH_BINWORDS)
// word w of bin b of axis a.  Word-major: the 32 bins of one word are 32 consecutive words, so the lanes of a warp
// that update different bins hit different shared-memory banks (bin-major, 8 words per bin, put every bin on one of
// 4 banks: an 8-way conflict on each of the 21 atomics of a triangle), and the sweep's lane-per-bin reads coalesce
#define RTK_SAH_BIN_AT(a, b, w) ((((a) * RTK_SAH_BINWORDS + (w)) * RTK_SAH_BINS) + (b))

struct rtkd_sah {
	const float4 *pb;             // [2n] AABBs in Morton order: pb[2j] = lo, pb[2j+1] = hi
	uint32_t *idx0, *idx1;        // ping-pong permutation of Morton positions
	uint32_t *idx_final;          // final leaf order (Morton positions)
	int *left, *right, *first, *last;
	float4 *blo, *bhi;
	uint32_t *ndepth;
	uint32_t *counters;           // [0] node_alloc [2] n_small [3] err; [8 + 2 * (level & 1)] large nodes, [9 + 2 * (level & 1)] chunks of a level
	uint32_t *act_in, *act_out;   // large nodes of this / the next level
	uint32_t *small_list;         // node id | (buffer << 31)
	uint32_t *chunk_base;         // [n_act] first chunk of each active node of this level
	uint32_t *chunk_node;         // [chunks] the active node each chunk of this level belongs to
	uint32_t *chunk_base_out, *chunk_node_out;   // the same for the next level (filled by k_sah_split_large)
	uint32_t chunk_cap;           // capacity of chunk_node
	uint32_t *bins;               // [n_act][3][8][32]
	uint32_t *binpack;            // [n] the three bins of the triangle at each position of the level's permutation (large levels)
	int4 *split;                  // per active node: axis (-1: equal split), bin, n_left, first child
	uint32_t *cursor;             // per active node: left / right write cursors
	uint32_t node_cap;
	uint32_t act_cap, small_cap;  // capacities of act_in / act_out and small_list (never reached: see carve(); guarded all the same)
};

#define RTK_SAH_LEVEL_NODES(s, par) ((s).counters[8 + 2 * (par)])
#define RTK_SAH_LEVEL_CHUNKS(s, par) ((s).counters[9 + 2 * (par)])

// bin of a triangle on one axis, rtk.c:892-902
RTK_DEV int rtk_sah_bin(float lo, float hi, float nmin, float nmax)
{
	float min_2x = nmin + nmin;
	float rcp_scale_2x = (0.5f * (float)RTK_SAH_BINS) / (nmax - nmin);
	float f = ((lo + hi) - min_2x) * rcp_scale_2x;
	if (!(f >= 0.0f)) return 0;
	if (f >= (float)RTK_SAH_BINS) return RTK_SAH_BINS - 1;
	return (int)f;
}

RTK_DEV void rtk_sah_bins_clear(uint32_t *bins, int tid, int nthreads)
{
	const uint32_t pinf = rtk_f2ord(+RTK_INF_F), ninf = rtk_f2ord(-RTK_INF_F);
	for (int i = tid; i < RTK_SAH_NODEBINS; i += nthreads) {
		int w = (i / RTK_SAH_BINS) & 7;
		bins[i] = w < 3 ? pinf : (w < 6 ? ninf : 0u);
	}
}

// add one AABB to the three axes' bins (shared or global memory)
RTK_DEV uint32_t rtk_sah_bin_add(uint32_t *bins, float4 lo, float4 hi, float4 nlo, float4 nhi)
{
	int b[3];
	b[0] = rtk_sah_bin(lo.x, hi.x, nlo.x, nhi.x);
	b[1] = rtk_sah_bin(lo.y, hi.y, nlo.y, nhi.y);
	b[2] = rtk_sah_bin(lo.z, hi.z, nlo.z, nhi.z);
	uint32_t ol[3] = { rtk_f2ord(lo.x), rtk_f2ord(lo.y), rtk_f2ord(lo.z) };
	uint32_t oh[3] = { rtk_f2ord(hi.x), rtk_f2ord(hi.y), rtk_f2ord(hi.z) };
	for (int a = 0; a < 3; a++) {
		uint32_t *p = bins + RTK_SAH_BIN_AT(a, b[a], 0);
		atomicMin(p + 0 * RTK_SAH_BINS, ol[0]); atomicMin(p + 1 * RTK_SAH_BINS, ol[1]); atomicMin(p + 2 * RTK_SAH_BINS, ol[2]);
		atomicMax(p + 3 * RTK_SAH_BINS, oh[0]); atomicMax(p + 4 * RTK_SAH_BINS, oh[1]); atomicMax(p + 5 * RTK_SAH_BINS, oh[2]);
		atomicAdd(p + 6 * RTK_SAH_BINS, 1u);
	}
	return (uint32_t)(b[0] | (b[1] << 8) | (b[2] << 16));
}

#ifndef RTK_SAH_REDUX_MIN
#define RTK_SAH_REDUX_MIN 6          // lanes that must share a bin before the warp reduces them with REDUX
#endif
// Per axis, the lanes of a warp that fall into the same bin are reduced with the hardware warp reductions (REDUX)
// and ONE lane issues the 7 atomics, as long as such a group has at least RTK_SAH_REDUX_MIN lanes; smaller groups
// update the bins lane by lane.  Morton-neighbours share bins: on the upper levels a warp is one group (round 2:
// 2.13 -> 2.07 ms at 1M triangles against the shuffle tree of round 1) or two or three -- a warp that straddles a
// bin boundary used to fall back to 32 lanes fighting over two addresses, which made levels 1-4 the slowest.  A
// version that walked ALL distinct bins with REDUX was 3x slower on the deep levels, where a warp's triangles
// spread over many bins: the walk stops after two small groups.
RTK_DEV uint32_t rtk_sah_bin_add_warp(uint32_t *bins, float4 lo, float4 hi, float4 nlo, float4 nhi, bool valid)
{
	const uint32_t FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const uint32_t vm = __ballot_sync(FULL, valid);
	if (vm == 0) return 0u;
	const int b[3] = { rtk_sah_bin(lo.x, hi.x, nlo.x, nhi.x), rtk_sah_bin(lo.y, hi.y, nlo.y, nhi.y), rtk_sah_bin(lo.z, hi.z, nlo.z, nhi.z) };
	const uint32_t ol[3] = { rtk_f2ord(lo.x), rtk_f2ord(lo.y), rtk_f2ord(lo.z) };
	const uint32_t oh[3] = { rtk_f2ord(hi.x), rtk_f2ord(hi.y), rtk_f2ord(hi.z) };
#pragma unroll
	for (int a = 0; a < 3; a++) {
		uint32_t rem = vm, solo = 0;
		int small = 0;
		while (rem && small < 2) {                      // warp-uniform
			const int leader = __ffs((int)rem) - 1;
			const int lb = __shfl_sync(FULL, b[a], leader);
			const uint32_t grp = __ballot_sync(FULL, valid && b[a] == lb) & rem;
			rem &= ~grp;
			if (__popc(grp) < RTK_SAH_REDUX_MIN) { solo |= grp; small++; continue; }
			const bool in = (grp >> lane) & 1u;
			const uint32_t m0 = __reduce_min_sync(FULL, in ? ol[0] : 0xffffffffu);
			const uint32_t m1 = __reduce_min_sync(FULL, in ? ol[1] : 0xffffffffu);
			const uint32_t m2 = __reduce_min_sync(FULL, in ? ol[2] : 0xffffffffu);
			const uint32_t x0 = __reduce_max_sync(FULL, in ? oh[0] : 0u);
			const uint32_t x1 = __reduce_max_sync(FULL, in ? oh[1] : 0u);
			const uint32_t x2 = __reduce_max_sync(FULL, in ? oh[2] : 0u);
			if (lane == leader) {
				uint32_t *p = bins + RTK_SAH_BIN_AT(a, lb, 0);
				atomicMin(p + 0 * RTK_SAH_BINS, m0); atomicMin(p + 1 * RTK_SAH_BINS, m1); atomicMin(p + 2 * RTK_SAH_BINS, m2);
				atomicMax(p + 3 * RTK_SAH_BINS, x0); atomicMax(p + 4 * RTK_SAH_BINS, x1); atomicMax(p + 5 * RTK_SAH_BINS, x2);
				atomicAdd(p + 6 * RTK_SAH_BINS, (uint32_t)__popc(grp));
			}
		}
		if (((solo | rem) >> lane) & 1u) {
			uint32_t *p = bins + RTK_SAH_BIN_AT(a, b[a], 0);
			atomicMin(p + 0 * RTK_SAH_BINS, ol[0]); atomicMin(p + 1 * RTK_SAH_BINS, ol[1]); atomicMin(p + 2 * RTK_SAH_BINS, ol[2]);
			atomicMax(p + 3 * RTK_SAH_BINS, oh[0]); atomicMax(p + 4 * RTK_SAH_BINS, oh[1]); atomicMax(p + 5 * RTK_SAH_BINS, oh[2]);
			atomicAdd(p + 6 * RTK_SAH_BINS, 1u);
		}
	}
	return (uint32_t)(b[0] | (b[1] << 8) | (b[2] << 16));
}

struct rtk_sah_choice {
	int axis, bin;               // axis < 0: no valid split
	uint32_t n_left;
	float llo[3], lhi[3], rlo[3], rhi[3];
	float cost;                  // only meaningful between rtk_sah_sweep_axis and rtk_sah_pick_axis
};

// One axis of the sweep of rtk.c:909-945 with one warp: lane i owns bin i and prices the split after it.
// Returns the cost (RTK_INF_F: not a valid split), the triangles on the left and the two boxes.
RTK_DEV float rtk_sah_axis_candidate(const uint32_t *bins, int axis, uint32_t count, float rcp_parent, uint32_t &nl_out, float (&L)[6], float (&R)[6])
{
	const uint32_t FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const uint32_t *p = bins + RTK_SAH_BIN_AT(axis, lane, 0);
	float lo[3] = { rtk_ord2f(p[0 * RTK_SAH_BINS]), rtk_ord2f(p[1 * RTK_SAH_BINS]), rtk_ord2f(p[2 * RTK_SAH_BINS]) };
	float hi[3] = { rtk_ord2f(p[3 * RTK_SAH_BINS]), rtk_ord2f(p[4 * RTK_SAH_BINS]), rtk_ord2f(p[5 * RTK_SAH_BINS]) };
	uint32_t cnt = p[6 * RTK_SAH_BINS];
	float Llo[3] = { lo[0], lo[1], lo[2] }, Lhi[3] = { hi[0], hi[1], hi[2] };
	float Rlo[3] = { lo[0], lo[1], lo[2] }, Rhi[3] = { hi[0], hi[1], hi[2] };
	uint32_t nl = cnt;
	for (int o = 1; o < 32; o <<= 1) {
		for (int k = 0; k < 3; k++) {
			float a = __shfl_up_sync(FULL, Llo[k], o), b = __shfl_up_sync(FULL, Lhi[k], o);
			float c = __shfl_down_sync(FULL, Rlo[k], o), d = __shfl_down_sync(FULL, Rhi[k], o);
			if (lane >= o) { Llo[k] = rtk_fmin(Llo[k], a); Lhi[k] = rtk_fmax(Lhi[k], b); }
			if (lane + o < 32) { Rlo[k] = rtk_fmin(Rlo[k], c); Rhi[k] = rtk_fmax(Rhi[k], d); }
		}
		uint32_t e = __shfl_up_sync(FULL, nl, o);
		if (lane >= o) nl += e;
	}
	// split after bin `lane`: left = bins 0..lane (mine), right = bins lane+1..31 (neighbour's suffix)
	for (int k = 0; k < 3; k++) {
		L[k] = Llo[k]; L[3 + k] = Lhi[k];
		R[k] = __shfl_down_sync(FULL, Rlo[k], 1);
		R[3 + k] = __shfl_down_sync(FULL, Rhi[k], 1);
	}
	nl_out = nl;
	const uint32_t nr = count - nl;
	if (!(lane < RTK_SAH_BINS - 1 && nl > 0 && nr > 0)) return RTK_INF_F;
	float lx = L[3] - L[0], ly = L[4] - L[1], lz = L[5] - L[2];
	float rx = R[3] - R[0], ry = R[4] - R[1], rz = R[5] - R[2];
	float area_l = 2.0f * (lx * ly + ly * lz + lz * lx);                // rtk.c:729-733
	float area_r = 2.0f * (rx * ry + ry * rz + rz * rx);
	float cost_l = (float)((nl + RTK_LEAF_MAX - 1u) / RTK_LEAF_MAX), cost_r = (float)((nr + RTK_LEAF_MAX - 1u) / RTK_LEAF_MAX);  // rtk.c:934-935
	return 1.0f + (area_l * cost_l + area_r * cost_r) * rcp_parent;     // rtk.c:936, split cost 1
}

RTK_DEV float rtk_sah_rcp_parent(float4 plo, float4 phi)
{
	float px = phi.x - plo.x, py = phi.y - plo.y, pz = phi.z - plo.z;
	return 1.0f / (2.0f * (px * py + py * pz + pz * px));          // rtk.c:880
}

RTK_DEV void rtk_sah_choice_none(rtk_sah_choice &c, float4 plo, float4 phi, uint32_t count)
{
	c.axis = -1; c.bin = 0; c.n_left = count / 2; c.cost = RTK_INF_F;
	for (int k = 0; k < 3; k++) { c.llo[k] = c.rlo[k] = k == 0 ? plo.x : (k == 1 ? plo.y : plo.z); c.lhi[k] = c.rhi[k] = k == 0 ? phi.x : (k == 1 ? phi.y : phi.z); }
}

// The whole sweep with one warp.  Returns the same choice in every lane.  Ties go to the lower axis, then the
// lower bin (the reference's strict '<' in axis-major, bin-minor order, rtk.c:938).
RTK_DEV rtk_sah_choice rtk_sah_sweep_warp(const uint32_t *bins, float4 plo, float4 phi, uint32_t count)
{
	const uint32_t FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const float rcp_parent = rtk_sah_rcp_parent(plo, phi);
	float best_cost = RTK_INF_F;
	int best_axis = -1;
	uint32_t best_nl = 0;
	float bl[6] = { 0, 0, 0, 0, 0, 0 }, br[6] = { 0, 0, 0, 0, 0, 0 };
	for (int axis = 0; axis < 3; axis++) {
		float L[6], R[6];
		uint32_t nl;
		const float cost = rtk_sah_axis_candidate(bins, axis, count, rcp_parent, nl, L, R);
		if (cost < best_cost) {
			best_cost = cost; best_axis = axis; best_nl = nl;
			for (int k = 0; k < 6; k++) { bl[k] = L[k]; br[k] = R[k]; }
		}
	}
	// warp argmin over (cost, axis, bin)
	float m = best_cost;
	for (int o = 16; o > 0; o >>= 1) m = rtk_fmin(m, __shfl_xor_sync(FULL, m, o));
	uint32_t key = (best_axis >= 0 && best_cost == m) ? (uint32_t)(best_axis * 32 + lane) : 0xffffffffu;
	for (int o = 16; o > 0; o >>= 1) key = rtk_umin(key, __shfl_xor_sync(FULL, key, o));
	rtk_sah_choice c;
	if (key == 0xffffffffu) { rtk_sah_choice_none(c, plo, phi, count); return c; }
	int src = (int)(key & 31u);
	c.axis = (int)(key >> 5); c.bin = src; c.cost = m;
	// the winning lane may hold a different axis as its own best: only `src` is read
	c.n_left = __shfl_sync(FULL, best_nl, src);
	for (int k = 0; k < 3; k++) {
		c.llo[k] = __shfl_sync(FULL, bl[k], src); c.lhi[k] = __shfl_sync(FULL, bl[3 + k], src);
		c.rlo[k] = __shfl_sync(FULL, br[k], src); c.rhi[k] = __shfl_sync(FULL, br[3 + k], src);
	}
	return c;
}

// One axis of the sweep with one warp, for callers that give each axis its own warp (the three scans are the
// longest dependency chain of a split).  The best split of the axis -- lowest cost, then lowest bin -- in every lane;
// axis < 0 when the axis has no valid split.  rtk_sah_pick_axis then takes the lowest cost, lower axis on ties:
// the same choice as rtk_sah_sweep_warp.
RTK_DEV rtk_sah_choice rtk_sah_sweep_axis(const uint32_t *bins, int axis, float4 plo, float4 phi, uint32_t count)
{
	const uint32_t FULL = 0xffffffffu;
	float L[6], R[6];
	uint32_t nl;
	const float cost = rtk_sah_axis_candidate(bins, axis, count, rtk_sah_rcp_parent(plo, phi), nl, L, R);
	float m = cost;
	for (int o = 16; o > 0; o >>= 1) m = rtk_fmin(m, __shfl_xor_sync(FULL, m, o));
	const uint32_t hit = __ballot_sync(FULL, cost < RTK_INF_F && cost == m);
	rtk_sah_choice c;
	if (!hit) { rtk_sah_choice_none(c, plo, phi, count); return c; }
	const int src = __ffs((int)hit) - 1;
	c.axis = axis; c.bin = src; c.cost = m;
	c.n_left = __shfl_sync(FULL, nl, src);
	for (int k = 0; k < 3; k++) {
		c.llo[k] = __shfl_sync(FULL, L[k], src); c.lhi[k] = __shfl_sync(FULL, L[3 + k], src);
		c.rlo[k] = __shfl_sync(FULL, R[k], src); c.rhi[k] = __shfl_sync(FULL, R[3 + k], src);
	}
	return c;
}

RTK_DEV rtk_sah_choice rtk_sah_pick_axis(const rtk_sah_choice *cand)
{
	int best = 0;                                        // all three invalid: any of them is the "no split" choice
	for (int a = 1; a < 3; a++) if (cand[a].axis >= 0 && (cand[best].axis < 0 || cand[a].cost < cand[best].cost)) best = a;
	return cand[best];
}

// depth rule of rtk.c:1429-1443: if the remaining levels cannot bring the node down to
// RTK_LEAF_MAX by halving, split it evenly now
RTK_DEV bool rtk_sah_must_halve(uint32_t count, uint32_t depth)
{
	uint32_t left = RTK_SAH_MAX_DEPTH - 1 - rtk_umin(depth, RTK_SAH_MAX_DEPTH - 1);
	if (left > 31) return false;
	return (count >> left) > RTK_LEAF_MAX;
}

// ---------------------------------------------------------------------------------------------
// per-triangle AABBs in Morton order
// ---------------------------------------------------------------------------------------------

__global__ void k_sah_prim_bounds(const float4 *tri_orig, const uint32_t *svals, uint32_t n, float4 *pb, uint32_t *idx0)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	uint32_t prim = svals[j];
	float4 a = tri_orig[3ull * prim], b = tri_orig[3ull * prim + 1], c = tri_orig[3ull * prim + 2];
	pb[2ull * j] = make_float4(rtk_fmin(rtk_fmin(a.x, b.x), c.x), rtk_fmin(rtk_fmin(a.y, b.y), c.y), rtk_fmin(rtk_fmin(a.z, b.z), c.z), 0.0f);
	pb[2ull * j + 1] = make_float4(rtk_fmax(rtk_fmax(a.x, b.x), c.x), rtk_fmax(rtk_fmax(a.y, b.y), c.y), rtk_fmax(rtk_fmax(a.z, b.z), c.z), 0.0f);
	idx0[j] = j;
}

// root node from the scene bounds; one warp (the lanes fill the root's chunk table)
__global__ void k_sah_root(rtkd_sah s, const uint32_t *bounds, uint32_t n)
{
	if (blockIdx.x) return;
	const uint32_t chunks = n > RTK_SAH_SMALL ? (n + RTK_SAH_CHUNK - 1) / RTK_SAH_CHUNK : 0u;
	for (uint32_t c = threadIdx.x; c < chunks && c < s.chunk_cap; c += blockDim.x) s.chunk_node[c] = 0;
	if (threadIdx.x) return;
	s.first[0] = 0; s.last[0] = (int)n - 1; s.left[0] = -1; s.right[0] = -1; s.ndepth[0] = 0;
	s.blo[0] = make_float4(rtk_ord2f(bounds[0]), rtk_ord2f(bounds[1]), rtk_ord2f(bounds[2]), 0.0f);
	s.bhi[0] = make_float4(rtk_ord2f(bounds[3]), rtk_ord2f(bounds[4]), rtk_ord2f(bounds[5]), 0.0f);
	for (int k = 0; k < 12; k++) s.counters[k] = 0;
	s.counters[0] = 1;
	if (n > RTK_SAH_SMALL) { s.act_in[0] = 0; s.chunk_base[0] = 0; RTK_SAH_LEVEL_NODES(s, 0) = 1; RTK_SAH_LEVEL_CHUNKS(s, 0) = chunks; }
	else { s.small_list[0] = 0; s.counters[2] = 1; }
}

// ---------------------------------------------------------------------------------------------
// large nodes, one level.  `par` = level & 1 selects the level's pair of counters (nodes, chunks); the kernels of a
// level are launched with upper-bound grids and read the pair on the device.  k_sah_split_large builds the NEXT
// level's node list, chunk table (chunk -> node, node -> first chunk) and counters with atomics as it emits the
// children -- the order of the nodes within a level is irrelevant -- so there is no planning kernel between the
// levels (a single-block scan, 4.6 us x 14 levels of a 1M-triangle build).  k_sah_bin_large zeroes the next pair
// before that: its level's predecessor, which used the pair, has finished.
// ---------------------------------------------------------------------------------------------

__global__ void k_sah_bins_clear(rtkd_sah s, int par)
{
	if (blockIdx.x >= RTK_SAH_LEVEL_NODES(s, par)) return;
	rtk_sah_bins_clear(s.bins + (size_t)blockIdx.x * RTK_SAH_NODEBINS, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256) k_sah_bin_large(rtkd_sah s, int src_buf, int par)
{
	__shared__ uint32_t s_bins[RTK_SAH_NODEBINS];
	if (blockIdx.x == 0 && threadIdx.x == 0) { RTK_SAH_LEVEL_NODES(s, par ^ 1) = 0; RTK_SAH_LEVEL_CHUNKS(s, par ^ 1) = 0; }
	if (blockIdx.x >= RTK_SAH_LEVEL_CHUNKS(s, par)) return;  // the grid is an upper bound on the level's chunks
	const uint32_t a = s.chunk_node[blockIdx.x];
	rtk_sah_bins_clear(s_bins, threadIdx.x, 256);
	__syncthreads();
	const uint32_t node = s.act_in[a];
	const uint32_t first = (uint32_t)s.first[node], last = (uint32_t)s.last[node];
	const uint32_t begin = first + (blockIdx.x - s.chunk_base[a]) * RTK_SAH_CHUNK;
	const uint32_t end = rtk_umin(begin + RTK_SAH_CHUNK, last + 1);
	const float4 nlo = s.blo[node], nhi = s.bhi[node];
	const uint32_t *idx = src_buf ? s.idx1 : s.idx0;
	// four triangles per thread at a time: the index loads, then the dependent box gathers, are all in
	// flight together before the first bin is touched (the kernel was bound by that two-load latency
	// chain, 8 times per chunk).  The trip count is warp-uniform.
	// Two ways of dealing the chunk's positions to the lanes.  Large nodes (the upper levels): a warp takes 32
	// consecutive positions, Morton neighbours, which share bins -- the warp reduces each group with REDUX.  Smaller
	// nodes: the bins are narrower than such a cluster but not by much, groups of 4-8 lanes are too small for the
	// reductions to pay and too large for the shared-memory atomics (same-address updates serialise); there a lane
	// takes positions 64 apart from its neighbours' and the warp's triangles mostly fall into different bins.
	// (Measured at 1M triangles: consecutive positions win clearly on the upper levels -- 27 against 46 us on level 1,
	// where the order used to be interleaved 8 ways -- and the two dealings are within 10 % of each other from level
	// 5 down, 34-40 us; thresholds of 8 Ki, 32 Ki and 128 Ki triangles build within 1 % of one another.)
	const bool compact = last - first + 1 >= RTK_SAH_COMPACT_MIN;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (uint32_t it = 0; it < RTK_SAH_CHUNK / 1024; it++) {
		if (begin + it * 1024 >= end && compact) break;
		uint32_t j[4], pos[4];
		bool valid[4];
		float4 lo[4], hi[4];
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const uint32_t p = compact ? begin + it * 1024 + u * 256 + threadIdx.x : begin + lane * (RTK_SAH_CHUNK / 32) + warp * 8 + it * 4 + u;
			pos[u] = p;
			valid[u] = p < end;
			j[u] = valid[u] ? idx[p] : 0u;
		}
#pragma unroll
		for (int u = 0; u < 4; u++) {
			lo[u] = make_float4(0, 0, 0, 0); hi[u] = make_float4(0, 0, 0, 0);
			if (valid[u]) { lo[u] = s.pb[2ull * j[u]]; hi[u] = s.pb[2ull * j[u] + 1]; }
		}
#pragma unroll
		for (int u = 0; u < 4; u++) {
			// the partition kernel of this level reads the bins back (coalesced) instead of gathering the boxes again
			uint32_t packed = 0;
			if (compact) packed = rtk_sah_bin_add_warp(s_bins, lo[u], hi[u], nlo, nhi, valid[u]);
			else if (valid[u]) packed = rtk_sah_bin_add(s_bins, lo[u], hi[u], nlo, nhi);
			if (valid[u]) s.binpack[pos[u]] = packed;
		}
	}
	__syncthreads();
	uint32_t *g = s.bins + (size_t)a * RTK_SAH_NODEBINS;
	for (int i = threadIdx.x; i < RTK_SAH_NODEBINS; i += 256) {
		int w = (i / RTK_SAH_BINS) & 7;
		uint32_t v = s_bins[i];
		if (w < 3) { if (v != rtk_f2ord(+RTK_INF_F)) atomicMin(g + i, v); }
		else if (w < 6) { if (v != rtk_f2ord(-RTK_INF_F)) atomicMax(g + i, v); }
		else if (w == 6 && v) atomicAdd(g + i, v);
	}
}

// children of a split node: allocate, fill, classify for the next step
// `local_alloc`: a shared-memory cursor into node numbers the caller has reserved (k_sah_small), or NULL to take two
// from the global counter
// For a child that stays large (child_buf >= 0 only) `large[k]` receives its index in the next level's node list,
// its first chunk and its number of chunks (the caller fills the chunk -> node map with all its threads).
RTK_DEV void rtk_sah_emit_children(rtkd_sah &s, uint32_t node, const rtk_sah_choice &c, uint32_t depth,
                                   int child_buf, uint32_t &child0, uint32_t *local_alloc = NULL, int par_next = 0, uint4 *large = NULL)
{
	uint32_t ch = local_alloc ? atomicAdd(local_alloc, 2u) : atomicAdd(&s.counters[0], 2u);
	if (ch + 2 > s.node_cap) { atomicOr(&s.counters[3], 1u); ch = 0; }
	child0 = ch;
	const uint32_t first = (uint32_t)s.first[node], last = (uint32_t)s.last[node];
	s.left[node] = (int)ch; s.right[node] = (int)ch + 1;
	s.first[ch] = (int)first; s.last[ch] = (int)(first + c.n_left) - 1;
	s.first[ch + 1] = (int)(first + c.n_left); s.last[ch + 1] = (int)last;
	s.blo[ch] = make_float4(c.llo[0], c.llo[1], c.llo[2], 0.0f); s.bhi[ch] = make_float4(c.lhi[0], c.lhi[1], c.lhi[2], 0.0f);
	s.blo[ch + 1] = make_float4(c.rlo[0], c.rlo[1], c.rlo[2], 0.0f); s.bhi[ch + 1] = make_float4(c.rhi[0], c.rhi[1], c.rhi[2], 0.0f);
	for (uint32_t k = 0; k < 2; k++) {
		uint32_t id = ch + k;
		uint32_t cnt = k == 0 ? c.n_left : (last - first + 1) - c.n_left;
		s.left[id] = -1; s.right[id] = -1; s.ndepth[id] = depth + 1;
		if (child_buf >= 0) {
			if (cnt > RTK_SAH_SMALL) {
				const uint32_t v = (cnt + RTK_SAH_CHUNK - 1) / RTK_SAH_CHUNK;
				const uint32_t at = atomicAdd(&RTK_SAH_LEVEL_NODES(s, par_next), 1u);
				const uint32_t c0 = atomicAdd(&RTK_SAH_LEVEL_CHUNKS(s, par_next), v);
				if (at < s.act_cap && c0 + v <= s.chunk_cap) {
					s.act_out[at] = id; s.chunk_base_out[at] = c0;
					large[k] = make_uint4(at, c0, v, 0u);
				} else atomicOr(&s.counters[3], 2u);
			} else {
				const uint32_t at = atomicAdd(&s.counters[2], 1u);
				if (at < s.small_cap) s.small_list[at] = id | ((uint32_t)child_buf << 31); else atomicOr(&s.counters[3], 4u);
			}
		}
	}
}

// one block of three warps per active large node: a warp per axis
__global__ void __launch_bounds__(96) k_sah_split_large(rtkd_sah s, uint32_t depth, int dst_buf, int par)
{
	__shared__ rtk_sah_choice s_cand[3];
	__shared__ uint4 s_large[2];
	const uint32_t a = blockIdx.x;
	if (a >= RTK_SAH_LEVEL_NODES(s, par)) return;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t node = s.act_in[a];
	const uint32_t count = (uint32_t)(s.last[node] - s.first[node] + 1);
	const float4 plo = s.blo[node], phi = s.bhi[node];
	const rtk_sah_choice mine = rtk_sah_sweep_axis(s.bins + (size_t)a * RTK_SAH_NODEBINS, warp, plo, phi, count);
	if (lane == 0) s_cand[warp] = mine;
	__syncthreads();
	if (threadIdx.x == 0) {
		rtk_sah_choice c = rtk_sah_pick_axis(s_cand);
		if (rtk_sah_must_halve(count, depth) && c.axis >= 0) {
			// forced equal split (rtk.c:1440-1443): position halves, children keep the parent's box
			rtk_sah_choice_none(c, plo, phi, count);
		}
		uint32_t ch;
		s_large[0] = make_uint4(0u, 0u, 0u, 0u); s_large[1] = make_uint4(0u, 0u, 0u, 0u);
		rtk_sah_emit_children(s, node, c, depth, dst_buf, ch, NULL, par ^ 1, s_large);
		s.split[a] = make_int4(c.axis, c.bin, (int)c.n_left, (int)ch);
		s.cursor[2 * a] = 0; s.cursor[2 * a + 1] = 0;
	}
	__syncthreads();
	// chunk -> node map of the children that stay large (up to n / 4096 entries each on the upper levels)
	for (int k = 0; k < 2; k++) {
		const uint4 lg = s_large[k];
		for (uint32_t i = threadIdx.x; i < lg.z; i += 96) s.chunk_node_out[lg.y + i] = lg.x;
	}
}

__global__ void __launch_bounds__(256) k_sah_partition_large(rtkd_sah s, int src_buf, int par)
{
	__shared__ uint32_t s_cntl[8], s_basel, s_baser;
	// the bins of the NEXT level's nodes are cleared here (their number is final once k_sah_split_large has run; this
	// level's bins have been read): a launch less per level.  There are never more than twice as many nodes on the
	// next level as there are blocks in this grid (the host launches at least one block per node of this level).
	{
		const uint32_t n_next = rtk_umin(RTK_SAH_LEVEL_NODES(s, par ^ 1), s.act_cap);
		for (uint32_t a2 = 2 * blockIdx.x; a2 < 2 * blockIdx.x + 2 && a2 < n_next; a2++)
			rtk_sah_bins_clear(s.bins + (size_t)a2 * RTK_SAH_NODEBINS, threadIdx.x, 256);
	}
	if (blockIdx.x >= RTK_SAH_LEVEL_CHUNKS(s, par)) return;
	const uint32_t a = s.chunk_node[blockIdx.x];
	const uint32_t node = s.act_in[a];
	const int4 sp = s.split[a];
	const uint32_t first = (uint32_t)s.first[node], last = (uint32_t)s.last[node];
	const uint32_t begin = first + (blockIdx.x - s.chunk_base[a]) * RTK_SAH_CHUNK;
	const uint32_t end = rtk_umin(begin + RTK_SAH_CHUNK, last + 1);
	const uint32_t *src = src_buf ? s.idx1 : s.idx0;
	uint32_t *dst = src_buf ? s.idx0 : s.idx1;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// The whole 2048-triangle chunk in one pass: every thread classifies its 8 triangles (all index
	// loads, then all box gathers, in flight together), the block scans the per-thread counts and
	// reserves its left / right ranges with ONE pair of atomics (it used to take a pair, and three
	// barriers, per 256 triangles).  The order inside a side is irrelevant to the tree but not to the next level's
	// speed: a thread takes PER CONSECUTIVE positions, so thread order is position order and the partition is
	// stable -- the Morton order survives inside each side of a chunk, and with it the warps whose triangles share
	// a bin (with positions dealt out round-robin, every level interleaved the order 8 ways and the binning of
	// levels 1-4 took twice as long as that of the sorted level 0).
	constexpr int PER = RTK_SAH_CHUNK / 256;
	uint32_t j[PER];
	uint32_t vmask = 0, lmask = 0;
#pragma unroll
	for (int u = 0; u < PER; u++) {
		const uint32_t p = begin + threadIdx.x * PER + u;
		j[u] = 0;
		if (p < end) { vmask |= 1u << u; j[u] = src[p]; }
	}
	if (sp.x < 0) {
#pragma unroll
		for (int u = 0; u < PER; u++) {
			const uint32_t p = begin + threadIdx.x * PER + u;
			if (((vmask >> u) & 1u) && (p - first) < (uint32_t)sp.z) lmask |= 1u << u;      // equal split by position
		}
	} else {
		// the triangle's bin on the split axis, as k_sah_bin_large computed it for this position (rtk.c:973-977
		// recomputes it from the box: a gather of 32 bytes per triangle that this level has already done once)
		const int sh = 8 * sp.x;
#pragma unroll
		for (int u = 0; u < PER; u++) {
			const uint32_t p = begin + threadIdx.x * PER + u;
			if (((vmask >> u) & 1u) && (int)((s.binpack[p] >> sh) & 0xffu) <= sp.y) lmask |= 1u << u;
		}
	}
	const uint32_t cl = (uint32_t)__popc(lmask), cr = (uint32_t)__popc(vmask & ~lmask);
	// exclusive scan of (cl, cr) over the block: packed in one word (a chunk holds at most 2048)
	uint32_t x = cl | (cr << 16);
	uint32_t incl = x;
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += y;
	}
	if (lane == 31) s_cntl[warp] = incl;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < 8; w++) { uint32_t v = s_cntl[w]; s_cntl[w] = t; t += v; }
		const uint32_t tl = t & 0xffffu, tr = t >> 16;
		s_basel = tl ? atomicAdd(&s.cursor[2 * a], tl) : 0;
		s_baser = tr ? atomicAdd(&s.cursor[2 * a + 1], tr) : 0;
	}
	__syncthreads();
	const uint32_t excl = s_cntl[warp] + incl - x;
	uint32_t pl = first + s_basel + (excl & 0xffffu);
	uint32_t pr = first + (uint32_t)sp.z + s_baser + (excl >> 16);
#pragma unroll
	for (int u = 0; u < PER; u++) {
		if ((vmask >> u) & 1u) {
			if ((lmask >> u) & 1u) dst[pl++] = j[u]; else dst[pr++] = j[u];
		}
	}
}

// ---------------------------------------------------------------------------------------------
// small subtrees: one CTA, everything in shared memory
// ---------------------------------------------------------------------------------------------

#ifndef RTK_SAH_SMALL_THREADS
#define RTK_SAH_SMALL_THREADS 128
#endif
#define RTK_SAH_SMALL_WARPS (RTK_SAH_SMALL_THREADS / 32)
#ifndef RTK_SAH_WARP_MAX
#define RTK_SAH_WARP_MAX 64           // subtrees of at most this many triangles are finished by one warp
#endif

struct rtk_sah_task { uint32_t node, begin, count, depth; float lo[3], hi[3]; };

// split one node whose permutation range lives in shared memory; `nthreads` threads with ids
// `tid` cooperate (a whole CTA with __syncthreads, or one warp with __syncwarp), `wl/wr` are
// nthreads/32 words of scratch.  Returns the two child tasks through a and b (valid in every thread).
template <bool WARP>
RTK_DEV void rtk_sah_split_shared(rtkd_sah &s, const rtk_sah_task &t, uint32_t *bins, const float4 *s_lo, const float4 *s_hi,
                                  unsigned short *perm0, unsigned short *perm1, rtk_sah_choice *s_choice, uint32_t *s_child,
                                  uint32_t *wl, uint32_t *wr, int tid, int nthreads, uint32_t *node_alloc, rtk_sah_task &a, rtk_sah_task &b)
{
#define RTK_SYNC() do { if (WARP) __syncwarp(); else __syncthreads(); } while (0)
	const int lane = tid & 31, warp = tid >> 5;
	const float4 nlo = make_float4(t.lo[0], t.lo[1], t.lo[2], 0.0f), nhi = make_float4(t.hi[0], t.hi[1], t.hi[2], 0.0f);
	rtk_sah_bins_clear(bins, tid, nthreads);
	RTK_SYNC();
	{
		// transposed walk: the lanes of a warp take triangles count/32 apart.  Neighbours in the (Morton-like) order
		// share bins, and 32 lanes updating 2 or 3 addresses serialise in the shared-memory atomics; triangles
		// from 32 different stretches of the range mostly fall into different bins.
		const uint32_t rows = (t.count + 31u) / 32u;
		for (uint32_t i = tid; i < rows * 32u; i += nthreads) {
			const uint32_t e = (i & 31u) * rows + (i >> 5);
			if (e < t.count) {
				const uint32_t k = perm0[t.begin + e];
				rtk_sah_bin_add(bins, s_lo[k], s_hi[k], nlo, nhi);
			}
		}
	}
	RTK_SYNC();
	rtk_sah_choice c;
	uint32_t child = 0;
	if (WARP) {
		// every lane holds the choice; only the child number travels (a shuffle, not a round trip through shared memory)
		c = rtk_sah_sweep_warp(bins, nlo, nhi, t.count);
		if (rtk_sah_must_halve(t.count, t.depth) && c.axis >= 0) rtk_sah_choice_none(c, nlo, nhi, t.count);
		if (lane == 0) rtk_sah_emit_children(s, t.node, c, t.depth, -1, child, node_alloc);
		child = __shfl_sync(0xffffffffu, child, 0);
	} else {
		// a warp per axis (s_choice[1..3] belong to the warps' own splits of phase 2: free during phase 1)
		if (warp < 3) {
			const rtk_sah_choice mine = rtk_sah_sweep_axis(bins, warp, nlo, nhi, t.count);
			if (lane == 0) s_choice[1 + warp] = mine;
		}
		__syncthreads();
		if (tid == 0) {
			rtk_sah_choice pick = rtk_sah_pick_axis(s_choice + 1);
			if (rtk_sah_must_halve(t.count, t.depth) && pick.axis >= 0) rtk_sah_choice_none(pick, nlo, nhi, t.count);
			uint32_t ch;
			rtk_sah_emit_children(s, t.node, pick, t.depth, -1, ch, node_alloc);
			*s_choice = pick; *s_child = ch;
		}
		__syncthreads();
		c = *s_choice;
		child = *s_child;
	}
	// stable partition of perm0[begin, begin+count) into perm1, then copy back
	const float amin = c.axis <= 0 ? nlo.x : (c.axis == 1 ? nlo.y : nlo.z);
	const float amax = c.axis <= 0 ? nhi.x : (c.axis == 1 ? nhi.y : nhi.z);
	uint32_t run_l = 0, run_r = 0;
	for (uint32_t base = 0; base < t.count; base += nthreads) {
		uint32_t i = base + tid;
		bool valid = i < t.count;
		unsigned short k = valid ? perm0[t.begin + i] : (unsigned short)0;
		bool goes_left = false;
		if (valid) {
			if (c.axis < 0) goes_left = i < c.n_left;
			else {
				float l = c.axis == 0 ? s_lo[k].x : (c.axis == 1 ? s_lo[k].y : s_lo[k].z);
				float h = c.axis == 0 ? s_hi[k].x : (c.axis == 1 ? s_hi[k].y : s_hi[k].z);
				goes_left = rtk_sah_bin(l, h, amin, amax) <= c.bin;
			}
		}
		uint32_t ml = __ballot_sync(0xffffffffu, valid && goes_left);
		uint32_t mr = __ballot_sync(0xffffffffu, valid && !goes_left);
		uint32_t offl = run_l, offr = run_r, totl, totr;
		if (WARP) { totl = __popc(ml); totr = __popc(mr); }
		else {
			if (lane == 0) { wl[warp] = __popc(ml); wr[warp] = __popc(mr); }
			__syncthreads();
			totl = 0; totr = 0;
			for (int w = 0; w < nthreads / 32; w++) {
				if (w < warp) { offl += wl[w]; offr += wr[w]; }
				totl += wl[w]; totr += wr[w];
			}
		}
		if (valid) {
			uint32_t lt = (1u << lane) - 1u;
			uint32_t pos = goes_left ? offl + __popc(ml & lt) : c.n_left + offr + __popc(mr & lt);
			perm1[t.begin + pos] = k;
		}
		run_l += totl; run_r += totr;
		RTK_SYNC();
	}
	for (uint32_t i = tid; i < t.count; i += nthreads) perm0[t.begin + i] = perm1[t.begin + i];
	a.node = child; a.begin = t.begin; a.count = c.n_left; a.depth = t.depth + 1;
	b.node = child + 1; b.begin = t.begin + c.n_left; b.count = t.count - c.n_left; b.depth = t.depth + 1;
	for (int k = 0; k < 3; k++) { a.lo[k] = c.llo[k]; a.hi[k] = c.lhi[k]; b.lo[k] = c.rlo[k]; b.hi[k] = c.rhi[k]; }
	RTK_SYNC();
#undef RTK_SYNC
}

// One CTA per small subtree.  Phase 1: the whole CTA splits nodes above RTK_SAH_WARP_MAX
// triangles; phase 2: the warps take the remaining subtrees from a shared queue and finish them
// independently with warp-level synchronisation only.
__global__ void __launch_bounds__(RTK_SAH_SMALL_THREADS) k_sah_small(rtkd_sah s, uint32_t n_small)
{
	__shared__ float4 s_lo[RTK_SAH_SMALL], s_hi[RTK_SAH_SMALL];
	__shared__ uint32_t s_gid[RTK_SAH_SMALL];
	__shared__ unsigned short s_perm[2][RTK_SAH_SMALL];
	__shared__ uint32_t s_bins[1 + RTK_SAH_SMALL_WARPS][RTK_SAH_NODEBINS];
	__shared__ rtk_sah_task s_stack[16];
	__shared__ rtk_sah_task s_wq[64];
	__shared__ rtk_sah_task s_wstack[RTK_SAH_SMALL_WARPS][12];
	__shared__ int s_sp, s_nwq, s_wq_next;
	__shared__ uint32_t s_node_next, s_node_end;
	__shared__ rtk_sah_choice s_choice[1 + RTK_SAH_SMALL_WARPS];
	__shared__ uint32_t s_child[1 + RTK_SAH_SMALL_WARPS], s_wl[RTK_SAH_SMALL_WARPS], s_wr[RTK_SAH_SMALL_WARPS];

	if (blockIdx.x >= n_small) return;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t entry = s.small_list[blockIdx.x];
	const uint32_t root = entry & 0x7fffffffu;
	const uint32_t *src = (entry >> 31) ? s.idx1 : s.idx0;
	const uint32_t gfirst = (uint32_t)s.first[root];
	const uint32_t total = (uint32_t)(s.last[root] - s.first[root] + 1);
	for (uint32_t i = tid; i < total; i += RTK_SAH_SMALL_THREADS) {
		uint32_t j = src[gfirst + i];
		s_gid[i] = j;
		s_lo[i] = s.pb[2ull * j]; s_hi[i] = s.pb[2ull * j + 1];
		s_perm[0][i] = (unsigned short)i;
	}
	if (tid == 0) {
		rtk_sah_task t;
		float4 lo = s.blo[root], hi = s.bhi[root];
		t.node = root; t.begin = 0; t.count = total; t.depth = s.ndepth[root];
		t.lo[0] = lo.x; t.lo[1] = lo.y; t.lo[2] = lo.z; t.hi[0] = hi.x; t.hi[1] = hi.y; t.hi[2] = hi.z;
		s_stack[0] = t;
		s_sp = 1; s_nwq = 0; s_wq_next = 0;
		// Node numbers for the whole subtree in ONE global atomic: a subtree of `total` triangles has at most
		// 2 * total - 2 nodes below its root.  (One returning atomic per split -- 125 000 of them on one address
		// for 1M triangles -- serialised the CTAs of this kernel on the L2.)  The reservations of all subtrees
		// plus the nodes of the large levels never exceed 2n - 1, so node_cap = 2n + 2 holds them; numbers that
		// stay unused are marked as empty ranges below (k_collapse_prep skips them, nothing refers to them).
		const uint32_t need = total > RTK_LEAF_MAX ? 2u * total - 2u : 0u;
		uint32_t base = need ? atomicAdd(&s.counters[0], need) : 0u;
		if (base + need > s.node_cap) { atomicOr(&s.counters[3], 1u); base = 0; }
		s_node_next = base; s_node_end = base + need;
	}
	__syncthreads();

	// ---- phase 1: whole CTA ----------------------------------------------------------------
	while (s_sp > 0) {
		const rtk_sah_task t = s_stack[s_sp - 1];
		__syncthreads();
		if (tid == 0) s_sp--;
		if (t.count <= RTK_LEAF_MAX) { __syncthreads(); continue; }          // a leaf: nothing to do
		if (t.count <= RTK_SAH_WARP_MAX) {                                    // left to one warp
			if (tid == 0) s_wq[s_nwq++] = t;
			__syncthreads();
			continue;
		}
		rtk_sah_task a, b;
		rtk_sah_split_shared<false>(s, t, s_bins[0], s_lo, s_hi, s_perm[0], s_perm[1], &s_choice[0], &s_child[0],
		                            s_wl, s_wr, tid, RTK_SAH_SMALL_THREADS, &s_node_next, a, b);
		if (tid == 0) {
			// the larger child is pushed first so that the stack stays logarithmic
			int sp = s_sp;
			if (a.count >= b.count) { s_stack[sp] = a; s_stack[sp + 1] = b; }
			else { s_stack[sp] = b; s_stack[sp + 1] = a; }
			s_sp = sp + 2;
		}
		__syncthreads();
	}
	__syncthreads();

	// ---- phase 2: one warp per remaining subtree --------------------------------------------
	const int nwq = s_nwq;
	for (;;) {
		int q = 0;
		if (lane == 0) q = atomicAdd(&s_wq_next, 1);
		q = __shfl_sync(0xffffffffu, q, 0);
		if (q >= nwq) break;
		rtk_sah_task *st = s_wstack[warp];
		int sp = 1;
		if (lane == 0) st[0] = s_wq[q];
		__syncwarp();
		while (sp > 0) {
			const rtk_sah_task t = st[sp - 1];
			__syncwarp();
			sp--;
			if (t.count <= RTK_LEAF_MAX) continue;
			rtk_sah_task a, b;
			rtk_sah_split_shared<true>(s, t, s_bins[1 + warp], s_lo, s_hi, s_perm[0], s_perm[1], &s_choice[1 + warp], &s_child[1 + warp],
			                           NULL, NULL, lane, 32, &s_node_next, a, b);
			if (lane == 0) {
				if (a.count >= b.count) { st[sp] = a; st[sp + 1] = b; }
				else { st[sp] = b; st[sp + 1] = a; }
			}
			sp += 2;
			__syncwarp();
		}
	}
	__syncthreads();
	for (uint32_t i = tid; i < total; i += RTK_SAH_SMALL_THREADS) s.idx_final[gfirst + i] = s_gid[s_perm[0][i]];
	for (uint32_t id = s_node_next + tid; id < s_node_end; id += RTK_SAH_SMALL_THREADS) {
		s.first[id] = 0; s.last[id] = -1; s.left[id] = -1; s.right[id] = -1;      // reserved, not used: an empty range
	}
}

// final leaf order as original triangle numbers
__global__ void k_sah_compose(const uint32_t *idx_final, const uint32_t *svals, uint32_t n, uint32_t *out)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = svals[idx_final[i]];
}
