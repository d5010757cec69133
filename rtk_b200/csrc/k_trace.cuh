// k_trace.cuh -- closest-hit traversal, brute-force check kernel and hit expansion (sm_100a).
//
// k_trace replaces the reference's per-ray rtk_trace_ray (rtk.c:543-577) with a persistent,
// warp-cooperative kernel:
//   * one ray per group of LANES lanes (2 by default: 16 rays per warp).  Lane c of a group owns
//     the child slots c, c+LANES, ... of the current 8-wide node -- one 256-bit load per child, the
//     lanes of a ray reading adjacent slots -- instead of 8 scattered fetches per thread;
//   * persistent CTAs pull batches of 32 rays per warp from a global counter and stage them in
//     shared memory with asynchronous copies (the next batch is in flight while the current one
//     is traced); the warp prepares a whole batch at once, one ray per lane (PD, see below);
//   * rays that reach a leaf wait until the warp tests the leaves of up to four rays together,
//     ONE triangle per lane (PD);
//   * the traversal stack (distance key + reference, 8 bytes) lives in shared memory,
//     interleaved across groups, and spills to a per-group slab in global memory when a ray
//     needs more than RTK_STACK_SMEM entries;
//   * node order: nearest hit child first, the others are pushed with their entry distance so
//     they can be culled when popped (the reference sorts all four, rtk.c:489-536);
//   * exact ties in t go to the lowest triangle number, which makes the result independent
//     of traversal order (the reference keeps whichever it met first, rtk.c:371).
//
// CULL selects the distance used to cull a box against the current best hit:
//   CULL = 1 (default, provable): entry/exit of the slab of the ray's dominant axis only.  A
//            triangle accepted by the fp32 watertight test always has its computed t inside
//            that (padded) slab interval of every box containing it, even when the triangle is
//            seen edge-on and t is numerically ill-conditioned.
//   CULL = 0: entry/exit of the full box (tighter, not provable for edge-on triangles).
#pragma once
#include "rtk_common.cuh"
#include "rtk_math.cuh"

#define RTK_TRACE_WARPS 8
#define RTK_TRACE_THREADS (RTK_TRACE_WARPS * 32)
#define RTK_GROUPS_PER_CTA_MAX (RTK_TRACE_WARPS * 16)   // 2 lanes per ray
#ifndef RTK_STACK_SMEM
#define RTK_STACK_SMEM 16
#endif
#define RTK_RAY_BATCH 32
#ifndef RTK_TRACE_MINB
#define RTK_TRACE_MINB 4                 // resident CTAs per SM the register allocation must allow
#endif
#ifndef RTK_TRI_PERIOD
#define RTK_TRI_PERIOD 5                 // the leaf phase runs every RTK_TRI_PERIOD-th iteration (or when no ray has node work)
#endif
#ifndef RTK_ASSIGN_MIN
#define RTK_ASSIGN_MIN 3                 // idle rays of a warp wait for new work until this many are idle
#endif
#ifndef RTK_PD_MIN
#define RTK_PD_MIN 4                     // PD leaf phase: run as soon as this many rays of the warp wait at a leaf
#endif
#ifndef RTK_PD_FULL_ROUNDS
#define RTK_PD_FULL_ROUNDS 1
#endif
#ifndef RTK_TRACE_LANES
#define RTK_TRACE_LANES 2                // default lanes per ray (8, 4 or 2); RTK_B200_LANES overrides at run time
#endif

struct rtkd_hit16 { float t, u, v; uint32_t prim; };

struct rtkd_trace_args {
	rtkd_arrays sc;
	const float4 *rays;          // rtk_ray = 2 x float4: (o.xyz, d.x) (d.yz, min_t, max_t)
	float4 *out;                 // rtkd_hit16 per ray
	uint32_t nrays;
	uint32_t *counter;           // global ray cursor (zeroed before launch)
	uint2 *overflow;             // [groups_in_grid * ovf_entries] stack spill
	uint32_t ovf_entries;
	uint32_t *status;            // the scene's sticky status word (host-mapped): bit 1 = stack exhausted
	unsigned long long *stats;   // [6] when STATS
};

// LANES = lanes per ray (8, 4 or 2): each lane owns 8/LANES children of the current node and
// 8/LANES triangles of the current leaf.  Fewer lanes per ray put more rays in a warp (4, 8, 16),
// which divides the per-ray serial work (setup, stack, control) by the same factor at the price of
// more independent loads per instruction.
// ANY = true turns the kernel into the occlusion (shadow-ray) query: a ray ends at the first
// triangle accepted in (min_t, max_t) and one byte per ray is written instead of a hit record.
//
// PD = true ("pair distributed" leaf phase): the (ray, triangle) pairs of all the warp's rays that
// wait at a leaf are dealt out to the 32 lanes -- four leaves of up to 8 triangles per round, each
// lane testing ONE triangle for some ray of the warp -- instead of every ray's own lanes walking
// their leaf while the lanes of the other rays idle (ncu, round 1: the leaf phase was 42 % of the
// instructions at 6 of 32 threads active).  The per-ray constants of the triangle test live in
// shared memory (written once at ray setup), candidates are min-reduced per ray with a 64-bit
// shared-memory atomicMin on (ordered t, triangle number) -- which is exactly the tie rule --
// and the winner deposits (t, u, v).  A lane's 8 triangles are 128 contiguous bytes per array, so
// a round is 3 x 4 full-line requests.
template <int LANES, int CULL, bool STATS, bool ANY = false, bool PD = false>
__global__ void __launch_bounds__(RTK_TRACE_THREADS, RTK_TRACE_MINB) k_trace(rtkd_trace_args p)
{
	constexpr int CPL = 8 / LANES;                       // children / triangles per lane
	constexpr int GW = 32 / LANES;                       // rays (groups) per warp
	constexpr int GROUPS = RTK_TRACE_WARPS * GW;         // rays per CTA
	// !PD: two raw batches (double buffer).  PD: buffer 0 receives the raw batch in flight, s_prep
	// holds the current batch after the warp-wide preparation pass (rtk_ray_prepare)
	__shared__ float4 s_rays[RTK_TRACE_WARPS][PD ? 1 : 2][RTK_RAY_BATCH * 2];
	__shared__ float4 s_prep[PD ? RTK_TRACE_WARPS : 1][PD ? RTK_RAY_BATCH * 3 : 1];
	__shared__ uint2 s_stack[RTK_STACK_SMEM][GROUPS];
	// PD: per-ray record for the distributed triangle test
	__shared__ float4 s_recA[PD ? GROUPS : 1];                // origin, min_t            (q0 of rtk_ray_prepare)
	__shared__ float4 s_recB[PD ? GROUPS : 1];                // shear sx, sy, sz, kz|sgn (q1)
	__shared__ unsigned long long s_key[PD ? GROUPS : 1];     // (ordered best t) << 32 | best triangle
	__shared__ uint32_t s_ref[PD ? GROUPS : 1];               // leaf the ray waits at
	__shared__ float4 s_hit[PD ? GROUPS : 1];                 // t, u, v of the best hit

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int c = lane & (LANES - 1), g = lane / LANES;
	const int gcta = warp * GW + g;
	const unsigned long long gglobal = (unsigned long long)blockIdx.x * GROUPS + gcta;
	const uint32_t FULL = 0xffffffffu;
	const float4 *__restrict__ nodes = p.sc.nodes;

	// ---- warp-level ray batches ------------------------------------------------------------
	uint32_t cur_base = 0, cur_cnt = 0, cur_pos = 0, next_base = 0, next_cnt = 0;
	int cur_buf = 0;
	if (PD) {
		// first batch into the raw buffer; the first pass of the assignment loop prepares it
		uint32_t b = 0;
		if (lane == 0) b = atomicAdd(p.counter, (uint32_t)RTK_RAY_BATCH);
		b = __shfl_sync(FULL, b, 0);
		next_base = b;
		next_cnt = b < p.nrays ? rtk_umin(RTK_RAY_BATCH, p.nrays - b) : 0u;
		if ((uint32_t)lane < next_cnt) {
			rtk_cp_async16(&s_rays[warp][0][lane * 2], p.rays + 2ull * (b + lane));
			rtk_cp_async16(&s_rays[warp][0][lane * 2 + 1], p.rays + 2ull * (b + lane) + 1);
		}
		rtk_cp_async_commit();
	} else {
		// first batch into buffer 0, second into buffer 1
		uint32_t b = 0;
		if (lane == 0) b = atomicAdd(p.counter, (uint32_t)RTK_RAY_BATCH);
		b = __shfl_sync(FULL, b, 0);
		cur_base = b;
		cur_cnt = b < p.nrays ? rtk_umin(RTK_RAY_BATCH, p.nrays - b) : 0u;
		if ((uint32_t)lane < cur_cnt) {
			rtk_cp_async16(&s_rays[warp][0][lane * 2], p.rays + 2ull * (b + lane));
			rtk_cp_async16(&s_rays[warp][0][lane * 2 + 1], p.rays + 2ull * (b + lane) + 1);
		}
		rtk_cp_async_commit();
		if (cur_cnt == RTK_RAY_BATCH) {
			if (lane == 0) b = atomicAdd(p.counter, (uint32_t)RTK_RAY_BATCH);
			b = __shfl_sync(FULL, b, 0);
			next_base = b;
			next_cnt = b < p.nrays ? rtk_umin(RTK_RAY_BATCH, p.nrays - b) : 0u;
			if ((uint32_t)lane < next_cnt) {
				rtk_cp_async16(&s_rays[warp][PD ? 0 : 1][lane * 2], p.rays + 2ull * (b + lane));
				rtk_cp_async16(&s_rays[warp][PD ? 0 : 1][lane * 2 + 1], p.rays + 2ull * (b + lane) + 1);
			}
			rtk_cp_async_commit();
		}
		rtk_cp_async_wait_all();
		__syncwarp();
	}

	// ---- per-ray state (identical in the lanes of a group) ---------------------------------
	rtk_ray_ctx rc;
	bool has_ray = false;
	uint32_t ray_index = 0;
	float best_t = 0.0f, best_u = 0.0f, best_v = 0.0f;
	uint32_t best_prim = RTK_MISS;
	uint32_t cur_ref = RTK_REF_EMPTY;
	int sp = 0;
	uint32_t st_nodes = 0, st_leaves = 0, st_tris = 0, st_stack = 0;
	int tri_tick = 0;

#define RTK_STACK_WRITE(pos, val) do { \
		int _p = (pos); \
		if (_p < RTK_STACK_SMEM) s_stack[_p][gcta] = (val); \
		else if ((uint32_t)(_p - RTK_STACK_SMEM) < p.ovf_entries) p.overflow[gglobal * p.ovf_entries + (uint32_t)(_p - RTK_STACK_SMEM)] = (val); \
		else *(volatile uint32_t*)p.status = 2u; \
	} while (0)
#define RTK_STACK_POP() do { \
		cur_ref = RTK_REF_EMPTY; \
		if (sp <= RTK_STACK_SMEM) { \
			/* common case: everything the ray has pushed lives in shared memory */ \
			while (sp > 0) { \
				--sp; \
				const uint2 _e = s_stack[sp][gcta]; \
				if (__uint_as_float(_e.x) <= best_t) { cur_ref = _e.y; break; } \
			} \
		} else while (sp > 0) { \
			--sp; \
			uint2 _e; \
			if (sp < RTK_STACK_SMEM) _e = s_stack[sp][gcta]; \
			else if ((uint32_t)(sp - RTK_STACK_SMEM) < p.ovf_entries) _e = __ldcg(&p.overflow[gglobal * p.ovf_entries + (uint32_t)(sp - RTK_STACK_SMEM)]); \
			else continue; \
			if (__uint_as_float(_e.x) <= best_t) { cur_ref = _e.y; break; } \
		} \
	} while (0)

	for (;;) {
		// ---- (1) hand rays to idle groups --------------------------------------------------
		uint32_t need_mask = __ballot_sync(FULL, !has_ray && c == 0);
#if RTK_ASSIGN_MIN > 1
		// ray setup is serial per ray but costs a warp-wide instruction stream: let finished rays
		// wait until a few of the warp's rays are idle, then set them up together
		if (__popc(need_mask) < RTK_ASSIGN_MIN && __popc(need_mask) < GW) need_mask = 0;
#endif
		while (need_mask) {
			uint32_t avail = cur_cnt - cur_pos;
			if (avail == 0) {
				if (next_cnt == 0) break;                     // input exhausted
				rtk_cp_async_wait_all();
				__syncwarp();
				if (PD) {
					// the staged batch becomes current: every lane prepares one of its rays
					cur_base = next_base; cur_cnt = next_cnt; cur_pos = 0;
					next_cnt = 0;
					if ((uint32_t)lane < cur_cnt) {
						float4 q0, q1, q2;
						rtk_ray_prepare(s_rays[warp][0][lane * 2], s_rays[warp][0][lane * 2 + 1], q0, q1, q2);
						s_prep[warp][lane * 3] = q0; s_prep[warp][lane * 3 + 1] = q1; s_prep[warp][lane * 3 + 2] = q2;
					}
					__syncwarp();
					if (cur_cnt == RTK_RAY_BATCH) {
						uint32_t b = 0;
						if (lane == 0) b = atomicAdd(p.counter, (uint32_t)RTK_RAY_BATCH);
						b = __shfl_sync(FULL, b, 0);
						next_base = b;
						next_cnt = b < p.nrays ? rtk_umin(RTK_RAY_BATCH, p.nrays - b) : 0u;
						if ((uint32_t)lane < next_cnt) {
							rtk_cp_async16(&s_rays[warp][0][lane * 2], p.rays + 2ull * (b + lane));
							rtk_cp_async16(&s_rays[warp][0][lane * 2 + 1], p.rays + 2ull * (b + lane) + 1);
						}
						rtk_cp_async_commit();
					}
					continue;
				}
				cur_buf ^= 1; cur_base = next_base; cur_cnt = next_cnt; cur_pos = 0;
				next_cnt = 0;
				if (cur_cnt == RTK_RAY_BATCH) {
					uint32_t b = 0;
					if (lane == 0) b = atomicAdd(p.counter, (uint32_t)RTK_RAY_BATCH);
					b = __shfl_sync(FULL, b, 0);
					next_base = b;
					next_cnt = b < p.nrays ? rtk_umin(RTK_RAY_BATCH, p.nrays - b) : 0u;
					if ((uint32_t)lane < next_cnt) {
						rtk_cp_async16(&s_rays[warp][PD ? 0 : (cur_buf ^ 1)][lane * 2], p.rays + 2ull * (b + lane));
						rtk_cp_async16(&s_rays[warp][PD ? 0 : (cur_buf ^ 1)][lane * 2 + 1], p.rays + 2ull * (b + lane) + 1);
					}
					rtk_cp_async_commit();
				}
				continue;
			}
			uint32_t rank = __popc(need_mask & ((1u << (g * LANES)) - 1u));
			if (!has_ray && rank < avail) {
				uint32_t slot = cur_pos + rank;
				if (PD) {
					const float4 q0 = s_prep[warp][slot * 3], q1 = s_prep[warp][slot * 3 + 1], q2 = s_prep[warp][slot * 3 + 2];
					rtk_ray_node_ctx(rc, q0, q1, q2, p.sc.abs_max);
					best_t = q2.w;                                                       // max_t, rtk.c:548
					if (c == 0) { s_recA[gcta] = q0; s_recB[gcta] = q1; }
				} else {
					float4 r0 = s_rays[warp][PD ? 0 : cur_buf][slot * 2], r1 = s_rays[warp][PD ? 0 : cur_buf][slot * 2 + 1];
					rtk_ray_setup(rc, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, p.sc.abs_max);
					best_t = r1.w;
				}
				ray_index = cur_base + slot;
				best_u = 0.0f; best_v = 0.0f; best_prim = RTK_MISS;
				sp = 0;
				cur_ref = p.sc.num_nodes ? 0u : RTK_REF_EMPTY;                        // root = node 0
				has_ray = true;
			}
			cur_pos += rtk_umin((uint32_t)__popc(need_mask), avail);
			need_mask = __ballot_sync(FULL, !has_ray && c == 0);
		}
		if (!__any_sync(FULL, has_ray)) break;

		// ---- (2) leaf: the group tests up to 8 triangles (rtk.c:181-388) ---------------------
		const bool is_leaf = has_ray && cur_ref != RTK_REF_EMPTY && rtk_ref_is_leaf(cur_ref);
		if (PD) {
			uint32_t lm = __ballot_sync(FULL, is_leaf && c == 0);
			const int nl = __popc(lm);
			const bool any_node0 = __any_sync(FULL, has_ray && cur_ref != RTK_REF_EMPTY && !rtk_ref_is_leaf(cur_ref));
			// a round serves four leaves: run when a full round is waiting, or when waiting longer
			// would only idle lanes (no node work) or has lasted RTK_TRI_PERIOD iterations
			const bool force = !any_node0 || (nl && ++tri_tick >= RTK_TRI_PERIOD);
			if (nl >= RTK_PD_MIN || (nl && force)) {
				tri_tick = 0;
				if (!force && RTK_PD_FULL_ROUNDS) {
					// full rounds only: the last (nl mod 4) rays keep waiting
					for (int r = nl & 3; r > 0; r--) lm &= ~(0x80000000u >> __clz((int)lm));
				}
				const bool mine = is_leaf && ((lm >> (g * LANES)) & 1u);
				if (mine && c == 0) {
					s_key[gcta] = ((unsigned long long)rtk_f2ord(best_t) << 32) | best_prim;
					s_ref[gcta] = cur_ref;
				}
				__syncwarp();
				const int q = lane >> 3, k = lane & 7;
				while (lm) {
					// lanes 8q..8q+7 take the q-th waiting ray
					uint32_t m = lm;
					int b = -1;
#pragma unroll
					for (int j = 0; j < 4; j++) {
						if (q == j && m) b = __ffs((int)m) - 1;
						m &= m - 1u;
					}
					lm = m;
					bool acc = false;
					unsigned long long mykey = 0;
					float tt = 0.0f, uu = 0.0f, vv = 0.0f;
					int slot = 0;
					if (b >= 0) {
						slot = warp * GW + b / LANES;
						const uint32_t ref = s_ref[slot];
						const uint32_t first = rtk_leaf_first(ref), cnt = rtk_leaf_count(ref);
						if ((uint32_t)k < cnt) {
							const unsigned long long key0 = s_key[slot];
							const float bt = rtk_ord2f((uint32_t)(key0 >> 32));
							const uint32_t bp = (uint32_t)key0;
							const float4 A = s_recA[slot], B = s_recB[slot];
							rtk_ray_ctx tc;
							rtk_ray_tri_ctx(tc, A, B);
							float4 p0 = __ldg(&p.sc.tv0[first + k]);
							float4 p1 = __ldg(&p.sc.tv1[first + k]);
							float4 p2 = __ldg(&p.sc.tv2[first + k]);
							const uint32_t id = __float_as_uint(p0.w);
							// strict '<' against the recorded best (rtk.c:354, 371); an exact tie is taken
							// only from a lower triangle number than the recorded hit
							if (rtk_tri_test(tc, p0, p1, p2, bt, tt, uu, vv) && (tt < bt || (bp != RTK_MISS && id < bp))) {
								acc = true;
								// + 0.0f: -0.0 and +0.0 are the same distance
								mykey = ((unsigned long long)rtk_f2ord(__fadd_rn(tt, 0.0f)) << 32) | id;
								atomicMin(&s_key[slot], mykey);
							}
						}
					}
					__syncwarp();
					if (acc && s_key[slot] == mykey) s_hit[slot] = make_float4(tt, uu, vv, 0.0f);
				}
				__syncwarp();
				if (mine) {
					const unsigned long long key = s_key[gcta];
					best_t = rtk_ord2f((uint32_t)(key >> 32));
					best_prim = (uint32_t)key;
					if (STATS) { st_leaves++; st_tris += rtk_leaf_count(cur_ref); }
					if (ANY && best_prim != RTK_MISS) { sp = 0; cur_ref = RTK_REF_EMPTY; }   // occluded: nothing else matters
					else RTK_STACK_POP();
				}
				__syncwarp();
			}
		} else {
#if RTK_TRI_PERIOD > 1
		// Both phases cost a full warp instruction stream however few of the warp's rays take part.
		// With 16 rays per warp about 1 in 5 is at a leaf at any time, so the leaf phase is run only
		// every RTK_TRI_PERIOD-th iteration: rays that reach a leaf wait (idle lanes, no issue slots,
		// no speculative traversal) and are then tested together.  It still runs at once when no
		// ray of the warp has node work.
		const bool any_leaf = __any_sync(FULL, is_leaf);
		const bool any_node0 = __any_sync(FULL, has_ray && cur_ref != RTK_REF_EMPTY && !rtk_ref_is_leaf(cur_ref));
		if (any_leaf && (!any_node0 || ++tri_tick >= RTK_TRI_PERIOD)) {
			tri_tick = 0;
#else
		if (__any_sync(FULL, is_leaf)) {
#endif
			float t = INFINITY, u = 0.0f, v = 0.0f;
			uint32_t prim = RTK_MISS;
			if (is_leaf) {
				const uint32_t first = rtk_leaf_first(cur_ref), cnt = rtk_leaf_count(cur_ref);
				// running best of this lane: starts at the ray's best so that each later triangle
				// only has to beat what an earlier one of the same lane found
				float lt = best_t;
				uint32_t lp = best_prim;
#pragma unroll
				for (int j = 0; j < CPL; j++) {
					const uint32_t k = (uint32_t)(c * CPL + j);
					if (k < cnt) {
						float4 p0 = __ldg(&p.sc.tv0[first + k]);
						float4 p1 = __ldg(&p.sc.tv1[first + k]);
						float4 p2 = __ldg(&p.sc.tv2[first + k]);
						uint32_t id = __float_as_uint(p0.w);
						float tt, uu, vv;
						if (rtk_tri_test(rc, p0, p1, p2, lt, tt, uu, vv)) {
							// strict '<' against the running best (rtk.c:354, 371); an exact tie is
							// taken only from a lower triangle number than the recorded hit
							if (tt < lt || (lp != RTK_MISS && id < lp)) { lt = tt; lp = id; t = tt; u = uu; v = vv; prim = id; }
						}
					}
				}
				if (STATS) { st_leaves++; st_tris += cnt; }
			}
			float tmin = t;
#pragma unroll
			for (int o = 1; o < LANES; o <<= 1) tmin = rtk_fmin(tmin, __shfl_xor_sync(FULL, tmin, o));
			uint32_t pmin = (t == tmin) ? prim : RTK_MISS;
#pragma unroll
			for (int o = 1; o < LANES; o <<= 1) pmin = rtk_umin(pmin, __shfl_xor_sync(FULL, pmin, o));
			const bool winner = prim != RTK_MISS && t == tmin && prim == pmin;
			uint32_t wm = (__ballot_sync(FULL, winner) >> (g * LANES)) & ((1u << LANES) - 1u);
			int src = wm ? __ffs(wm) - 1 : 0;
			float wu = __shfl_sync(FULL, u, src, LANES), wv = __shfl_sync(FULL, v, src, LANES);
			if (is_leaf) {
				if (wm) { best_t = tmin; best_prim = pmin; best_u = wu; best_v = wv; }
				if (ANY && wm) { sp = 0; cur_ref = RTK_REF_EMPTY; }       // occluded: nothing else matters
				else RTK_STACK_POP();
			}
			__syncwarp();
		}
		}   // !PD

		// ---- (3) node: the group tests the 8 children (rtk.c:457-473) ------------------------
		const bool is_node = has_ray && cur_ref != RTK_REF_EMPTY && !rtk_ref_is_leaf(cur_ref);
		if (__any_sync(FULL, is_node)) {
			uint32_t hitbits = 0;                 // bit (child index) for the children of this lane
			uint32_t ok = 0x7fffffffu;            // best ordering key of this lane (a signed comparison, see below) ...
			uint32_t okref = RTK_REF_EMPTY;       // ... and the child it belongs to
			float key[CPL];
			uint32_t ref[CPL];
#pragma unroll
			for (int j = 0; j < CPL; j++) { key[j] = 0.0f; ref[j] = RTK_REF_EMPTY; }
			if (is_node) {
				const float4 *np = nodes + 16ull * cur_ref;
				const bool nx = rc.sgn & 1u, ny = rc.sgn & 2u, nz = rc.sgn & 4u;
#pragma unroll
				for (int j = 0; j < CPL; j++) {
					const int k = j * LANES + c;             // child slot: the lanes of a ray read adjacent slots
					float4 lo, hi;
					rtk_ldg256(np + 2 * k, lo, hi);
					ref[j] = __float_as_uint(lo.w);
					float tnx = fmaf(nx ? hi.x : lo.x, rc.idx, rc.cnx), tfx = fmaf(nx ? lo.x : hi.x, rc.idx, rc.cfx);
					float tny = fmaf(ny ? hi.y : lo.y, rc.idy, rc.cny), tfy = fmaf(ny ? lo.y : hi.y, rc.idy, rc.cfy);
					float tnz = fmaf(nz ? hi.z : lo.z, rc.idz, rc.cnz), tfz = fmaf(nz ? lo.z : hi.z, rc.idz, rc.cfz);
					float tn = rtk_fmax(rtk_fmax(tnx, tny), tnz);
					float tf = rtk_fmin(rtk_fmin(tfx, tfy), tfz);
					float kn, kf;
					if (CULL == 1) {
						kn = rc.kz == 0 ? tnx : (rc.kz == 1 ? tny : tnz);
						kf = rc.kz == 0 ? tfx : (rc.kz == 1 ? tfy : tfz);
					} else { kn = tn; kf = tf; }
					// an empty slot is an inverted box (lo = +RTK_INF, hi = -RTK_INF): tn <= tf already fails for it, so
					// the slot's reference needs no test of its own (measured: 9.84 -> 9.69 ms per 16.7M C3 rays)
					const bool hit = tn <= tf && kn <= best_t && kf >= rc.min_t;
					key[j] = kn;
					if (hit) {
						hitbits |= 1u << (j * LANES);              // a constant per j: the lane's offset c is added once, after the loop
						// nearest hit child: entry distance with the child number in the low 3 bits
						// (the reference tags 2 bits the same way, rtk.c:496)
						// compared as SIGNED integers: a negative entry distance (the origin is inside the box) sorts
						// before every positive one without being clamped to zero first (9.69 -> 9.55 ms per 16.7M C3 rays;
						// the order among children that contain the origin is irrelevant to the result)
						const uint32_t kk = (__float_as_uint(tn) & ~7u) | (uint32_t)k;
						if ((int)kk < (int)ok) { ok = kk; okref = ref[j]; }
					}
				}
				if (STATS) st_nodes++;
			}
			hitbits <<= c;                            // (9.55 -> 9.42 ms per 16.7M C3 rays against a shift per child)
			uint32_t om = ok, gm = hitbits;
#pragma unroll
			for (int o = 1; o < LANES; o <<= 1) {
				om = (uint32_t)rtk_imin((int)om, (int)__shfl_xor_sync(FULL, om, o));
				if (CPL > 1) gm |= __shfl_xor_sync(FULL, gm, o);
			}
			if (CPL == 1) gm = (__ballot_sync(FULL, hitbits != 0) >> (g * LANES)) & 0xffu;
			const int cmin = (int)(om & 7u);
			const uint32_t nref = __shfl_sync(FULL, okref, cmin & (LANES - 1), LANES);
			if (is_node) {
				const uint32_t others = gm & ~(1u << cmin);
				const int npush = __popc(others);
				if (sp + npush <= RTK_STACK_SMEM) {
					// common case: the pushes fit in shared memory, no bounds checks
					// position of child k = j * LANES + c among the pushed ones: those below the lane's offset c, plus those of
					// the mask shifted down by c below bit j * LANES -- a constant mask per j (9.42 -> 9.35 ms per 16.7M C3 rays
					// against a variable shift per child)
					const uint32_t osh = others >> c;
					const int pbase = sp + __popc(others & ((1u << c) - 1u));
#pragma unroll
					for (int j = 0; j < CPL; j++) {
						if ((osh >> (j * LANES)) & 1u) s_stack[pbase + __popc(osh & ((1u << (j * LANES)) - 1u))][gcta] = make_uint2(__float_as_uint(key[j]), ref[j]);
					}
				} else {
#pragma unroll
					for (int j = 0; j < CPL; j++) {
						const int k = j * LANES + c;
						if ((others >> k) & 1u) {
							int pos = sp + __popc(others & ((1u << k) - 1u));
							RTK_STACK_WRITE(pos, make_uint2(__float_as_uint(key[j]), ref[j]));
						}
					}
				}
				sp += npush;
				if (STATS) st_stack = rtk_umax(st_stack, (uint32_t)sp);
			}
			__syncwarp();
			if (is_node) {
				if (gm) cur_ref = nref;
				else RTK_STACK_POP();
			}
			__syncwarp();
		}

		// ---- (4) finished rays --------------------------------------------------------------
		if (has_ray && cur_ref == RTK_REF_EMPTY) {
			if (c == 0) {
				const bool got = best_prim != RTK_MISS;                 // a hit is only ever recorded below max_t (rtk.c:571)
				if (ANY) ((unsigned char*)p.out)[ray_index] = got ? 1 : 0;
				else {
					float4 o = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(RTK_MISS));
					if (got) {
						if (PD) { o = s_hit[gcta]; o.w = __uint_as_float(best_prim); }
						else o = make_float4(best_t, best_u, best_v, __uint_as_float(best_prim));
					}
					p.out[ray_index] = o;
				}
				if (STATS) {
					atomicAdd(&p.stats[1], (unsigned long long)(got ? 1 : 0));
					atomicAdd(&p.stats[2], (unsigned long long)st_nodes);
					atomicAdd(&p.stats[3], (unsigned long long)st_leaves);
					atomicAdd(&p.stats[4], (unsigned long long)st_tris);
					atomicMax(&p.stats[5], (unsigned long long)st_stack);
				}
			}
			st_nodes = st_leaves = st_tris = st_stack = 0;
			has_ray = false;
		}
	}
#undef RTK_STACK_WRITE
#undef RTK_STACK_POP
}

// ---------------------------------------------------------------------------------------------
// Exhaustive check kernel: every ray against every triangle, same arithmetic, same tie rule.
// One ray per thread; triangles are staged through shared memory 128 at a time.
// ---------------------------------------------------------------------------------------------

#define RTK_BRUTE_TILE 128

__global__ void __launch_bounds__(128) k_trace_brute(rtkd_arrays sc, const float4 *rays, float4 *out, uint32_t nrays)
{
	__shared__ float4 s_v0[RTK_BRUTE_TILE], s_v1[RTK_BRUTE_TILE], s_v2[RTK_BRUTE_TILE];
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	rtk_ray_ctx rc;
	float best_t = 0.0f, best_u = 0.0f, best_v = 0.0f, max_t = 0.0f;
	uint32_t best_prim = RTK_MISS;
	bool live = i < nrays;
	if (live) {
		float4 r0 = rays[2ull * i], r1 = rays[2ull * i + 1];
		rtk_ray_setup(rc, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, sc.abs_max);
		best_t = max_t = r1.w;
	}
	for (uint32_t base = 0; base < sc.num_tv; base += RTK_BRUTE_TILE) {
		uint32_t k = base + threadIdx.x;
		__syncthreads();
		if (threadIdx.x < RTK_BRUTE_TILE && k < sc.num_tv) {
			s_v0[threadIdx.x] = sc.tv0[k]; s_v1[threadIdx.x] = sc.tv1[k]; s_v2[threadIdx.x] = sc.tv2[k];
		}
		__syncthreads();
		uint32_t cnt = rtk_umin(RTK_BRUTE_TILE, sc.num_tv - base);
		if (live) {
			for (uint32_t j = 0; j < cnt; j++) {
				float t, u, v;
				const uint32_t id = __float_as_uint(s_v0[j].w);
				if (id != RTK_MISS && rtk_tri_test(rc, s_v0[j], s_v1[j], s_v2[j], best_t, t, u, v)) {
					if (t < best_t || (best_prim != RTK_MISS && id < best_prim)) {
						best_t = t; best_u = u; best_v = v; best_prim = id;
					}
				}
			}
		}
	}
	if (live) {
		bool got = best_t < max_t;
		out[i] = got ? make_float4(best_t, best_u, best_v, __uint_as_float(best_prim))
		             : make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(RTK_MISS));
	}
}

// ---------------------------------------------------------------------------------------------
// Hit expansion: compact record -> rtk_hit (68 bytes), reference rtk.c:371-381.  A block stages
// its 128 hits in shared memory and writes the 17 words of each with coalesced stores; rows of
// rays that missed are left untouched (rtk.c:571-576).
// ---------------------------------------------------------------------------------------------

#define RTK_RESOLVE_THREADS 128

// DENSE: the rows of the rays that hit are packed instead -- block b's rows are contiguous from
// row block_base[b] of `hits`, blocks in the order they reserved their range from *hit_count --
// which is what the host-buffer path sends over PCIe (rtk_place.c puts them back in place).
template <bool DENSE>
__global__ void __launch_bounds__(RTK_RESOLVE_THREADS) k_resolve(rtkd_arrays sc, const float4 *hit16, uint32_t *hits,
                                                                 unsigned char *mask, uint32_t nrays,
                                                                 unsigned long long *hit_count, uint32_t *block_base)
{
	__shared__ uint32_t s_row[RTK_RESOLVE_THREADS][17];
	__shared__ unsigned char s_hit[RTK_RESOLVE_THREADS];
	__shared__ uint32_t s_warp[RTK_RESOLVE_THREADS / 32 + 1];
	const uint32_t base = blockIdx.x * RTK_RESOLVE_THREADS;
	const uint32_t i = base + threadIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	bool got = false;
	float4 h = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	if (i < nrays) {
		h = hit16[i];
		got = __float_as_uint(h.w) != RTK_MISS;
		if (mask) mask[i] = got ? 1 : 0;
	}
	const uint32_t m = __ballot_sync(0xffffffffu, got);
	uint32_t slot = threadIdx.x;                 // row of the staging buffer this thread fills
	if (DENSE) {
		if (lane == 0) s_warp[warp] = (uint32_t)__popc(m);
		__syncthreads();
		uint32_t before = 0, total = 0;
		for (int w = 0; w < RTK_RESOLVE_THREADS / 32; w++) { uint32_t c = s_warp[w]; if (w < warp) before += c; total += c; }
		slot = before + (uint32_t)__popc(m & ((1u << lane) - 1u));
		if (threadIdx.x == 0) {
			uint32_t b = total ? (uint32_t)atomicAdd(hit_count, (unsigned long long)total) : 0u;
			s_warp[RTK_RESOLVE_THREADS / 32] = b;
			block_base[blockIdx.x] = b;
		}
	} else if (hit_count && lane == 0 && m) atomicAdd(hit_count, (unsigned long long)__popc(m));
	if (got) {
		const uint32_t prim = __float_as_uint(h.w);
		// mesh of the triangle: mesh_first is a short sorted table
		uint32_t lo = 0, hi = sc.num_meshes;
		while (hi - lo > 1) {
			uint32_t mid = (lo + hi) >> 1;
			if (sc.mesh_first[mid] <= prim) lo = mid; else hi = mid;
		}
		const float4 *tv = sc.tri_orig + 3ull * prim;
		float4 a = __ldg(tv), b = __ldg(tv + 1), cc = __ldg(tv + 2);
		uint32_t *r = s_row[slot];
		r[0] = __float_as_uint(h.x); r[1] = __float_as_uint(h.y); r[2] = __float_as_uint(h.z);
		r[3] = __float_as_uint(a.x); r[4] = __float_as_uint(a.y); r[5] = __float_as_uint(a.z); r[6] = __float_as_uint(a.w);
		r[7] = __float_as_uint(b.x); r[8] = __float_as_uint(b.y); r[9] = __float_as_uint(b.z); r[10] = __float_as_uint(b.w);
		r[11] = __float_as_uint(cc.x); r[12] = __float_as_uint(cc.y); r[13] = __float_as_uint(cc.z); r[14] = __float_as_uint(cc.w);
		r[15] = lo;
		r[16] = prim - sc.mesh_first[lo];
	}
	s_hit[threadIdx.x] = got ? 1 : 0;
	__syncthreads();
	if (DENSE) {
		uint32_t total = 0;
		for (int w = 0; w < RTK_RESOLVE_THREADS / 32; w++) total += s_warp[w];
		uint32_t *dst = hits + 17ull * s_warp[RTK_RESOLVE_THREADS / 32];
		const uint32_t *src = &s_row[0][0];
		for (uint32_t w = threadIdx.x; w < total * 17u; w += RTK_RESOLVE_THREADS) dst[w] = src[w];
	} else {
		uint32_t rows = rtk_umin(RTK_RESOLVE_THREADS, nrays > base ? nrays - base : 0u);
		uint32_t *dst = hits + 17ull * base;
		for (uint32_t w = threadIdx.x; w < rows * 17u; w += RTK_RESOLVE_THREADS) {
			uint32_t row = w / 17u;
			if (s_hit[row]) dst[w] = s_row[row][w - row * 17u];
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Row push of the direct host path: rows expanded in device memory (k_resolve, rows in place at a
// 68-byte pitch) go to the caller's page-locked rtk_hit array, hit rows only, and the mask bytes to the
// caller's mask array.  A handful of persistent CTAs (they run on the SMs the traversal grid leaves
// free) loop over the chunk's 128-ray blocks; a block's 2176 words are written 32 consecutive words
// per warp instruction -- one aligned 128-byte line of host memory, minus the words of rays that missed.
// ---------------------------------------------------------------------------------------------

#define RTK_PUSH_THREADS 256

__global__ void __launch_bounds__(RTK_PUSH_THREADS) k_push_rows(const uint32_t *rows, const unsigned char *mask, uint32_t *host_hits,
                                                                unsigned char *host_mask, uint32_t nrays)
{
	__shared__ unsigned char s_m[RTK_RESOLVE_THREADS];
	const uint32_t nblocks = (nrays + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS;
	for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
		const uint32_t base = b * RTK_RESOLVE_THREADS;
		const uint32_t cnt = rtk_umin(RTK_RESOLVE_THREADS, nrays - base);
		__syncthreads();
		if (threadIdx.x < RTK_RESOLVE_THREADS) {
			const unsigned char m = threadIdx.x < cnt ? mask[base + threadIdx.x] : (unsigned char)0;
			s_m[threadIdx.x] = m;
			if (host_mask && threadIdx.x < cnt) host_mask[base + threadIdx.x] = m;
		}
		__syncthreads();
		const uint32_t *src = rows + 17ull * base;
		uint32_t *dst = host_hits + 17ull * base;
		for (uint32_t w = threadIdx.x; w < cnt * 17u; w += RTK_PUSH_THREADS) {
			if (s_m[w / 17u]) dst[w] = src[w];
		}
	}
}
