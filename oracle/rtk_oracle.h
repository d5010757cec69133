/*
 * rtk_oracle.h -- CPU restatement of the reference's closest-hit path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, loaded by or
 * called from the product library (librtk_b200.so).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker or as the timed CPU baseline.
 *
 * Parity pin: the scalar restatement below is checked bit-for-bit against the
 * UNMODIFIED reference leaf code (oracle/_ref/librtk_ref.so, compiled from
 * /root/reference/rtk.c with oracle/shim.h force-included) by
 * tests/test_oracle_vs_reference.py, and against the committed known-answer
 * vectors in tests/golden/ (generated from that same unmodified build by
 * tests/golden/make_golden.py).
 */
#ifndef RTK_ORACLE_H
#define RTK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same layout as rtk_ray (rtk.h:29-34) */
typedef struct orc_ray { float o[3]; float d[3]; float min_t, max_t; } orc_ray;
/* compact hit: global triangle number or 0xffffffff */
typedef struct orc_hit { float t, u, v; uint32_t prim; } orc_hit;

/* one triangle = 9 floats (v0 xyz, v1 xyz, v2 xyz) */

/* Brute force over all triangles with the canonical semantics:
 * closest t wins, exact ties -> lowest triangle number.  `threads` <= 0 uses
 * all online cores.  Follows rtk.c:543-577 (setup), :256-354 (arithmetic),
 * :366-385 (commit), with the own-lane fp64 rule. */
void orc_trace_brute(const float *tri9, size_t ntris, const orc_ray *rays, size_t nrays,
                     orc_hit *out, int threads);

/* The same answers, faster (blocked over triangles, AVX2 pre-filter of the sign test; the scalar
 * code remains the arbiter of every hit -- see rtk_oracle.c).  Returns 1 when the vector path ran,
 * 0 when it fell back to orc_trace_brute (no AVX2). */
int orc_trace_brute_blocked(const float *tri9, size_t ntris, const orc_ray *rays, size_t nrays,
                            orc_hit *out, int threads);

/* A single ray/triangle evaluation: returns 1 and fills t,u,v when the
 * triangle is accepted for (min_t, max_t). */
int orc_ray_triangle(const orc_ray *ray, const float *tri9, float max_t, float *t, float *u, float *v);

/* ---- reference-format scene (rtk.c:64-86, :1737-1765) ----------------- */

typedef struct orc_blob_stats {
	uint64_t size_bytes, num_nodes4, num_leaves, num_build_nodes;
	uint32_t max_depth;
	double   build_seconds;
} orc_blob_stats;

/* Binned-SAH build (rtk.c:867-1019 with the cost constants the header leaves
 * unset chosen as item=1, split=1) + 2->4 collapse (rtk.c:1570-1622) + packing
 * into the blob layout that rtk_trace_ray (rtk.c:543) consumes.  mesh_of /
 * tri_in_mesh / vidx give the ids stored with each triangle (may be NULL:
 * mesh 0, tri = global number, vertex index = 3*i+k).  Returns a malloc'ed,
 * 64-byte aligned blob (free with orc_free). */
void *orc_build_reference_blob(const float *tri9, size_t ntris, const uint32_t *mesh_of,
                               const uint32_t *tri_in_mesh, const uint32_t *vidx3,
                               orc_blob_stats *stats);
void orc_free(void *p);

/* One all-enclosing leaf of <= 60 triangles behind a root node: the "flat"
 * blob that drives the unmodified reference leaf code without touching the
 * defective multi-child push.  blob must hold
 * orc_flat_blob_size(n) bytes and be 64-byte aligned. */
size_t orc_flat_blob_size(size_t ntris);
void   orc_write_flat_blob(void *blob, const float *tri9, size_t ntris, uint32_t first_prim);

/* Drive a reference build's rtk_trace_ray (function pointer obtained with
 * dlsym by the caller) --------------------------------------------------- */
typedef int (*orc_ref_trace_fn)(const void *scene, const void *ray, void *hit);

/* brute force through the UNMODIFIED reference leaf test: chunks of 60
 * triangles in id order, chained through ray.max_t.  out[i].prim is the
 * global triangle number. */
void orc_trace_flat_reference(orc_ref_trace_fn fn, const float *tri9, size_t ntris,
                              const orc_ray *rays, size_t nrays, orc_hit *out, int threads);

/* full-BVH traversal through a (patched) reference build over a blob from
 * orc_build_reference_blob; out gets t,u,v and prim = global number
 * reconstructed from (mesh_index, triangle_index) via mesh_first[] (NULL: one
 * mesh).  Returns seconds spent in the timed loop (threads run concurrently;
 * wall time). */
double orc_trace_reference_blob(orc_ref_trace_fn fn, const void *blob, const orc_ray *rays,
                                size_t nrays, orc_hit *out, const uint32_t *mesh_first,
                                int threads);

int orc_num_cores(void);

#ifdef __cplusplus
}
#endif
#endif
