/*
 * rtk.h -- public C ABI of the rtk_b200 library.
 *
 * This header is the drop-in boundary: it declares the same types, the same
 * nine entry points and the same struct layouts as the reference's public
 * header (reference rtk.h:11-130), so a program written against the reference
 * links against librtk_b200.so unchanged.  It was written from the ABI facts
 * (sizes / offsets), not copied; every declaration cites
 * the reference line it replaces.  tests/test_abi.py pins each sizeof/offsetof.
 *
 * What is different behind the boundary: the BVH build and the ray queries run
 * as CUDA kernels on an NVIDIA B200 (sm_100a).  There is no CPU implementation
 * in this library; every entry point fails loudly (NULL / false + a message on
 * stderr) when no CUDA device is usable.  The batched, data-parallel entry
 * points live in rtk_cuda.h.
 */
#ifndef RTK_B200_RTK_H
#define RTK_B200_RTK_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Largest distance a ray may carry in max_t.  Deliberately a hair below
 * FLT_MAX (reference rtk.h:11); a true +inf is outside the domain. */
#define RTK_INF (3.402823e+38f)

/* Scalar type of all coordinates (reference rtk.h:13). */
typedef float rtk_real;

/* 12 bytes, align 4.  Components are reachable by name or by index
 * (reference rtk.h:15-22). */
struct rtk_vec3 {
	union {
		rtk_real v[3];                       /* by index ...            */
		struct { rtk_real x; rtk_real y; rtk_real z; };   /* ... or by name, same storage */
	};
};
typedef struct rtk_vec3 rtk_vec3;

/* 16 bytes: a position plus the index the vertex had in its mesh's position
 * buffer (reference rtk.h:24-27). */
struct rtk_vertex {
	rtk_vec3 position;                       /* @0  */
	uint32_t index;                          /* @12 */
};
typedef struct rtk_vertex rtk_vertex;

/* 32 bytes.  The direction need not be normalised: hit distances are in units
 * of |direction| (reference rtk.h:29-34). */
struct rtk_ray {
	rtk_vec3 origin;      /* @0  */
	rtk_vec3 direction;   /* @12 */
	rtk_real min_t;       /* @24  hits need t >  min_t */
	rtk_real max_t;       /* @28  hits need t <  max_t, max_t <= RTK_INF */
};
typedef struct rtk_ray rtk_ray;

/* 68 bytes.  u is the barycentric weight of vertex[0], v of vertex[1],
 * vertex[2] carries 1-u-v (reference rtk.h:36-43). */
struct rtk_hit {
	rtk_real   t;               /* @0  */
	rtk_real   u;               /* @4  */
	rtk_real   v;               /* @8  */
	rtk_vertex vertex[3];       /* @12 */
	uint32_t   mesh_index;      /* @60 position of the mesh in rtk_scene_desc::meshes */
	uint32_t   triangle_index;  /* @64 index of the triangle inside that mesh */
};
typedef struct rtk_hit rtk_hit;

/* Element types of mesh buffers; the enumerator order is ABI
 * (reference rtk.h:45-52). */
typedef enum rtk_type {
	RTK_TYPE_DEFAULT,   /* positions: REAL, indices: U32 */
	RTK_TYPE_F32,
	RTK_TYPE_F64,
	RTK_TYPE_REAL,      /* == F32 in this build (sizeof(rtk_real) == 4) */
	RTK_TYPE_U16,
	RTK_TYPE_U32,
} rtk_type;

/* A strided view; stride 0 means tightly packed (reference rtk.h:54-58). */
struct rtk_buffer {
	const void *data;       /* @0  first element */
	size_t      stride;     /* @8  bytes between vertices / index triples, 0 = packed */
	rtk_type    type;       /* @16 */
};
typedef struct rtk_buffer rtk_buffer;

typedef struct rtk_mesh rtk_mesh;

/* Optional pull callbacks (reference rtk.h:61-62).  position_cb receives
 * 3*count vertex indices and must write 3*count positions; index_cb must write
 * the 3*count indices of triangles [offset, offset+count).  They are invoked on
 * the calling thread in chunks of at most 128 triangles. */
typedef void rtk_position_callback_fn(void *user, const rtk_mesh *mesh, rtk_vec3 *positions_out,
                                      const uint32_t *vertex_indices, size_t num_triangles);
typedef void rtk_index_callback_fn(void *user, const rtk_mesh *mesh, uint32_t *indices_out,
                                   size_t first_triangle, size_t num_triangles);

/* 96 bytes (reference rtk.h:64-76). */
struct rtk_mesh {
	void  *user;
	size_t num_triangles;

	rtk_buffer position;   /* default element type: REAL, 3 per vertex */
	rtk_buffer index;      /* default element type: U32, 3 per triangle; no data => 3i,3i+1,3i+2 */

	rtk_position_callback_fn *position_cb;
	void                     *position_cb_user;

	rtk_index_callback_fn *index_cb;
	void                  *index_cb_user;
};

/* 56-byte header at the start of every scene blob (reference rtk.h:78-89).
 * The blob is relocatable: it may be copied, written to disk and handed back
 * to rtk_trace_ray / rtk_trace_rays at a different address. */
typedef struct rtk_scene {
	char     magic[8];        /* "\0RTK\r\n\x1a\n" */
	uint16_t endian;          /* 0xaabb as written by the producer */
	uint8_t  sizeof_real;     /* 4 */
	uint8_t  pad_0;
	uint32_t version;         /* reference writes 1; this library writes 0x00B20001 */
	uint32_t pad_1;
	uint64_t size_in_bytes;
	uint64_t node_offset;
	uint64_t leaf_offset;
	uint64_t vertex_offset;
} rtk_scene;

typedef struct rtk_build    rtk_build;
typedef struct rtk_task     rtk_task;
typedef struct rtk_task_ctx rtk_task_ctx;

/* Progress messages (reference rtk.h:95). */
typedef void rtk_log_fn(void *user, rtk_build *build, const char *message);

/* 32 bytes (reference rtk.h:97-105). */
struct rtk_scene_desc {
	const rtk_mesh *meshes;       /* @0  */
	size_t          num_meshes;   /* @8  */
	rtk_log_fn     *log_fn;       /* @16 may be NULL */
	void           *log_user;     /* @24 */
};
typedef struct rtk_scene_desc rtk_scene_desc;

/* User-pumped build tasks, 40 bytes each (reference rtk.h:108-115). */
typedef void rtk_task_fn(const rtk_task *task, rtk_task_ctx *context);
struct rtk_task {
	rtk_build   *build;
	rtk_task_fn *fn;
	double       cost;
	size_t       index;
	uintptr_t    arg;
};

typedef bool rtk_filter_fn(void *user, const rtk_ray *ray, const rtk_hit *candidate);

/* Split-phase build (reference rtk.h:119-124).  rtk_start_build ingests the
 * meshes; running *first_task performs the whole GPU build synchronously and
 * queues nothing, so a task pump written for the reference terminates after
 * one call.  rtk_finish_build_to returns NULL (and keeps the build alive) when
 * the buffer is too small. */
rtk_build *rtk_start_build(const rtk_scene_desc *description, rtk_task *out_first_task);
size_t     rtk_run_task(const rtk_task *task_to_run, rtk_task *new_tasks, size_t new_tasks_capacity);
size_t     rtk_get_build_size(const rtk_build *finished_build);
rtk_scene *rtk_finish_build_to(rtk_build *finished_build, void *blob_memory, size_t blob_capacity);
rtk_scene *rtk_finish_build(rtk_build *finished_build);

/* One-shot build and release (reference rtk.h:126-127). */
rtk_scene *rtk_build_scene(const rtk_scene_desc *description);
void       rtk_free_scene(rtk_scene *scene_or_null);

/* Closest hit of one ray (reference rtk.h:129).  *hit is written only when the
 * function returns true.  Implemented as a one-ray batch on the GPU: correct,
 * but latency-bound -- use rtk_trace_rays (rtk_cuda.h) for throughput. */
bool rtk_trace_ray(const rtk_scene *scene, const rtk_ray *ray, rtk_hit *hit_out);

/* Reference rtk.h:130.  The reference ships a stub that returns true and
 * ignores its arguments (rtk.c:579-582); here: closest hit, then the filter is
 * consulted once on the host; a rejected hit reports a miss. */
bool rtk_trace_ray_filter(const rtk_scene *scene, const rtk_ray *ray, rtk_hit *hit_out,
                          rtk_filter_fn *accept, void *accept_user);

#ifdef __cplusplus
}
#endif

#endif /* RTK_B200_RTK_H */
