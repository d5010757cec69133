/*
 * rtk_device.h -- the thin C-ABI layer between the C host code (rtk_host.c) and the CUDA
 * translation unit (rtk_device.cu).  Internal: not installed, not part of the public ABI.
 * Plain C types only.
 */
#ifndef RTK_DEVICE_H
#define RTK_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A scene resident on the current device.  Pointers are device pointers. */
typedef struct rtkd_scene {
	uint64_t id;                 /* unique per process, also stored in serialised blobs */
	uint32_t num_tris, num_meshes, num_nodes, num_leaves, depth, build_mode;
	void *tri_orig;              /* float4[3*num_tris] */
	void *tv0, *tv1, *tv2;       /* float4[num_tv] each: one 8-entry slot per leaf */
	uint32_t num_tv, tv_cap;
	void *nodes;                 /* float4[16*num_nodes] */
	uint32_t nodes_cap;
	unsigned char *node_level;   /* uint8[num_nodes]: depth of each wide node (NULL for a scene loaded from a blob) */
	void *mesh_first;            /* uint32[num_meshes+1] */
	uint32_t *h_mesh_first;      /* host copy */
	float bounds_min[3], bounds_max[3], abs_max;
	double build_device_ms, build_total_ms, sah_cost;
	/* traversal scratch, created on first use */
	void *scratch;               /* counter, err, stats */
	void *overflow; size_t overflow_entries, overflow_groups;
	void *hit16; size_t hit16_cap;   /* compact hits of rtk_trace_rays_device */
	void *filter_bits;           /* uint32[(num_tris+31)/32] or NULL: triangle filter baked into the leaf slots */
} rtkd_scene;

typedef struct rtkd_trace_stats {
	uint64_t rays, hits, node_visits, leaf_visits, tri_tests, stack_max;
} rtkd_trace_stats;

int         rtkd_init(int device);           /* 0 or negative rtk_cuda_status */
int         rtkd_bind_thread(void);          /* make the library's device current on the calling thread (initialises device 0 if needed) */
void        rtkd_shutdown(void);
const char *rtkd_last_error(void);
void        rtkd_set_error(const char *fmt, ...);
int         rtkd_reserve_sms(int sms);       /* the traversal grid leaves this many SMs free */
int         rtkd_device_info(int *sm_count, size_t *l2_bytes, int *ctas_per_sm, int *threads_per_cta);

int         rtkd_read_bandwidth(size_t bytes, int passes, double *gbs);   /* read probe: L2 (small buffer) or HBM */

rtkd_scene *rtkd_scene_new(uint32_t num_tris, uint32_t num_meshes, const uint32_t *mesh_first);
void        rtkd_scene_free(rtkd_scene *s);

/* staging helpers for the host ingest */
void *rtkd_upload(const void *host, size_t bytes, void *stream);
void  rtkd_free_async(void *dev, void *stream);
int   rtkd_sync(void *stream);
int   rtkd_max_index(const void *d_idx, size_t stride, int idx_bytes, uint32_t ntris, uint32_t *out, void *stream);

/* decode one mesh (device buffers) into the scene's corner records */
int rtkd_decode_mesh(rtkd_scene *s, uint32_t first_prim, uint32_t ntris,
                     const void *d_pos, size_t pos_stride, int pos_f64,
                     const void *d_idx, size_t idx_stride, int idx_bytes, int pregathered, void *stream);

/* the same with an instance transform baked in: xf12 is a row-major 3x4 matrix (NULL = none) */
int rtkd_decode_mesh_xf(rtkd_scene *s, uint32_t first_prim, uint32_t ntris,
                        const void *d_pos, size_t pos_stride, int pos_f64,
                        const void *d_idx, size_t idx_stride, int idx_bytes, int pregathered,
                        const float *xf12, void *stream);

/* build the BVH from the decoded triangles; synchronous with respect to `stream` on return */
int rtkd_build(rtkd_scene *s, int mode, void *stream);

/* the vertices of the decoded triangles moved, the topology did not: refit all boxes in place */
int rtkd_refit(rtkd_scene *s, void *stream);

/* triangle filter: bit i of the bitset = triangle i takes part in every query on
 * this scene.  bits == NULL removes it.  on_device: bits is a device pointer.  Survives refits
 * and rebuilds of the scene. */
int rtkd_set_filter(rtkd_scene *s, const void *bits, size_t num_words, int on_device, void *stream);

/* queries: device pointers, asynchronous on stream.  cull_mode bit 0: provable culling,
 * bit 1: occlusion query (d_hit16 is then one byte per ray) */
int rtkd_trace(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, int cull_mode,
               rtkd_trace_stats *stats, void *stream);
int rtkd_trace_brute(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, void *stream);
int rtkd_resolve(rtkd_scene *s, const void *d_hit16, void *d_hits, void *d_mask, size_t n, void *stream);
void *rtkd_scene_hit16(rtkd_scene *s, size_t n);   /* scene-owned compact hit buffer of >= n records */

/* hit gather over NVLink peer memory (CUDA IPC): a window owned by the gathering process, opened
 * by the others, filled with copy-engine pushes.  handle64 is a cudaIpcMemHandle_t. */
int rtkd_peer_create(size_t bytes, void **d_window, unsigned char *handle64);
int rtkd_peer_open(const unsigned char *handle64, void **d_window);
int rtkd_peer_close(void *d_window);
int rtkd_peer_destroy(void *d_window);
int rtkd_peer_push(void *d_dst, const void *d_src, size_t bytes, void *stream);

/* wavefront ray generation: cam20 = eye, forward, right, up (3 floats each), tan(half vertical fov) */
int rtkd_gen_primary(const float *cam20, uint32_t width, uint32_t height, unsigned long long seed, uint32_t sample,
                     unsigned long long first_pixel, size_t count, void *d_rays, void *stream);
int rtkd_gen_bounce(rtkd_scene *s, const void *d_rays_in, const void *d_hit16, void *d_rays_out, void *d_alive,
                    size_t n, unsigned long long seed, uint32_t bounce, unsigned long long first_ray, uint32_t flags, void *stream);

/* host-buffer batch: H2D, trace, resolve, D2H; returns hits or -1 */
long long rtkd_trace_host(rtkd_scene *s, const void *rays, void *hits, unsigned char *mask, size_t n);

/* host-buffer batch with compact results: one 16-byte record per ray straight into the caller's array */
int rtkd_trace_host_compact(rtkd_scene *s, const void *rays, void *hit16, size_t n);

/* host placement of dense hit rows (rtk_place.c): for every ray i of the chunk with mask[i] != 0
 * the next 68-byte row of its 128-ray block -- block b's rows start at rows[block_base[b]] --
 * is copied to hits[first_ray + i]; mask_out (may be NULL) receives the mask bytes. */
typedef struct rtkd_place_desc {
	void *hits;                       /* caller's rtk_hit array (whole batch) */
	unsigned char *mask_out;          /* caller's mask array (whole batch) or NULL */
	const void *rows;                 /* dense rows of this chunk */
	const unsigned char *mask;        /* mask bytes of this chunk */
	const uint32_t *block_base;       /* first row of each 128-ray block */
	size_t first_ray, nrays;          /* the chunk's position in the batch */
} rtkd_place_desc;
int  rtkd_place_submit(const rtkd_place_desc *d);   /* returns a ticket (or -1: done synchronously) */
void rtkd_place_wait(int ticket);
int  rtkd_place_threads(void);

/* serialisation of the device layout into a relocatable blob (payload after the 128-byte
 * header block that rtk_host.c writes) */
size_t      rtkd_blob_payload_size(const rtkd_scene *s);
int         rtkd_blob_write(const rtkd_scene *s, void *payload);          /* device -> host */
rtkd_scene *rtkd_blob_read(const void *payload, size_t payload_size);     /* host -> device */

#ifdef __cplusplus
}
#endif
#endif
